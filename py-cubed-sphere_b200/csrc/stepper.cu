// Host-side driver of the fused advection step (src/advection_timestep.py:19-43 for the ET-DG schemes):
// sequencing of ghost fill, step kernel, multi-GPU exchange; deferred MF-PR projection; separable
// wind; CUDA-graph replay; and the small kernels around the step kernel (Lagrange ghost fill,
// projection flush, ring restore).
//
// Three shapes of a step:
//
//   serial (one GPU)       ghost fill (folds the pending MF-PR term into the ghost cells, hands the
//                          coefficient to the step kernel) -> step kernel over the whole grid; the two are
//                          chained by a programmatic dependent launch inside the replayed graph.
//
//   one kernel (2 GPUs;    ONE launch of the GH = 2 flavour: a CTA whose staged rectangle touches a panel edge
//   PYCS_ONEKERNEL=1)      fills those ghost cells itself first (on a sharded handle after the peers' dflags),
//                          the boundary CTAs ship their rows to the peers at the end of their march.  Uniform
//                          CTA table, a single graph node per step.
//
//   split (4 / 8 GPUs,     the handle's stream (high priority):  boundary CTAs (GH = 1, they ship their rows to the
//   PYCS_SPLIT=1 on one)                             peers themselves) -> ghost fill of the NEXT step (raw: no projection term)
//                          second stream:            interior CTAs (GH = 0)
//                          The boundary CTAs are everything a ghost cell or a peer reads: first / last
//                          column strip of every panel, first / last chunk of the slab.  Exchange,
//                          flag latency and ghost fill run beside the interior update; the two launches
//                          share the partial sums and the ticket, whichever finishes last closes the
//                          step (publishes the MF-PR sum, raises sflag).  Both wait in-kernel for the
//                          peers' sflags; the ghost fill waits for their dflags.
//
// MF-PR (src/discrete_operators.py:98-101) needs a global sum: a step kernel writes Q - dt*div and the
// sum of pxdF + pydF; the projection term sqrtg * (-sum / a2) is added when the next consumer loads Q
// (the next step, or the flush kernel before anything else reads Q).  Algebraically identical, one
// rounding apart from the reference order.
//
// Everything that differs from one step to the next lives in the device-side control block (StepCtl):
// the launches of a step depend only on which of the two ping-pong Q arrays is read, so they are
// captured once per parity into a CUDA graph and replayed (PYCS_GRAPH=0 turns that off).
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <map>
#include <vector>
#include "pycs_common.cuh"
#include "fused_args.cuh"
#include "mgpu.cuh"
#include "ghost_core.cuh"
#include "cube_edges.cuh"

namespace {

constexpr int WS_CAP = 4096;        // entries of the separable-wind time-factor table (power of two)
constexpr int WS_BATCH = 1024;      // entries filled per refill

// sum of n partials in a fixed order (every CTA gets the same bits)
__device__ double reduce_partials(const double* part, int n) {
  double v = 0.0;
  for (int k = 0; k < n; ++k) v += part[k];
  return v;
}

// projection coefficient of the step that ended last: -sum / a2, the expression of the step kernel.  Plain
// (cached, broadcast) loads: the kernels that call this run after the step that wrote the control block and
// nothing changes it while they run (volatile loads from every thread cost 300 us at N = 1536).
__device__ __forceinline__ double pending_corr(const StepCtl* ctl, const MgSync* sync, int world, double inv_a2) {
  if (!ctl->pend) return 0.0;
  const double sm = world > 1 ? reduce_partials(sync->psum[ctl->steps & 1], world) : ctl->sum;
  return -sm * inv_a2;
}

// ---- Lagrange ghost fill (src/interpolation.py:154-314) in ONE launch ---------------------------
//   * blocks [0, nb_sn): the S / N edge ghost cells at positions k in [k0, k1) -- the rows the caller
//     needs (all of them on one GPU, the slab +- 3 on several);
//   * blocks [nb_sn, nb_sn + nb_ew): the E / W edge ghost cells (sides selected by ew_mask);
//   * last 3 blocks: the 4 x 4 corners (src/interpolation.py:250-314).  A corner stencil reads the
//     neighbour's strip, whose ends are that neighbour's phase-1 ghosts; instead of waiting for them
//     they are recomputed in registers (same arithmetic, same bits), so corners depend on interior
//     cells only and the whole fill is one dependency-free kernel.
// fold != 0 (serial path): the pending MF-PR term is folded in -- the fill is linear, so
//   ghost(Q + corr*sqrtg) = ghost(Q) + corr * ghost(sqrtg), and gs = ghost(sqrtg) is static -- and the
//   coefficient is written to *corr_out for the step kernel.
// flags != null (several GPUs): every CTA first waits until each peer has delivered the exchange
//   ctl->xcount (bounded wait, mgpu.cuh).
struct GhostFillArgs {
  Geo g;
  HaloMaps maps;
  double* q;
  const int* kminE;
  const double* wE;
  int order;
  const double* gs;
  const StepCtl* ctl;
  int fold;
  double inv_a2;
  double* corr_out;
  const double* corr_in;     // != null: use this coefficient (ring restore after a run)
  const long long* flags;
  int world;
  int* mg_err;
  unsigned long long mg_timeout_ns;
  int k0, k1, nbx_sn, nbx_ew, ew_mask;
};

__global__ void dg_fill_kernel(const __grid_constant__ GhostFillArgs a) {
  const Geo& g = a.g;
  pdl_trigger();               // PDL (pycs_common.cuh): no-ops unless launched with the attribute
  pdl_wait();
  if (a.flags) {
    if ((int)threadIdx.x < a.world) {
      const long long xc = *((const volatile long long*)&a.ctl->xcount);
      mg_wait_flag(a.flags + threadIdx.x, xc, a.mg_err, a.mg_timeout_ns);     // acquire: see mgpu.cuh
    }
    __syncthreads();
  }
  double corr = 0.0;
  bool add = false;
  if (a.corr_in) {
    corr = *a.corr_in;
    add = true;
  } else if (a.fold) {
    corr = pending_corr(a.ctl, nullptr, 1, a.inv_a2);
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.corr_out = corr;
    add = true;
  }
  const int nb_sn = a.nbx_sn * 4 * 12, nb_ew = a.nbx_ew * 4 * 12;
  int blk = blockIdx.x;
  if (blk < nb_sn + nb_ew) {
    int bx, gl, p, s, k;
    if (blk < nb_sn) {
      bx = blk % a.nbx_sn;
      gl = (blk / a.nbx_sn) & 3;
      const int ps = blk / (4 * a.nbx_sn);           // 0 .. 11
      p = ps >> 1;
      s = SIDE_N + (ps & 1);
      k = a.k0 + bx * blockDim.x + threadIdx.x;
      if (k >= a.k1) return;
    } else {
      blk -= nb_sn;
      bx = blk % a.nbx_ew;
      gl = (blk / a.nbx_ew) & 3;
      const int ps = blk / (4 * a.nbx_ew);
      p = ps >> 1;
      s = SIDE_E + (ps & 1);
      if (!((a.ew_mask >> s) & 1)) return;
      k = g.lo + bx * blockDim.x + threadIdx.x;
      if (k >= g.hi) return;
    }
    double acc = dg_phase1_value(g, a.maps, a.q, a.kminE, a.wE, a.order, p, s, gl, k);
    int i, j;
    if (s == SIDE_E) { i = g.hi + gl; j = k; }
    else if (s == SIDE_W) { i = gl; j = k; }
    else if (s == SIDE_N) { i = k; j = g.hi + gl; }
    else { i = k; j = gl; }
    const long long id = gidx(g, p, i, j);
    if (add) acc = fma(a.gs[id], corr, acc);
    a.q[id] = acc;
    return;
  }
  // corners: 12 (panel, E|W) x 4 layers x 8 positions
  const int t = (blk - nb_sn - nb_ew) * blockDim.x + threadIdx.x;
  if (t >= 12 * 32) return;
  const int c32 = t & 31, pe = t >> 5;
  const int gl = c32 >> 3, c = c32 & 7;
  const int k = (c < 4) ? c : g.hi + (c - 4);
  const int p = pe >> 1, s = pe & 1;
  if (!((a.ew_mask >> s) & 1)) return;
  double acc = dg_corner_value(g, a.maps, a.q, a.kminE, a.wE, a.order, p, s, gl, k);
  const int i = (s == SIDE_E) ? g.hi + gl : gl;
  const long long id = gidx(g, p, i, k);
  if (add) acc = fma(a.gs[id], corr, acc);
  a.q[id] = acc;
}

// sqrtg of the single metric panel copied into the interior of all six panels of dst
__global__ void spread_metric_kernel(Geo g, const double* __restrict__ sgc, double* __restrict__ dst) {
  int j = g.lo + blockIdx.x * blockDim.x + threadIdx.x, i = g.lo + blockIdx.y, p = blockIdx.z;
  if (j >= g.hi) return;
  dst[gidx(g, p, i, j)] = sgc[gidx(g, 0, i, j)];
}

// add the pending projection term to the interior rows [i0, i1) (before anything else reads Q)
__global__ void flush_corr_kernel(Geo g, double* __restrict__ q, const double* __restrict__ sgc, const StepCtl* ctl,
                                  const MgSync* sync, int world, double inv_a2, int i0) {
  const double corr = pending_corr(ctl, sync, world, inv_a2);
  int j = g.lo + blockIdx.x * blockDim.x + threadIdx.x, i = i0 + blockIdx.y, p = blockIdx.z;
  if (j >= g.hi) return;
  long long id = gidx(g, p, i, j);
  q[id] = fma(sgc[gidx(g, 0, i, j)], corr, q[id]);
}

// ghost ring of the buffer the last step read (filled at the start of that step) -> the buffer it
// wrote, so that Q looks exactly like the reference's after the step; raw != 0: the ring was filled
// without the projection term of that step (split path), add it here
__global__ void copy_ring_kernel(Geo g, double* __restrict__ dst, const double* __restrict__ src, int raw,
                                 const double* __restrict__ gs, const StepCtl* ctl) {
  int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, p = blockIdx.z;
  if (j >= g.P) return;
  if (i >= g.lo && i < g.hi && j >= g.lo && j < g.hi) return;
  long long id = gidx(g, p, i, j);
  double v = src[id];
  if (raw) v = fma(gs[id], ctl->corr_applied, v);
  dst[id] = v;
}

// MF-AF (src/edges_treatment.py:231-278): the outer fluxes on the 12 cube edges are replaced by the average of
// the two panels' values (with the sign / flip of the edge pair) before the flux differences are taken.  The step
// kernel has already formed Q + (pxdF + pydF) / sqrtg with each panel's own edge flux and recorded those fluxes;
// the update is linear in them, so the averaging is a correction of the cells next to the cube edges:
//   side lo: Q[lo]     += (f_avg - f_own) / sqrtg,     side hi: Q[hi - 1] -= (f_avg - f_own) / sqrtg.
// pass = 0 corrects the ends that are x-edges, pass = 1 the y-edges: within a pass no two ends touch the same
// cell (a corner cell belongs to one x- and one y-edge), so the result does not depend on thread order.
__global__ void mf_af_patch_kernel(Geo g, const double* __restrict__ ef, const double* __restrict__ rgc,
                                   double* __restrict__ q, int pass) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= g.N) return;
  const CubeEdge ce = cube_edge(blockIdx.y);
  const int tb = ce.flip ? g.N - 1 - t : t;
  const double sg = ce.a.side == ce.b.side ? -1.0 : 1.0;
  auto rec = [&](const EdgeEnd& e, int pos) { return ef[((long long)e.panel * 4 + 2 * e.dir + e.side) * g.N + pos]; };
  const double fa = rec(ce.a, t), fb = rec(ce.b, tb);
  const double v = 0.5 * fa + sg * (0.5 * fb);           // :242-278, a = b = 1/2
  auto fix = [&](const EdgeEnd& e, int pos, double delta) {
    if (e.dir != pass) return;
    const int c = e.side == 0 ? g.lo : g.hi - 1;         // the cell next to the edge
    const int i = e.dir == 0 ? c : g.lo + pos, j = e.dir == 0 ? g.lo + pos : c;
    const long long id = gidx(g, e.panel, i, j);
    const double d = e.side == 0 ? delta : -delta;
    q[id] = fma(d, rgc[gidx(g, 0, i, j)], q[id]);
  };
  fix(ce.a, t, v - fa);
  fix(ce.b, tb, sg * v - fb);
}

__global__ void recip_kernel(Geo g, const double* __restrict__ s, double* __restrict__ d) {
  int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j > g.P) return;
  long long id = gidx(g, 0, i, j);
  double v = s[id];
  d[id] = (v != 0.0) ? 1.0 / v : 0.0;
}

// time factors of the separable wind for n consecutive steps: entry (s0 + i) & (WS_CAP-1) serves the
// step with control-block index s0 + i, which is reference step k0 + i and reads the wind of
// t = (k0 + i - 1) dt (src/advection_ic.py:301-305, src/advection_sphere.py:45-57)
__global__ void ws_fill_kernel(double* tab, long long s0, long long k0, int n, double dt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  tab[(s0 + i) & (WS_CAP - 1)] = cos(PYCS_PI * ((double)(k0 + i - 1) * dt) / PYCS_WIND_PERIOD);
}

__global__ void ctl_set_kernel(StepCtl* ctl, int what) {
  if (what == 0) {            // reset
    ctl->steps = 0;
    ctl->xcount = 0;
    ctl->pend = 0;
    ctl->corr_applied = 0.0;
    ctl->sum = 0.0;
  } else {                    // nothing pending any more
    ctl->pend = 0;
  }
}

}  // namespace

// fused-path state kept next to the handle (one per handle, keyed by pointer)
struct FusedState {
  double* rgc = nullptr;       // 1/sqrtg_pc
  double* gs = nullptr;        // ghost cells: Lagrange fill of the sqrtg field (static)
  double* part = nullptr;
  unsigned* counter = nullptr; // last-writer ticket of the step kernels
  StepCtl* ctl = nullptr;
  double* ws_tab = nullptr;    // [WS_CAP] + one constant entry for the timing launches
  long long steps_host = 0;    // host mirror of ctl->steps (every enqueued step completes in order)
  long long ws_off = 0;        // table filled for steps with k - index == ws_off ...
  long long ws_until = -1;     // ... and index < ws_until
  int prof = -1;               // PYCS_STEP_PROFILE: CUDA events around the kernels of every step
  std::vector<cudaEvent_t> ev; // 4 per profiled step
  std::vector<cudaEvent_t> ev2; // split step: 6 per profiled step (boundary stream: start, after boundary CTAs, after
                                // exchange kernel, after ghost fill; interior stream: start, end)
  int npart_cap = 0;
  double* bu = nullptr;        // separable wind: ucontra(t = 0) incl. ghost edges
  double* bv = nullptr;        //                 vcontra(t = 0)
  int base_valid = 0;
  // basis winds (wind mode 2): the contravariant wind incl. ghost edges of every basis field of the wind
  // (wind.cu), the per-step combination (instantaneous + departure-point averaged) and its coefficients
  double* basis_u[4] = {nullptr, nullptr, nullptr, nullptr};
  double* basis_v[4] = {nullptr, nullptr, nullptr, nullptr};
  int nbasis = 0, basis_valid = 0;
  double *wua = nullptr, *wum = nullptr, *wva = nullptr, *wvm = nullptr;
  double* wcoef = nullptr;     // [WS_CAP][8]
  long long wc_off = 0, wc_until = -1;
  int pending = 0;             // the sum of the last step waits to be applied (mirror of ctl->pend)
  int ring_pending = 0;        // ghost ring of the current buffer is stale
  int ring_raw = 0;            // ... and the ring of the buffer the last step read carries no projection term
  int ghost_ready = 0;         // split path: the current buffer's ghost cells are already (raw) filled
  int npart = 0;
  int tb = 160, rows = 0, nstrips = 0, wcols = 0, nchunks = 0;
  int impl = 0;                // 4: v2b (fused2b.cu), 2: v2 (fused.cu, limited reconstructions)
  // split step
  int split = 0;               // 0 undecided, 1 on, -1 off
  int4* cta_tab = nullptr;     // CTA table of the split step: n_b boundary CTAs, then n_i interior CTAs
  std::vector<CtaDesc> cta_host;
  int* xjob_off = nullptr;     // several GPUs: per boundary CTA, the pieces of its rows that peers read
  int4* xjobs = nullptr;
  unsigned* xcounter = nullptr;
  int xkernel = 0;             // PYCS_MG_XKERNEL=1: separate exchange kernel instead
  double* edge_flux = nullptr; // MF-AF: recorded outer fluxes on the panels' edge lines [6][4][N]
  int n_i = 0, n_b = 0;
  int band = 0, edge_rows = 0, irows = 0;
  cudaStream_t s2 = nullptr;
  cudaEvent_t e_fork = nullptr, e_join = nullptr;
  // graphs: one per ping-pong parity and wind mask
  int use_graph = -1;
  int use_pdl = -1;
  int onek = -1;               // one-kernel step (GH = 2: the step kernel fills its own ghost cells), PYCS_ONEKERNEL
  cudaGraphExec_t gexec[2][5] = {};
  const double* gq[2][5] = {};
  int gnodes[2][5] = {};
};

static std::map<pycs_handle, FusedState> g_fused;

StepCtl* k_fused_ctl(pycs_handle h) { return g_fused[h].ctl; }

static void drop_graphs(FusedState& fs) {
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 5; ++b) {
      if (fs.gexec[a][b]) cudaGraphExecDestroy(fs.gexec[a][b]);
      fs.gexec[a][b] = nullptr;
      fs.gq[a][b] = nullptr;
    }
}

int k_fused_supported(pycs_handle h) {
  // duo-grid ghost cells only (ET-S72/PL07 refill ghosts between the two stages and ET-PL07 couples
  // parabolas across panels).  MF-AF couples the outer fluxes across the cube edges: handled by recording
  // them in the step kernel and a correction pass (mf_af_patch_kernel) -- production kernel, one GPU.
  if (h->prm.et != 3) return 0;
  if (h->prm.mf == 2) return (pycs_fused2b_has(h->prm.recon, h->prm.opsplit) && !h->mg) ? 1 : 0;
  return 1;
}

// Makespan (in marched rows) of a split step with the given boundary / interior shapes, by list scheduling of
// its CTAs -- boundary CTAs first, as the high-priority stream dispatches them -- onto `slots` CTA slots.  A
// CTA costs its rows + 6 (the ramp of a chunk) + 6 (prologue, first staged rows, closing ticket), a boundary CTA
// 9 % more (ghost-cell projection term); the next ghost fill follows the last boundary CTA, ~20 rows' time.  Measured against
// it at N = 1536 on 2 GPUs: 108 rows predicted, 112 us per step observed (a marched row takes ~1 us when the
// SM is full).
static double split_shape_cost(int row_lo, int row_hi, int nstrips, int slots, int band, int edge_rows, int rows) {
  std::vector<CtaDesc> tab;
  const int nb = pycs_plan_split_ctas(row_lo, row_hi, nstrips, band, edge_rows, rows, &tab);
  std::vector<double> heap(slots, 0.0);            // min-heap of slot free times
  auto cmp = [](double a, double b) { return a > b; };
  double chain = 0.0, makespan = 0.0;
  for (int k = 0; k < (int)tab.size(); ++k) {
    std::pop_heap(heap.begin(), heap.end(), cmp);
    const double f = heap.back() + (tab[k].r1 - tab[k].r0 + 6) * (k < nb ? 1.09 : 1.0) + 6.0;   // + a CTA's fixed cost
    heap.back() = f;
    std::push_heap(heap.begin(), heap.end(), cmp);
    if (k < nb && f > chain) chain = f;
    if (f > makespan) makespan = f;
  }
  return makespan > chain + 20.0 ? makespan : chain + 20.0;
}

static int fused_setup(pycs_handle h, FusedState& fs) {
  const Geo& g = h->g;
  if (!fs.ctl) {
    CK(cudaMalloc(&fs.ctl, sizeof(StepCtl)));
    ctl_set_kernel<<<1, 1, 0, h->stream>>>(fs.ctl, 0);
    CKL(h);
    fs.steps_host = 0;
    CK(cudaMalloc(&fs.ws_tab, sizeof(double) * (WS_CAP + 1)));
    const double one = 0.999;                       // the timing launches' constant factor
    CK(cudaMemcpyAsync(fs.ws_tab + WS_CAP, &one, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  if (!fs.rgc) {
    double* sgc;
    TRY(pycs_field_ptr(h, PYCS_F_SQRTG_PC, &sgc));
    CK(cudaMalloc(&fs.rgc, sizeof(double) * g.ps));
    CK(cudaMemsetAsync(fs.rgc, 0, sizeof(double) * g.ps, h->stream));
    recip_kernel<<<dim3((g.P + 128) / 128, g.P + 1), 128, 0, h->stream>>>(g, sgc, fs.rgc);
    CKL(h);
  }
  if (fs.rows == 0) {
    const int nrows = h->row_hi - h->row_lo;       // rows this handle updates (multi-GPU: its slab)
    const char* er = getenv("PYCS_FUSED_ROWS");
    int rows = er ? atoi(er) : 0;
    int resident, lag;
    if (pycs_fused2b_has(h->prm.recon, h->prm.opsplit)) {
      fs.impl = 4;
      fs.tb = pycs_fused2b_threads();
      const int mask = (h->prm.dp == 2) ? 1 : 0;
      const int per_sm = pycs_fused2b_resident(h->prm.recon, h->prm.opsplit, mask);
      if (per_sm < 1) {
        pycs_set_error("fused2b kernel: occupancy query failed");
        return PYCS_ERR_CUDA;
      }
      resident = h->sm_count * per_sm;
      lag = 6;
    } else {                                       // limited reconstructions: v2
      if (h->mg) {
        pycs_set_error("multi-GPU handles need PPM-0 / PPM-PL07 (the limited reconstructions run on the "
                       "first-generation kernel, single GPU only)");
        return PYCS_ERR_ARG;
      }
      fs.impl = 2;
      fs.tb = pycs_fused_v2_threads();
      resident = h->sm_count * pycs_fused_v2_resident();
      lag = 4;
    }
    // strips of TB-6 columns, one column per thread; chunks sized so that the grid fills whole waves
    const int wmax = fs.tb - 6;
    fs.nstrips = (g.N + wmax - 1) / wmax;
    fs.wcols = (g.N + fs.nstrips - 1) / fs.nstrips;
    fs.wcols += fs.wcols & 1;                      // even: staged rows start 16-byte aligned
    const int cols = 6 * fs.nstrips;
    if (rows <= 0) {
      // whole waves of resident CTAs: time ~ waves * (rows + ramp)
      int best = 0;
      double best_cost = 1e30;
      for (int nch = 1; nch <= nrows; ++nch) {
        int rr = (nrows + nch - 1) / nch;
        if (rr < 8 && nch > 1) break;
        int nb = cols * ((nrows + rr - 1) / rr);
        int waves = (nb + resident - 1) / resident;
        double cost = (double)waves * (rr + lag);
        if (cost < best_cost) { best_cost = cost; best = rr; }
      }
      rows = best;
    }
    fs.rows = rows;
    fs.nchunks = (nrows + rows - 1) / rows;
    drop_graphs(fs);
  }
  if (fs.onek < 0) {
    const char* eo = getenv("PYCS_ONEKERNEL");
    // Default: on for a handle sharded over two GPUs, off otherwise -- measured at N = 1536 (profiles/r2_mgpu_onekernel.md):
    //   1 GPU   0.185 ms per step against 0.170 for the serial step: the CTAs at a panel edge pay ~8 us for their own
    //           fill and the late first row copies, and with two such CTAs in a slot the launch ends 15 us later;
    //   2 GPUs  0.108 against 0.113 for the two-stream split step (one launch and no events instead of three launches);
    //   8 GPUs  0.055 against 0.049: with 22-row chunks the fill of the edge CTAs is a quarter of their march, and
    //           the split step hides it behind the interior CTAs.
    const int dflt = (h->mg && h->mg->world == 2) ? 1 : 0;
    fs.onek = ((eo ? atoi(eo) : dflt) && fs.impl == 4 && h->prm.mf != 2) ? 1 : 0;   // (MF-AF: serial step)
  }
  if (fs.split == 0) {
    const char* es = getenv("PYCS_SPLIT");
    const bool want = h->mg || (es && atoi(es));   // several GPUs always run the split step
    fs.split = (want && fs.impl == 4 && h->prm.mf != 2) ? 1 : -1;   // (MF-AF runs the serial step)
    if (fs.split == 1) {
      // Boundary CTAs march few rows each, so that they are done early and the exchange + ghost fill of
      // the next step run beside the interior CTAs: bands of `band` rows at both ends of the slab, the
      // first / last strip in chunks of `edge_rows`, the interior in chunks of `irows`.  The three sizes
      // are chosen by simulating the launch (split_shape_cost); PYCS_SPLIT_BAND / _EDGE_ROWS / _ROWS override.
      const char* eb = getenv("PYCS_SPLIT_BAND");
      const char* ee = getenv("PYCS_SPLIT_EDGE_ROWS");
      const char* ei = getenv("PYCS_SPLIT_ROWS");
      const int nrows = h->row_hi - h->row_lo;
      const int per_sm = pycs_fused2b_resident(h->prm.recon, h->prm.opsplit, (h->prm.dp == 2) ? 1 : 0);
      const int slots = h->sm_count * (per_sm > 0 ? per_sm : 4);
      int band = 12, edge = 32, irows = nrows;
      double best = 1e30;
      const int bands[] = {8, 12, 16, 24}, edges[] = {12, 16, 20, 24, 28, 32, 40, 48, 64, 96};
      for (int b : bands)
        for (int e : edges) {
          const int inner = nrows - 2 * b;
          const int nmax = inner > 16 ? (inner / 8 < 64 ? inner / 8 : 64) : 1;
          for (int n = 1; n <= nmax; ++n) {
            const int ir = inner > 0 ? (inner + n - 1) / n : nrows;
            const double c = split_shape_cost(h->row_lo, h->row_hi, fs.nstrips, slots, b, e, ir);
            if (c < best) { best = c; band = b; edge = e; irows = ir; }
          }
        }
      if (fs.onek) {
        // one-kernel step: nothing has to finish early -- the ghost fill is part of the next launch, the exchange
        // only has to be out by the end of the step -- so the table is the uniform grid of the serial step (one
        // wave of equal chunks).  Measured at 2 GPUs, N = 1536: 0.1081 ms per step against 0.1142 with the short
        // boundary CTAs of the two-stream shape (profiles/r2_mgpu_onekernel.md).
        band = edge = irows = fs.rows;
      }
      if (eb) band = atoi(eb);
      if (ee) edge = atoi(ee);
      if (ei && atoi(ei) > 0) irows = atoi(ei);
      if (band < 4) band = 4;                      // >= the 4-wide E / W strips and the 3 halo rows
      fs.band = band;
      fs.edge_rows = edge;
      fs.irows = irows;
      if (getenv("PYCS_STEP_PROFILE"))
        fprintf(stderr, "[pycs split shape] rows [%d,%d): band %d, edge chunks %d, interior chunks %d rows; predicted %.0f rows' time\n",
                h->row_lo, h->row_hi, band, edge, irows,
                split_shape_cost(h->row_lo, h->row_hi, fs.nstrips, slots, band, edge, irows));
      std::vector<CtaDesc> tab;
      fs.n_b = pycs_plan_split_ctas(h->row_lo, h->row_hi, fs.nstrips, fs.band, fs.edge_rows, irows, &tab);
      fs.n_i = (int)tab.size() - fs.n_b;
      static_assert(sizeof(CtaDesc) == sizeof(int4), "CTA table entries are int4");
      CK(cudaMalloc(&fs.cta_tab, sizeof(int4) * tab.size()));
      CK(cudaMemcpy(fs.cta_tab, tab.data(), sizeof(int4) * tab.size(), cudaMemcpyHostToDevice));
      fs.cta_host = tab;
      if (const char* ex = getenv("PYCS_MG_XKERNEL")) fs.xkernel = atoi(ex);
      if (h->mg && h->mg->connected && !fs.xkernel) {
        // in-kernel exchange: cut the send rectangles of the plan along the boundary CTAs that write them
        std::vector<int> off(fs.n_b + 1, 0);
        std::vector<int4> jobs;
        for (int c = 0; c < fs.n_b; ++c) {
          const CtaDesc& d = tab[c];
          const int cj0 = g.lo + d.strip * fs.wcols, cj1 = cj0 + fs.wcols < g.hi ? cj0 + fs.wcols : g.hi;
          for (const MgRect& r : *h->mg->rects) {
            if (r.panel != d.panel) continue;
            const int i0 = r.i0 > d.r0 ? r.i0 : d.r0, i1 = r.i1 < d.r1 ? r.i1 : d.r1;
            const int j0 = r.j0 > cj0 ? r.j0 : cj0, j1 = r.j1 < cj1 ? r.j1 : cj1;
            if (i0 < i1 && j0 < j1) jobs.push_back(make_int4(r.peer | (i0 << 4), i1, j0, j1));
          }
          off[c + 1] = (int)jobs.size();
        }
        // every cell of the plan must be shipped by some boundary CTA (tests/test_host_cpu.py checks the cover)
        long long planned = 0, cut = 0;
        for (const MgRect& r : *h->mg->rects) planned += (long long)(r.i1 - r.i0) * (r.j1 - r.j0);
        for (const int4& x : jobs) cut += (long long)(x.y - (x.x >> 4)) * (x.w - x.z);
        if (planned != cut) {
          pycs_set_error("split step: the boundary CTAs do not cover the exchange plan");
          return PYCS_ERR_STATE;
        }
        if (jobs.empty()) jobs.push_back(make_int4(0, 0, 0, 0));
        CK(cudaMalloc(&fs.xjob_off, sizeof(int) * off.size()));
        CK(cudaMemcpy(fs.xjob_off, off.data(), sizeof(int) * off.size(), cudaMemcpyHostToDevice));
        CK(cudaMalloc(&fs.xjobs, sizeof(int4) * jobs.size()));
        CK(cudaMemcpy(fs.xjobs, jobs.data(), sizeof(int4) * jobs.size(), cudaMemcpyHostToDevice));
        if (!fs.xcounter) {
          CK(cudaMalloc(&fs.xcounter, sizeof(unsigned)));
          CK(cudaMemset(fs.xcounter, 0, sizeof(unsigned)));
        }
      }
      if (!fs.s2) {
        int lo_pri = 0, hi_pri = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
        CK(cudaStreamCreateWithPriority(&fs.s2, cudaStreamNonBlocking, lo_pri));   // interior CTAs: behind the handle's stream
        CK(cudaEventCreateWithFlags(&fs.e_fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&fs.e_join, cudaEventDisableTiming));
      }
    }
  }
  // MF-PR partial sums: one per CTA
  const int nb = (fs.split == 1) ? fs.n_b + fs.n_i : 6 * fs.nstrips * fs.nchunks;
  if (fs.npart_cap < nb) {
    if (fs.part) cudaFree(fs.part);
    CK(cudaMalloc(&fs.part, sizeof(double) * nb));
    fs.npart_cap = nb;
  }
  fs.npart = nb;
  if (h->prm.mf == 2 && !fs.edge_flux) {
    CK(cudaMalloc(&fs.edge_flux, sizeof(double) * 24 * g.N));
    CK(cudaMemsetAsync(fs.edge_flux, 0, sizeof(double) * 24 * g.N, h->stream));
  }
  if (!fs.counter) {
    CK(cudaMalloc(&fs.counter, sizeof(unsigned)));
    CK(cudaMemsetAsync(fs.counter, 0, sizeof(unsigned), h->stream));
  }
  if (fs.use_graph < 0) {
    // replay from a CUDA graph: on by default for the serial and the one-kernel step; the two-stream split step is
    // launched directly unless PYCS_GRAPH=1 (measured at 2 GPUs, N = 1536: 0.131 ms per step replayed, 0.122 ms launched)
    const char* eg = getenv("PYCS_GRAPH");
    fs.use_graph = eg ? (atoi(eg) ? 1 : 0) : ((fs.split == 1 && !fs.onek) ? 0 : 1);
  }
  if (fs.use_pdl < 0) {
    const char* ep = getenv("PYCS_PDL");
    fs.use_pdl = ep ? atoi(ep) : 1;
  }
  if (fs.prof < 0) fs.prof = getenv("PYCS_STEP_PROFILE") ? 1 : 0;
  return 0;
}

// ---- ghost fill launcher ---------------------------------------------------------------------------
// whole = fill every ghost cell of the sphere (single GPU, and the stand-alone fills); otherwise the
// cells of rows [mg->gf_lo, mg->gf_hi) only -- what this rank's step reads, and what the exchange
// plan delivers the sources of.
static int launch_ghost_fill(pycs_handle h, double* q, cudaStream_t st, const double* gs, const StepCtl* ctl, int fold,
                             double* corr_out, const double* corr_in, bool wait_peers, bool pdl = false) {
  const Geo& g = h->g;
  GhostFillArgs a;
  a.g = g;
  a.maps = h->maps;
  a.q = q;
  a.kminE = h->kminE;
  a.wE = h->wE;
  a.order = h->order;
  a.gs = gs;
  a.ctl = ctl;
  a.fold = fold;
  a.inv_a2 = (fold && h->a2_valid) ? 1.0 / h->a2 : 0.0;
  a.corr_out = corr_out;
  a.corr_in = corr_in;
  a.flags = nullptr;
  a.world = 0;
  a.mg_err = nullptr;
  a.mg_timeout_ns = 0;
  a.k0 = g.lo;
  a.k1 = g.hi;
  a.ew_mask = 3;
  if (h->mg) {
    a.k0 = h->mg->gf_lo < g.lo ? g.lo : h->mg->gf_lo;
    a.k1 = h->mg->gf_hi > g.hi ? g.hi : h->mg->gf_hi;
    a.ew_mask = (h->mg->gf_hi > g.hi ? 1 : 0) | (h->mg->gf_lo < g.lo ? 2 : 0);     // bit SIDE_E = 0, SIDE_W = 1
    if (wait_peers) {
      a.flags = h->mg->sync->dflag;
      a.world = h->mg->world;
      a.mg_err = &h->mg->sync->err;
      a.mg_timeout_ns = h->mg->timeout_ns;
    }
  }
  a.nbx_sn = (a.k1 - a.k0 + 127) / 128;
  a.nbx_ew = a.ew_mask ? (g.N + 127) / 128 : 0;
  const int nblk = a.nbx_sn * 4 * 12 + a.nbx_ew * 4 * 12 + 3;
  if (pdl) CK(pycs_launch_pdl(dg_fill_kernel, dim3(nblk), dim3(128), 0, st, true, a));
  else dg_fill_kernel<<<nblk, 128, 0, st>>>(a);
  CKL(h);
  return 0;
}

// The plain Lagrange ghost fill (src/interpolation.py:154-314) of any centre field in one launch:
// same arithmetic as dg_phase1_kernel + dg_phase2_kernel of halo.cu, without the dependency
// between the two phases (corners recompute the neighbour's edge ghosts they read).
int k_dg_fill_single(pycs_handle h, double* q) {
  if (!h->kminE) {
    pycs_set_error("ET-DG ghost fill needs pycs_upload_lagrange first");
    return PYCS_ERR_STATE;
  }
  MgpuState* mg = h->mg;       // a stand-alone fill is a whole-sphere operation
  h->mg = nullptr;
  const int r = launch_ghost_fill(h, q, h->stream, nullptr, nullptr, 0, nullptr, nullptr, false);
  h->mg = mg;
  return r;
}

// PYCS_STEP_PROFILE: per-kernel device time of the run that just ended
void k_fused_profile_report(pycs_handle h) {
  auto it = g_fused.find(h);
  if (it == g_fused.end()) return;
  FusedState& fs = it->second;
  if (fs.prof <= 0 || fs.ev.size() < 8) return;
  cudaStreamSynchronize(h->stream);
  const size_t ns = fs.ev.size() / 4, skip = ns > 8 ? 4 : 0;
  double t[4] = {0, 0, 0, 0};
  for (size_t k = skip; k < ns; ++k) {
    float ms;
    for (int j = 0; j < 3; ++j) {
      cudaEventElapsedTime(&ms, fs.ev[4 * k + j], fs.ev[4 * k + j + 1]);
      t[j] += ms;
    }
    if (k + 1 < ns) {
      cudaEventElapsedTime(&ms, fs.ev[4 * k + 3], fs.ev[4 * k + 4]);
      t[3] += ms;
    }
  }
  const double n = (double)(ns - skip);
  fprintf(stderr, "[pycs step profile] rank %d: %zu steps; prologue (ghost fill / winds) %.2f us, step kernel(s) %.2f us, "
                  "join %.2f us, gap to next step %.2f us\n",
          h->mg ? h->mg->rank : 0, ns - skip, 1e3 * t[0] / n, 1e3 * t[1] / n, 1e3 * t[2] / n, 1e3 * t[3] / n);
  for (auto e : fs.ev) cudaEventDestroy(e);
  fs.ev.clear();
  if (fs.ev2.size() >= 12) {
    const size_t n2 = fs.ev2.size() / 6, sk = n2 > 8 ? 4 : 0;
    double u[4] = {0, 0, 0, 0};
    for (size_t k = sk; k < n2; ++k) {
      float ms;
      cudaEventElapsedTime(&ms, fs.ev2[6 * k], fs.ev2[6 * k + 1]); u[0] += ms;
      cudaEventElapsedTime(&ms, fs.ev2[6 * k + 1], fs.ev2[6 * k + 2]); u[1] += ms;
      cudaEventElapsedTime(&ms, fs.ev2[6 * k + 2], fs.ev2[6 * k + 3]); u[2] += ms;
      cudaEventElapsedTime(&ms, fs.ev2[6 * k + 3], fs.ev2[6 * k + 4]); u[3] += ms;
    }
    const double m = (double)(n2 - sk);
    fprintf(stderr, "[pycs split profile] rank %d: boundary CTAs %.2f us, exchange kernel %.2f us, next ghost fill (+ wait "
                    "for the peers' data) %.2f us, then %.2f us until the interior CTAs are done\n",
            h->mg ? h->mg->rank : 0, 1e3 * u[0] / m, 1e3 * u[1] / m, 1e3 * u[2] / m, 1e3 * u[3] / m);
  }
  for (auto e : fs.ev2) cudaEventDestroy(e);
  fs.ev2.clear();
}

// A new Q was uploaded into PYCS_F_Q: whatever the fused path had pending belonged to the old state.
int k_fused_discard(pycs_handle h) {
  if (h->qcur == 1) {
    double* t = h->f[PYCS_F_Q];
    h->f[PYCS_F_Q] = h->f[PYCS_F_Q_NEXT];
    h->f[PYCS_F_Q_NEXT] = t;
    h->qcur = 0;
  }
  auto it = g_fused.find(h);
  if (it == g_fused.end()) return 0;
  FusedState& fs = it->second;
  if (fs.pending && fs.ctl) {
    ctl_set_kernel<<<1, 1, 0, h->stream>>>(fs.ctl, 1);
    CKL(h);
  }
  fs.pending = 0;
  fs.ring_pending = 0;
  fs.ghost_ready = 0;
  return 0;
}

// Several GPUs, after every rank uploaded its own rows of a new state into PYCS_F_Q: one exchange of the
// current buffer, so that each rank holds the peers' cells its first step reads (the regular exchange
// only ever ships a step's output).
int k_fused_share_slab(pycs_handle h) {
  if (!h->mg) return 0;
  FusedState& fs = g_fused[h];
  TRY(fused_setup(h, fs));
  double* q;
  TRY(pycs_field_ptr(h, h->qcur ? PYCS_F_Q_NEXT : PYCS_F_Q, &q));
  // nobody may still be reading or rewriting this buffer: every rank has finished the steps issued so far
  // and the flush that followed them
  TRY(k_mg_wait_steps(h, h->stream));
  TRY(k_mg_quiesce_wait(h, h->stream));
  TRY(k_mg_exchange(h, q, fs.ctl, h->stream));
  fs.ghost_ready = 0;
  return 0;
}

// apply the pending projection term to the current Q so that every other code
// path (download, operator kernels, diagnostics) sees the reference's Q
int k_fused_flush(pycs_handle h) {
  auto it = g_fused.find(h);
  if (it == g_fused.end()) return 0;
  FusedState& fs = it->second;
  if (!fs.pending && !fs.ring_pending) return 0;
  const bool had_pending = fs.pending != 0;
  const Geo& g = h->g;
  double *sgc, *q, *qo;
  TRY(pycs_field_ptr(h, PYCS_F_SQRTG_PC, &sgc));
  TRY(pycs_field_ptr(h, h->qcur ? PYCS_F_Q_NEXT : PYCS_F_Q, &q));
  TRY(pycs_field_ptr(h, h->qcur ? PYCS_F_Q : PYCS_F_Q_NEXT, &qo));
  if (fs.pending) {
    const MgSync* sync = nullptr;
    int world = 1;
    if (h->mg) {                     // per-rank sums of the last step, once they have all arrived
      TRY(k_mg_wait_steps(h, h->stream));
      sync = h->mg->sync;
      world = h->mg->world;
    }
    // the whole interior, also on a sharded handle: the peers' cells it holds (halo rows, ghost-fill
    // sources) must carry the term as well, the next run reads them without a pending coefficient
    flush_corr_kernel<<<dim3((g.N + 127) / 128, g.N, 6), 128, 0, h->stream>>>(g, q, sgc, fs.ctl, sync, world,
                                                                             1.0 / h->a2, g.lo);
    CKL(h);
    ctl_set_kernel<<<1, 1, 0, h->stream>>>(fs.ctl, 1);
    CKL(h);
    fs.pending = 0;
  }
  if (fs.ring_pending) {
    copy_ring_kernel<<<dim3((g.P + 127) / 128, g.P, 6), 128, 0, h->stream>>>(g, q, qo, fs.ring_raw && fs.gs ? 1 : 0,
                                                                             fs.gs, fs.ctl);
    CKL(h);
    fs.ring_pending = 0;
  }
  fs.ghost_ready = 0;
  if (h->mg && had_pending) TRY(k_mg_quiesce_raise(h, h->stream));
  return 0;
}

void k_fused_release(pycs_handle h) {
  auto it = g_fused.find(h);
  if (it == g_fused.end()) return;
  FusedState& fs = it->second;
  drop_graphs(fs);
  if (fs.rgc) cudaFree(fs.rgc);
  if (fs.gs) cudaFree(fs.gs);
  if (fs.part) cudaFree(fs.part);
  if (fs.counter) cudaFree(fs.counter);
  if (fs.ctl) cudaFree(fs.ctl);
  if (fs.ws_tab) cudaFree(fs.ws_tab);
  if (fs.bu) cudaFree(fs.bu);
  if (fs.bv) cudaFree(fs.bv);
  for (int m = 0; m < 4; ++m) {
    if (fs.basis_u[m]) cudaFree(fs.basis_u[m]);
    if (fs.basis_v[m]) cudaFree(fs.basis_v[m]);
  }
  for (double* p : {fs.wua, fs.wum, fs.wva, fs.wvm, fs.wcoef})
    if (p) cudaFree(p);
  if (fs.cta_tab) cudaFree(fs.cta_tab);
  if (fs.xjob_off) cudaFree(fs.xjob_off);
  if (fs.xjobs) cudaFree(fs.xjobs);
  if (fs.xcounter) cudaFree(fs.xcounter);
  if (fs.edge_flux) cudaFree(fs.edge_flux);
  if (fs.e_fork) cudaEventDestroy(fs.e_fork);
  if (fs.e_join) cudaEventDestroy(fs.e_join);
  if (fs.s2) cudaStreamDestroy(fs.s2);
  for (auto e : fs.ev) cudaEventDestroy(e);
  g_fused.erase(it);
}

// the rows this handle updates changed (pycs_mgpu_init): recompute the launch geometry and restart
// the step count (it is the epoch of the peers' flags, which start at zero)
void k_fused_reset_grid(pycs_handle h) {
  FusedState& fs = g_fused[h];
  fs.rows = 0;
  if (fs.cta_tab) cudaFree(fs.cta_tab);  // the CTA table of the split step belongs to the old grid
  fs.cta_tab = nullptr;
  if (fs.xjob_off) cudaFree(fs.xjob_off);
  if (fs.xjobs) cudaFree(fs.xjobs);
  fs.xjob_off = nullptr;
  fs.xjobs = nullptr;
  fs.n_i = fs.n_b = 0;
  fs.split = 0;                          // decided again by fused_setup
  fs.ghost_ready = 0;
  drop_graphs(fs);
  if (fs.ctl) {
    ctl_set_kernel<<<1, 1, 0, h->stream>>>(fs.ctl, 0);
    h->launches++;
    fs.steps_host = 0;
    fs.ws_until = -1;
    fs.wc_until = -1;
  }
}

// the exchange plan changed (new Lagrange tables on a sharded handle): the split step's tables follow it
void k_fused_replan_exchange(pycs_handle h) {
  auto it = g_fused.find(h);
  if (it == g_fused.end()) return;
  FusedState& fs = it->second;
  if (fs.cta_tab) cudaFree(fs.cta_tab);
  if (fs.xjob_off) cudaFree(fs.xjob_off);
  if (fs.xjobs) cudaFree(fs.xjobs);
  fs.cta_tab = nullptr;
  fs.xjob_off = nullptr;
  fs.xjobs = nullptr;
  fs.n_i = fs.n_b = 0;
  fs.split = 0;
  drop_graphs(fs);
}

// geometry was re-uploaded: 1/sqrtg and the t = 0 winds must be rebuilt
void k_fused_invalidate(pycs_handle h) {
  auto it = g_fused.find(h);
  if (it == g_fused.end()) return;
  FusedState& fs = it->second;
  if (fs.rgc) cudaFree(fs.rgc);
  fs.rgc = nullptr;
  if (fs.gs) cudaFree(fs.gs);
  fs.gs = nullptr;
  fs.base_valid = 0;
  fs.basis_valid = 0;
  drop_graphs(fs);
}

void k_fused_invalidate_ghost_metric(pycs_handle h) {
  auto it = g_fused.find(h);
  if (it == g_fused.end()) return;
  if (it->second.gs) cudaFree(it->second.gs);
  it->second.gs = nullptr;
  it->second.ghost_ready = 0;
  drop_graphs(it->second);
}

// kernel parameters that are baked into captured launches changed (dt)
void k_fused_invalidate_graphs(pycs_handle h) {
  auto it = g_fused.find(h);
  if (it == g_fused.end()) return;
  drop_graphs(it->second);
  it->second.ws_until = -1;
  it->second.wc_until = -1;
}

// Separable wind (vf = 3, RK1): the step kernel scales the contravariant wind of t = 0
// (interior + ghost edges, exactly what init_vars_adv leaves in ucontra_averaged,
// src/advection_vars.py:37-87) by cos(pi t / T).  The copy is private to the fused path:
// ucontra_averaged itself is overwritten by every non-separable step.  Building it
// overwrites U_pu / U_pv / U_pc; the lazy catch-up (wind_sync in capi.cu) restores them.
static int ensure_base_winds(pycs_handle h, FusedState& fs) {
  if (fs.base_valid) return 0;
  const size_t bytes = sizeof(double) * 6 * (size_t)h->g.ps;
  if (!fs.bu) CK(cudaMalloc(&fs.bu, bytes));
  if (!fs.bv) CK(cudaMalloc(&fs.bv, bytes));
  TRY(k_wind_interior(h, 0.0, 1, 1));
  TRY(k_wind_ghost_fill(h));
  double *u, *v;
  TRY(pycs_field_ptr(h, PYCS_F_PU_UCONTRA, &u));
  TRY(pycs_field_ptr(h, PYCS_F_PV_VCONTRA, &v));
  CK(cudaMemcpyAsync(fs.bu, u, bytes, cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaMemcpyAsync(fs.bv, v, bytes, cudaMemcpyDeviceToDevice, h->stream));
  fs.base_valid = 1;
  drop_graphs(fs);
  return 0;
}

// Basis winds (wind mode 2: field 2, or field 3 with RK2): every basis field of the wind pushed once
// through the wind pipeline of init_vars_adv (wind.cu: k_wind_basis_build), kept as contravariant
// fields incl. ghost edges.  Building them overwrites U_pu / U_pv / U_pc; the lazy catch-up restores them.
static int ensure_basis_winds(pycs_handle h, FusedState& fs) {
  if (fs.basis_valid) return 0;
  const size_t bytes = sizeof(double) * 6 * (size_t)h->g.ps;
  fs.nbasis = k_wind_basis_count(h);
  double *u, *v;
  TRY(pycs_field_ptr(h, PYCS_F_PU_UCONTRA, &u));
  TRY(pycs_field_ptr(h, PYCS_F_PV_VCONTRA, &v));
  for (int m = 0; m < fs.nbasis; ++m) {
    if (!fs.basis_u[m]) CK(cudaMalloc(&fs.basis_u[m], bytes));
    if (!fs.basis_v[m]) CK(cudaMalloc(&fs.basis_v[m], bytes));
    TRY(k_wind_basis_build(h, m));
    CK(cudaMemcpyAsync(fs.basis_u[m], u, bytes, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(fs.basis_v[m], v, bytes, cudaMemcpyDeviceToDevice, h->stream));
  }
  for (double** p : {&fs.wua, &fs.wum, &fs.wva, &fs.wvm})
    if (!*p) {
      CK(cudaMalloc(p, bytes));
      CK(cudaMemsetAsync(*p, 0, bytes, h->stream));
    }
  if (!fs.wcoef) CK(cudaMalloc(&fs.wcoef, sizeof(double) * 8 * WS_CAP));
  fs.basis_valid = 1;
  fs.wc_until = -1;
  drop_graphs(fs);
  return 0;
}

// Everything the reference's step k leaves in U_pu / U_pv / U_pc, rebuilt from the analytic wind
// after one or more fused steps that did not touch those arrays: the wind of t_{k-1} with -- for
// RK2 -- the wind of t_{k-2} in ucontra_old (src/advection_timestep.py:61-62), the ghost fill and the
// departure velocity of step k (:31-37), then update_adv(t_k) (:48-75).
int k_wind_catch_up(pycs_handle h, long long k) {
  if (k < 1 || h->prm.vf < 2) return 0;
  const double dt = h->g.dt;
  if (h->prm.dp == 2 && k >= 2) {
    TRY(k_wind_interior(h, (double)(k - 2) * dt, 1, 1));
    TRY(k_wind_ghost_fill(h));
    TRY(k_update_adv(h, (double)(k - 1) * dt));       // old <- wind(t_{k-2}) incl. its ghost edges; wind(t_{k-1})
  } else {
    TRY(k_wind_interior(h, (double)(k - 1) * dt, 1, 1));
    if (h->prm.dp == 2) {                              // first step: old = the wind of t = 0 after its ghost fill
      TRY(k_wind_ghost_fill(h));
      double *u, *v, *uo, *vo;
      TRY(pycs_field_ptr(h, PYCS_F_PU_UCONTRA, &u));
      TRY(pycs_field_ptr(h, PYCS_F_PV_VCONTRA, &v));
      TRY(pycs_field_ptr(h, PYCS_F_PU_UOLD, &uo));
      TRY(pycs_field_ptr(h, PYCS_F_PV_VOLD, &vo));
      const size_t bytes = sizeof(double) * 6 * (size_t)h->g.ps;
      CK(cudaMemcpyAsync(uo, u, bytes, cudaMemcpyDeviceToDevice, h->stream));
      CK(cudaMemcpyAsync(vo, v, bytes, cudaMemcpyDeviceToDevice, h->stream));
    }
  }
  TRY(k_wind_ghost_fill(h));
  TRY(k_time_averaged_velocity(h));
  return k_update_adv(h, (double)k * dt);
}

// ghost(sqrtg): the Lagrange fill applied to the metric field itself, once (see dg_fill_kernel)
static int ensure_gs(pycs_handle h, FusedState& fs) {
  if (h->prm.mf != 3 || fs.gs) return 0;
  if (!h->kminE) {
    pycs_set_error("fused step needs pycs_upload_lagrange first");
    return PYCS_ERR_STATE;
  }
  const Geo& g = h->g;
  double* sgc;
  TRY(pycs_field_ptr(h, PYCS_F_SQRTG_PC, &sgc));
  const size_t bytes = sizeof(double) * 6 * (size_t)g.ps;
  CK(cudaMalloc(&fs.gs, bytes));
  CK(cudaMemsetAsync(fs.gs, 0, bytes, h->stream));
  spread_metric_kernel<<<dim3((g.N + 127) / 128, g.N, 6), 128, 0, h->stream>>>(g, sgc, fs.gs);
  CKL(h);
  TRY(k_dg_fill_single(h, fs.gs));
  drop_graphs(fs);
  return 0;
}

// arguments common to every launch of the step kernel
static int step_args(pycs_handle h, FusedState& fs, const double* qcur, double* qnext, int mask, FusedArgs* out,
                     int wind_mode = 0) {
  const Geo& g = h->g;
  double *sgc, *sgu, *sgv, *ua, *va, *um, *vm;
  TRY(pycs_field_ptr(h, PYCS_F_SQRTG_PC, &sgc));
  TRY(pycs_field_ptr(h, PYCS_F_SQRTG_PU, &sgu));
  TRY(pycs_field_ptr(h, PYCS_F_SQRTG_PV, &sgv));
  TRY(pycs_field_ptr(h, PYCS_F_PU_UAVG, &ua));
  TRY(pycs_field_ptr(h, PYCS_F_PV_VAVG, &va));
  TRY(pycs_field_ptr(h, PYCS_F_PU_UCONTRA, &um));
  TRY(pycs_field_ptr(h, PYCS_F_PV_VCONTRA, &vm));
  FusedArgs a;
  memset(&a, 0, sizeof a);
  a.g = g;
  a.q = qcur; a.qn = qnext;
  if (mask == 2) { ua = fs.bu; va = fs.bv; }
  if (wind_mode == 2) { ua = fs.wua; va = fs.wva; um = fs.wum; vm = fs.wvm; }
  a.ua = ua; a.va = va; a.um = um; a.vm = vm;
  a.sgc = sgc; a.rgc = fs.rgc; a.sgu = sgu; a.sgv = sgv;
  a.part = fs.part;
  a.counter = fs.counter;
  a.ctl = fs.ctl;
  a.ws_tab = fs.ws_tab;
  a.ws_mask = WS_CAP - 1;
  a.apply_corr = (h->prm.mf == 3) ? 1 : 0;
  a.corr_ptr = nullptr;
  a.inv_a2 = (h->prm.mf == 3) ? 1.0 / h->a2 : 0.0;
  a.gs = fs.gs;
  a.rows_per_chunk = fs.rows; a.nstrips = fs.nstrips; a.wcols = fs.wcols;
  a.row_lo = h->row_lo; a.row_hi = h->row_hi;
  a.cdx = g.dt / g.dx; a.cdy = g.dt / g.dy;
  a.cta_tab = nullptr;
  a.cta_off = 0;
  a.nblk_total = fs.npart;
  a.pub.world = 0;
  a.timing = 0;
  *out = a;
  return 0;
}

static int launch_step(pycs_handle h, FusedState& fs, const FusedArgs& a, int mask, int gh, int nblocks, cudaStream_t st) {
  if (fs.impl == 4) CK(pycs_launch_fused2b(a, h->prm.recon, h->prm.opsplit, mask, gh, nblocks, st));
  else CK(pycs_launch_fused_v2(a, h->prm.recon, h->prm.opsplit, mask, nblocks, st));
  CKL(h);
  return 0;
}

// Device time of `reps` back-to-back launches of the step kernel alone over the whole grid (ping-pong
// buffers, no ghost fill): the roofline measurement of bench.py.  Leaves Q undefined.
int k_fused_time_kernel(pycs_handle h, int reps, int separable, float* ms) {
  FusedState& fs = g_fused[h];
  TRY(fused_setup(h, fs));
  if (h->prm.mf == 3 && !h->a2_valid) {
    TRY(k_sum_sq_metric(h, &h->a2));
    h->a2_valid = 1;
  }
  double *qa, *qb;
  TRY(pycs_field_ptr(h, PYCS_F_Q, &qa));
  TRY(pycs_field_ptr(h, PYCS_F_Q_NEXT, &qb));
  const int mask = separable ? 2 : ((h->prm.dp == 2) ? 1 : 0);
  if (separable) TRY(ensure_base_winds(h, fs));
  TRY(ensure_gs(h, fs));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaEventRecord(h->ev0, h->stream));
  for (int r = 0; r < reps; ++r) {
    FusedArgs a;
    TRY(step_args(h, fs, (r & 1) ? qb : qa, (r & 1) ? qa : qb, mask, &a));
    a.corr_ptr = h->red_out + 8;          // some small coefficient: the arithmetic is what is timed
    a.ws_tab = fs.ws_tab + WS_CAP;
    a.ws_mask = 0;
    a.timing = 1;
    a.nblk_total = 6 * fs.nstrips * fs.nchunks;
    TRY(launch_step(h, fs, a, mask, 0, 6 * fs.nstrips * fs.nchunks, h->stream));
  }
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaEventSynchronize(h->ev1));
  CK(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return k_fused_discard(h);
}

int k_fused_kernel_name(pycs_handle h, char* out, int len) {
  FusedState& fs = g_fused[h];
  TRY(fused_setup(h, fs));
  const int mask = (h->prm.vf == 3 && h->prm.dp == 1 && !h->no_separable) ? 2 : ((h->prm.dp == 2) ? 1 : 0);
  if (fs.impl == 4)
    snprintf(out, len, "fused2b_kernel<TB=%d,recon=%d,split=%d,mask=%d> (csrc/fused2b.cu)", fs.tb, h->prm.recon,
             h->prm.opsplit, mask);
  else
    snprintf(out, len, "fused_step_kernel<TB=%d,recon=%d,split=%d,D=6> (csrc/fused.cu)", fs.tb, h->prm.recon,
             h->prm.opsplit);
  return 0;
}

int k_fused_grid_info(pycs_handle h, int* tb, int* rows, int* nblocks) {
  FusedState& fs = g_fused[h];
  TRY(fused_setup(h, fs));
  *tb = fs.tb;
  *rows = fs.rows;
  *nblocks = 6 * fs.nstrips * fs.nchunks;
  return 0;
}

// ---- one step, enqueued (or captured) on the handle's stream ----------------------------------------
// serial: ghost fill with the projection term folded in, (wind kernels), step kernel over the whole grid
static int step_winds(pycs_handle h, FusedState& fs, bool winds, int wind_mode, bool pdl = false) {
  if (winds) {                    // the reference's wind kernels (src/advection_timestep.py:31-37)
    TRY(k_wind_ghost_fill(h));
    TRY(k_time_averaged_velocity(h));
  } else if (wind_mode == 2) {    // the same winds combined from the basis fields, one launch
    TRY(k_wind_basis_combine(h, fs.basis_u, fs.basis_v, fs.nbasis, fs.wua, fs.wum, fs.wva, fs.wvm, fs.wcoef,
                             WS_CAP - 1, &fs.ctl->steps, pdl ? 1 : 0));
  }
  return 0;
}

static int enqueue_serial(pycs_handle h, FusedState& fs, double* qcur, double* qnext, int mask, bool winds,
                          bool profile, int wind_mode) {
  auto mark = [&]() {
    if (!profile || fs.ev.size() >= 4 * 4096) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, h->stream);
    fs.ev.push_back(e);
  };
  mark();
  // The kernels of a serial step on one GPU are chained by programmatic dependent launches (pycs_common.cuh):
  // each may be scheduled while its predecessor drains and waits on the device for it to complete.
  // Measured (profiles/r2_pdl.log, N=1536): 0.1738 -> 0.1725 ms per step replayed from the graph, N=48: 18.4 -> 17.4 us,
  // N=384: 30.7 -> 28.7 us; launched directly the chain got slower (0.1749 -> 0.1770), and so did the basis-wind
  // step (two more kernels in the chain: 0.1538 -> 0.1568) -- PDL is used where it paid.
  const bool pdl = fs.use_pdl && fs.use_graph && !h->mg && !profile && !winds && wind_mode != 2;
  // 1. ghost cells of Q (src/advection_timestep.py:28), folding in the pending MF-PR term
  TRY(launch_ghost_fill(h, qcur, h->stream, fs.gs, fs.ctl, h->prm.mf == 3 ? 1 : 0, h->red_out + 8, nullptr, false, pdl));
  // 2. winds (src/advection_timestep.py:31-37)
  TRY(step_winds(h, fs, winds, wind_mode, pdl));
  mark();
  // 3. divergence + Q update
  FusedArgs a;
  TRY(step_args(h, fs, qcur, qnext, mask, &a, wind_mode));
  a.corr_ptr = (h->prm.mf == 3) ? h->red_out + 8 : nullptr;
  a.pdl = (pdl && fs.impl == 4) ? 1 : 0;
  if (h->prm.mf == 2) {          // MF-AF: record the outer edge fluxes (GH = 1 flavour), then average them
    a.edge_flux = fs.edge_flux;
    TRY(launch_step(h, fs, a, mask, 1, fs.npart, h->stream));
    const Geo& g = h->g;
    for (int pass = 0; pass < 2; ++pass) {
      mf_af_patch_kernel<<<dim3((g.N + 127) / 128, 12), 128, 0, h->stream>>>(g, fs.edge_flux, fs.rgc, qnext, pass);
      CKL(h);
    }
  } else {
    TRY(launch_step(h, fs, a, mask, 0, fs.npart, h->stream));
  }
  mark();
  mark();
  return 0;
}

// split: boundary CTAs (+ exchange) -> ghost fill of the next step on the handle's high-priority stream, interior
// CTAs on the second stream.  Precondition: the ghost cells of qcur are (raw) filled.
static int enqueue_split(pycs_handle h, FusedState& fs, double* qcur, double* qnext, int mask, bool winds,
                         bool profile, int wind_mode) {
  auto mark = [&]() {
    if (!profile || fs.ev.size() >= 4 * 4096) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, h->stream);
    fs.ev.push_back(e);
  };
  mark();
  TRY(step_winds(h, fs, winds, wind_mode));
  mark();
  FusedArgs a;
  TRY(step_args(h, fs, qcur, qnext, mask, &a, wind_mode));
  if (h->mg) {
    a.wait_flags = h->mg->sync->sflag;
    a.wait_world = h->mg->world;
    a.mg_err = &h->mg->sync->err;
    a.mg_timeout_ns = h->mg->timeout_ns;
    a.pub.world = h->mg->world;
    a.pub.rank = h->mg->rank;
    for (int d = 0; d < 8; ++d) a.pub.peer_sync[d] = d < h->mg->world ? h->mg->peer_sync[d] : nullptr;
  }
  // The boundary CTAs go on the handle's stream (high priority, no cross-stream dependency to resolve before
  // their launch): they take the CTA slots first.  The interior launch follows on the second stream.
  // (The other way round, the interior launch won the slots and the boundary CTAs trickled in behind it:
  // 92 us for the boundary launch at 4 GPUs instead of ~25, profiles/r2_mgpu_split_profile.md.)
  CK(cudaEventRecord(fs.e_fork, h->stream));            // everything before this step
  auto mark2 = [&](cudaStream_t st) {
    if (!profile || fs.ev2.size() >= 6 * 4096) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    fs.ev2.push_back(e);
  };
  mark2(h->stream);
  FusedArgs b = a;
  b.cta_tab = fs.cta_tab;
  b.cta_off = 0;
  const bool xin = h->mg && fs.xjob_off;                // the boundary CTAs ship their rows themselves
  if (xin) {
    b.xjob_off = fs.xjob_off;
    b.xjobs = fs.xjobs;
    b.xcounter = fs.xcounter;
    b.n_boundary = fs.n_b;
    const int idx = (qnext == h->mg->alloc[0]) ? 0 : 1;
    for (int d = 0; d < 8; ++d) b.xpeer_q[d] = d < h->mg->world ? h->mg->peer_q[idx][d] : nullptr;
  }
  TRY(launch_step(h, fs, b, mask, 1, fs.n_b, h->stream));   // reads ghost cells, feeds the peers
  mark2(h->stream);
  if (fs.n_i) {
    CK(cudaStreamWaitEvent(fs.s2, fs.e_fork, 0));
    FusedArgs c = a;
    c.cta_tab = fs.cta_tab;
    c.cta_off = fs.n_b;
    TRY(launch_step(h, fs, c, mask, 0, fs.n_i, fs.s2));     // reads no ghost cell
    CK(cudaEventRecord(fs.e_join, fs.s2));
  }
  if (h->mg && !xin) TRY(k_mg_exchange(h, qnext, fs.ctl, h->stream));
  mark2(h->stream);
  // the ghost cells the NEXT step reads: their sources are boundary cells of this step's output (own:
  // stream order; the peers': dflag), so the fill runs beside this step's interior CTAs
  TRY(launch_ghost_fill(h, qnext, h->stream, nullptr, fs.ctl, 0, nullptr, nullptr, true));
  mark2(h->stream);
  if (fs.n_i) CK(cudaStreamWaitEvent(h->stream, fs.e_join, 0));
  mark2(h->stream);
  mark2(h->stream);
  mark();
  mark();
  return 0;
}

// one kernel: every CTA of the step in ONE launch of the GH = 2 flavour -- the CTAs that stage ghost cells fill them
// first (the Lagrange fill inside the step kernel, raw; on several GPUs after waiting for the peers' exchange flags),
// the boundary CTAs of a sharded handle come first in the table and ship their rows to the peers as soon as they
// are written.  No ghost-fill kernel, no second stream, no events: the step is a single graph node.
static int enqueue_onekernel(pycs_handle h, FusedState& fs, double* qcur, double* qnext, int mask, int wind_mode) {
  TRY(step_winds(h, fs, false, wind_mode));
  FusedArgs a;
  TRY(step_args(h, fs, qcur, qnext, mask, &a, wind_mode));
  a.gf_maps = h->maps;
  a.gf_kmin = h->kminE;
  a.gf_w = h->wE;
  a.gf_order = h->order;
  if (fs.split == 1) {
    a.cta_tab = fs.cta_tab;
    a.cta_off = 0;
    a.n_boundary = fs.n_b;
  }
  if (h->mg) {
    a.wait_flags = h->mg->sync->sflag;
    a.wait_world = h->mg->world;
    a.mg_err = &h->mg->sync->err;
    a.mg_timeout_ns = h->mg->timeout_ns;
    a.pub.world = h->mg->world;
    a.pub.rank = h->mg->rank;
    for (int d = 0; d < 8; ++d) a.pub.peer_sync[d] = d < h->mg->world ? h->mg->peer_sync[d] : nullptr;
    a.gf_flags = h->mg->sync->dflag;
    if (!fs.xjob_off) {
      pycs_set_error("one-kernel step on several GPUs needs the in-kernel exchange (PYCS_MG_XKERNEL must be 0)");
      return PYCS_ERR_STATE;
    }
    a.xjob_off = fs.xjob_off;
    a.xjobs = fs.xjobs;
    a.xcounter = fs.xcounter;
    const int idx = (qnext == h->mg->alloc[0]) ? 0 : 1;
    for (int d = 0; d < 8; ++d) a.xpeer_q[d] = d < h->mg->world ? h->mg->peer_q[idx][d] : nullptr;
  }
  // launched directly (PYCS_GRAPH=0), the step kernels can chain by programmatic dependent launch (PYCS_PDL=2): the
  // next step's CTAs become resident and set up their shared memory while this step's last CTAs drain
  a.pdl = (fs.use_pdl == 2 && !fs.use_graph && wind_mode != 2) ? 1 : 0;
  TRY(launch_step(h, fs, a, mask, 2, fs.npart, h->stream));
  return 0;
}

// wind_mode: how the step gets its winds without the reference's per-step wind kernels
//   0  none of the below: steady winds (fields 1 and 4) are read as init_vars_adv left them, time-dependent
//      ones are refreshed by the wind kernels around the step kernel (PYCS_NO_SEPARABLE);
//   1  separable: wind field 3 with RK1 is U(0) cos(pi t / T) exactly (src/advection_ic.py:301-305): the step
//      kernel reads a private copy of the t = 0 winds and scales them;
//   2  basis winds: the instantaneous and the departure-point averaged wind are combined from the static
//      basis fields by one kernel (wind.cu: wind_basis_kernel).
// In modes 1 and 2 U_pu / U_pv / U_pc are not touched; the caller marks them stale (capi.cu: wind_sync).
int k_fused_step(pycs_handle h, long long k, double t, int wind_mode) {
  FusedState& fs = g_fused[h];
  TRY(fused_setup(h, fs));
  if (h->prm.mf == 3 && !h->a2_valid) {
    TRY(k_sum_sq_metric(h, &h->a2));
    h->a2_valid = 1;
  }
  if (!h->kminE) {
    pycs_set_error("fused step needs pycs_upload_lagrange first");
    return PYCS_ERR_STATE;
  }
  if (h->mg && !h->mg->connected) {
    pycs_set_error("multi-GPU step before pycs_mgpu_connect");
    return PYCS_ERR_STATE;
  }
  double *qa, *qb;
  TRY(pycs_field_ptr(h, PYCS_F_Q, &qa));
  TRY(pycs_field_ptr(h, PYCS_F_Q_NEXT, &qb));
  double* qcur = h->qcur ? qb : qa;
  double* qnext = h->qcur ? qa : qb;
  const bool separable = wind_mode == 1;
  if (separable) TRY(ensure_base_winds(h, fs));
  if (wind_mode == 2) TRY(ensure_basis_winds(h, fs));
  TRY(ensure_gs(h, fs));
  const int mask = separable ? 2 : ((h->prm.dp == 2) ? 1 : 0);    // RK1: averaged wind == instantaneous wind
  const bool split = fs.split == 1;
  const bool winds = (h->prm.vf == 2 || h->prm.vf == 3) && wind_mode == 0;
  const bool profile = fs.prof > 0;
  const bool onek = fs.onek == 1 && !winds && !profile && (!h->mg || (split && !fs.xkernel));

  // per-step coefficients of the lazy wind modes for this and the following steps
  const long long idx = fs.steps_host;
  if (separable && !(k - idx == fs.ws_off && idx < fs.ws_until)) {
    ws_fill_kernel<<<(WS_BATCH + 255) / 256, 256, 0, h->stream>>>(fs.ws_tab, idx, k, WS_BATCH, h->g.dt);
    CKL(h);
    fs.ws_off = k - idx;
    fs.ws_until = idx + WS_BATCH;
  }
  if (wind_mode == 2 && !(k - idx == fs.wc_off && idx < fs.wc_until)) {
    TRY(k_wind_coef_fill(h, fs.wcoef, WS_CAP - 1, idx, k, WS_BATCH));
    fs.wc_off = k - idx;
    fs.wc_until = idx + WS_BATCH;
  }
  if (split && !onek && !fs.ghost_ready) {
    // first step after something else touched Q: the raw ghost fill this step's boundary CTAs read
    TRY(launch_ghost_fill(h, qcur, h->stream, nullptr, fs.ctl, 0, nullptr, nullptr, true));
  }

  const bool graph = fs.use_graph && !winds && !profile;
  if (graph) {
    // nothing may allocate or configure inside the capture: touch every field and kernel attribute first
    FusedArgs warm;
    TRY(step_args(h, fs, qcur, qnext, mask, &warm, wind_mode));
    if (fs.impl == 4 && (pycs_fused2b_resident(h->prm.recon, h->prm.opsplit, mask, onek ? 2 : 0) < 1 ||
                         pycs_fused2b_resident(h->prm.recon, h->prm.opsplit, mask, (split || h->prm.mf == 2) ? 1 : 0) < 1)) {
      pycs_set_error("fused2b kernel: occupancy query failed");
      return PYCS_ERR_CUDA;
    }
    const int par = (qcur == qa) ? 0 : 1;      // keyed by the buffer that is read (f[Q] / f[Q_NEXT] may have been swapped)
    const int key = wind_mode == 2 ? 3 + mask : mask;      // (a handle runs either the one-kernel step or the other shapes)
    cudaGraphExec_t& ge = fs.gexec[par][key];
    if (ge && fs.gq[par][key] != qcur) {
      cudaGraphExecDestroy(ge);
      ge = nullptr;
    }
    if (!ge) {
      const long long l0 = h->launches;
      cudaGraph_t gr = nullptr;
      CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
      int r = onek ? enqueue_onekernel(h, fs, qcur, qnext, mask, wind_mode)
              : split ? enqueue_split(h, fs, qcur, qnext, mask, false, false, wind_mode)
                      : enqueue_serial(h, fs, qcur, qnext, mask, false, false, wind_mode);
      cudaError_t ce = cudaStreamEndCapture(h->stream, &gr);
      if (r) return r;
      CK(ce);
      CK(cudaGraphInstantiate(&ge, gr, 0));
      cudaGraphDestroy(gr);
      fs.gq[par][key] = qcur;
      fs.gnodes[par][key] = (int)(h->launches - l0);
      h->launches = l0;                        // captured, not run
    }
    CK(cudaGraphLaunch(ge, h->stream));
    h->launches += fs.gnodes[par][key];
  } else if (onek) {
    TRY(enqueue_onekernel(h, fs, qcur, qnext, mask, wind_mode));
  } else if (split) {
    TRY(enqueue_split(h, fs, qcur, qnext, mask, winds, profile, wind_mode));
  } else {
    TRY(enqueue_serial(h, fs, qcur, qnext, mask, winds, profile, wind_mode));
  }
  fs.steps_host += 1;
  h->last_step_kernel_launches++;
  h->qcur ^= 1;
  fs.pending = (h->prm.mf == 3) ? 1 : 0;
  fs.ring_pending = 1;
  fs.ring_raw = (split || onek) ? 1 : 0;
  fs.ghost_ready = (split && !onek) ? 1 : 0;
  // wind refresh for the next step (src/advection_timestep.py:48-75)
  if (winds) TRY(k_update_adv(h, t));
  return 0;
}
