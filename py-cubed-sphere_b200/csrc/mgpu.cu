// Multi-GPU layer of the fused step: row slabs of the six panels, peer-mapped halo stores, epoch flags.
//
// The reference is single process ("new design").  Rank k of W owns the rows
// [row_lo, row_hi) of all six panels -- 6 W tiles of (N / W) x N cells, six per GPU.  A step of rank k
// reads, besides its own rows,
//   * the 3 rows above and below the slab (the 7 x 7 dependence box of the split scheme), and
//   * the interior cells its ghost cells are interpolated from (src/interpolation.py:154-314): pieces
//     of the 4-wide boundary strips of the neighbouring panels, which panel rotation scatters over the
//     slabs of other ranks (src/halo_data.py:66-182: a transposed edge maps my rows to the first / last
//     rows of the neighbour = rank 0 or W-1, a flipped edge to the mirror rank);
//   * one scalar per rank for the MF-PR projection (src/discrete_operators.py:98-101).
// The plan (pycs_mgpu_plan_rects) derives, from the halo index maps and the Lagrange stencil table, the
// exact set of cells each peer needs from this rank and packs it into rectangles; nothing is broadcast.
//
// Per step (stepper.cu): the boundary CTAs of the step kernel run first on a high-priority stream, the
// exchange kernel below stores the rectangles straight into the peers' Q arrays at their natural
// positions (CUDA IPC mappings over NVLink: no staging, no NCCL call on the data path) and raises
// `dflag` on every peer, and the ghost fill of the next step follows at once -- all beside the interior
// CTAs.  The last CTA of the step kernel publishes the rank's MF-PR sum and raises `sflag`; the next
// step's kernels wait for the W sflags themselves.  Everything is stream ordered: no host
// synchronisation in the step loop.  Flags are epoch counters, so the two ping-pong Q arrays double as
// the double buffer of the exchange (a peer can be at most one step ahead).
#include <cstdlib>
#include <cstring>
#include "pycs_common.cuh"
#include "mgpu.cuh"
#include "fused_args.cuh"

namespace {

__global__ void mg_wait_steps_kernel(MgSync* sync, const StepCtl* ctl, int world, unsigned long long timeout_ns) {
  const int d = threadIdx.x;
  if (d >= world) return;
  const long long steps = *((const volatile long long*)&ctl->steps);
  mg_wait_flag(&sync->sflag[d], steps, &sync->err, timeout_ns);
  __threadfence_system();
}

struct PeerSync { MgSync* s[MG_MAX_WORLD]; };

__global__ void mg_barrier_kernel(PeerSync peers, MgSync* sync, int world, int rank, long long value,
                                  unsigned long long timeout_ns) {
  if ((int)threadIdx.x >= world) return;
  __threadfence_system();
  *((volatile long long*)&peers.s[threadIdx.x]->bflag[rank]) = value;
  mg_wait_flag(&sync->bflag[threadIdx.x], value, &sync->err, timeout_ns);
  __threadfence_system();
}
__global__ void mg_quiesce_raise_kernel(PeerSync peers, int world, int rank, long long value) {
  if ((int)threadIdx.x >= world) return;
  __threadfence_system();
  *((volatile long long*)&peers.s[threadIdx.x]->qflag[rank]) = value;
}
__global__ void mg_quiesce_wait_kernel(MgSync* sync, int world, long long value, unsigned long long timeout_ns) {
  const int d = threadIdx.x;
  if (d >= world) return;
  mg_wait_flag(&sync->qflag[d], value, &sync->err, timeout_ns);
  __threadfence_system();
}

// One launch per step, one CTA per rectangle: (1) copy the rectangle of one panel of src into the same
// position of a peer's array; (2) system-scope fence, ticket; (3) the last CTA raises dflag on every
// rank.  Latency, not bandwidth, is what matters here (a few hundred KB per rank and step).
__global__ void mg_exchange_kernel(Geo g, const MgScatterJob* __restrict__ jobs, const double* __restrict__ src,
                                   PeerSync peers, int world, int rank, StepCtl* ctl, unsigned* __restrict__ counter) {
  __shared__ int last;
  const long long xc = *((const volatile long long*)&ctl->xcount);
  {
    const MgScatterJob jb = jobs[blockIdx.x];
    const int w = jb.j1 - jb.j0, n = (jb.i1 - jb.i0) * w;
    const int p = jb.panel;
    if (((w | jb.j0) & 1) == 0) {             // 16-byte aligned rectangle: two cells per store
      const int w2 = w >> 1, n2 = n >> 1;
      for (int t = threadIdx.x; t < n2; t += blockDim.x) {
        const int i = jb.i0 + t / w2, j = jb.j0 + 2 * (t % w2);
        const long long id = gidx(g, p, i, j);
        *reinterpret_cast<double2*>(jb.dst + id) = *reinterpret_cast<const double2*>(src + id);
      }
    } else {
      for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int i = jb.i0 + t / w, j = jb.j0 + t % w;
        const long long id = gidx(g, p, i, j);
        jb.dst[id] = src[id];
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // one system-scope fence per CTA: it is cumulative over the stores the barrier ordered before it
    __threadfence_system();
    last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  if (threadIdx.x == 0) {
    *counter = 0;
    ctl->xcount = xc + 1;
  }
  if (threadIdx.x < world) {
    __threadfence_system();
    *((volatile long long*)&peers.s[threadIdx.x]->dflag[rank]) = xc + 1;
  }
}

}  // namespace

// ---- host-side plan (no GPU needed: exercised by the CPU tests) ----------------------------
void pycs_mgpu_rows(int N, int world, int rank, int* row_lo, int* row_hi) {
  const int base = N / world, rem = N % world;
  const int off = rank * base + (rank < rem ? rank : rem);
  *row_lo = PYCS_NG + off;
  *row_hi = *row_lo + base + (rank < rem ? 1 : 0);
}

namespace {
// the index logic of ghost_core.cuh on the host: mark the interior cells a ghost cell is interpolated from
struct Planner {
  Geo g;
  HaloMaps maps;
  const int* kmin;
  int order;
  std::vector<unsigned char>* need;    // [6][P][P]
  void mark(int p, int i, int j) { (*need)[((size_t)p * g.P + i) * g.P + j] = 1; }
  void phase1(int p, int s, int gl, int k) {
    const SideMap& m = maps.m[p][s];
    const int ge = (s == SIDE_E || s == SIDE_N) ? gl : PYCS_NG - 1 - gl;
    const int km = kmin[ge * g.P + k];
    for (int l = 0; l < order; ++l) {
      const int A = (s < 2) ? gl : km + l, B = (s < 2) ? km + l : gl;
      mark(m.nb, m.ci + m.ai * A + m.bi * B, m.cj + m.aj * A + m.bj * B);
    }
  }
  void corner(int p, int s, int gl, int k) {
    const SideMap& m = maps.m[p][s];
    const int ge = (s == SIDE_E) ? gl : PYCS_NG - 1 - gl;
    const int km = kmin[ge * g.P + k];
    for (int l = 0; l < order; ++l) {
      const int b_ = km + l;
      const int si = m.ci + m.ai * gl + m.bi * b_, sj = m.cj + m.aj * gl + m.bj * b_;
      const bool ii = si >= g.lo && si < g.hi, jj = sj >= g.lo && sj < g.hi;
      if (ii && jj) mark(m.nb, si, sj);
      else if (ii) phase1(m.nb, sj >= g.hi ? SIDE_N : SIDE_S, sj >= g.hi ? sj - g.hi : sj, si);
      else phase1(m.nb, si >= g.hi ? SIDE_E : SIDE_W, si >= g.hi ? si - g.hi : si, sj);
    }
  }
  void ghost(int p, int i, int j) {
    const bool ii = i >= g.lo && i < g.hi, jj = j >= g.lo && j < g.hi;
    if (ii) phase1(p, j >= g.hi ? SIDE_N : SIDE_S, j >= g.hi ? j - g.hi : j, i);
    else if (jj) phase1(p, i >= g.hi ? SIDE_E : SIDE_W, i >= g.hi ? i - g.hi : i, j);
    else corner(p, i >= g.hi ? SIDE_E : SIDE_W, i >= g.hi ? i - g.hi : i, j);
  }
};

Geo plan_geo(int N) {
  Geo g;
  memset(&g, 0, sizeof g);
  g.N = N;
  g.P = N + 2 * PYCS_NG;
  g.lo = PYCS_NG;
  g.hi = PYCS_NG + N;
  g.ld = ((PYCS_JOFF + g.P + 1 + 15) / 16) * 16;
  g.ps = (long long)(g.P + 1) * g.ld;
  return g;
}
}  // namespace

// Interior cells rank `reader` reads in one step: its rows +- 3, and the sources of every ghost cell
// of the rows [row_lo - 3, row_hi + 3) (all four ghost layers: the ghost-fill kernel fills them all).
static void need_mask(const Geo& g, const HaloMaps& maps, const int* kmin, int order, int world, int reader,
                      std::vector<unsigned char>* need) {
  need->assign((size_t)6 * g.P * g.P, 0);
  Planner pl{g, maps, kmin, order, need};
  int a, b;
  pycs_mgpu_rows(g.N, world, reader, &a, &b);
  const int ia = a - 3 < g.lo ? 0 : a - 3, ib = b + 3 > g.hi ? g.P : b + 3;   // with the W / E ghost rows at the ends
  for (int p = 0; p < 6; ++p)
    for (int i = ia; i < ib; ++i)
      for (int j = 0; j < g.P; ++j) {
        const bool inner = i >= g.lo && i < g.hi && j >= g.lo && j < g.hi;
        if (inner) pl.mark(p, i, j);
        else pl.ghost(p, i, j);
      }
}

void pycs_mgpu_plan_rects(int N, int world, int rank, const int* kmin_east, int order, std::vector<MgRect>* out) {
  out->clear();
  const Geo g = plan_geo(N);
  HaloMaps maps;
  pycs_build_halo_maps(g, &maps);
  int a, b;
  pycs_mgpu_rows(N, world, rank, &a, &b);
  std::vector<unsigned char> need;
  for (int d = 0; d < world; ++d) {
    if (d == rank) continue;
    need_mask(g, maps, kmin_east, order, world, d, &need);
    // rectangles of my rows: column runs per row, merged over consecutive rows with identical runs
    for (int p = 0; p < 6; ++p) {
      std::vector<MgRect> open;          // rectangles still growing at row i - 1
      for (int i = a; i <= b; ++i) {
        std::vector<MgRect> runs;
        if (i < b) {
          const unsigned char* row = &need[((size_t)p * g.P + i) * g.P];
          for (int j = g.lo; j < g.hi;) {
            if (!row[j]) { ++j; continue; }
            int j1 = j;
            while (j1 < g.hi && row[j1]) ++j1;
            runs.push_back(MgRect{d, p, i, i + 1, j, j1});
            j = j1;
          }
        }
        std::vector<MgRect> next;
        for (auto& r : runs) {
          bool merged = false;
          for (auto& o : open)
            if (o.i1 == i && o.j0 == r.j0 && o.j1 == r.j1) {
              o.i1 = i + 1;
              next.push_back(o);
              o.peer = -1;               // consumed
              merged = true;
              break;
            }
          if (!merged) next.push_back(r);
        }
        for (auto& o : open)
          if (o.peer >= 0) out->push_back(o);
        open.swap(next);
      }
      for (auto& o : open) out->push_back(o);
    }
  }
}

// ---- device-side state ---------------------------------------------------------------------
int k_mg_init(pycs_handle h, int rank, int world, unsigned char* handles_out) {
  if (world < 2 || world > MG_MAX_WORLD || rank < 0 || rank >= world) {
    pycs_set_error("pycs_mgpu_init: world must be 2..8 and 0 <= rank < world");
    return PYCS_ERR_ARG;
  }
  if (h->g.N / world < PYCS_NG) {
    pycs_set_error("pycs_mgpu_init: fewer than 4 rows per rank");
    return PYCS_ERR_ARG;
  }
  if (h->mg) {
    pycs_set_error("pycs_mgpu_init: already initialised");
    return PYCS_ERR_STATE;
  }
  if (!h->kmin_host) {
    pycs_set_error("pycs_mgpu_init: upload the Lagrange tables first (pycs_upload_lagrange): the exchange plan is derived from them");
    return PYCS_ERR_STATE;
  }
  MgpuState* mg = new MgpuState();
  memset(mg, 0, sizeof *mg);
  mg->rank = rank;
  mg->world = world;
  mg->rects = new std::vector<MgRect>();
  const char* et = getenv("PYCS_MG_TIMEOUT_S");
  const double tsec = et ? atof(et) : 30.0;
  mg->timeout_ns = (unsigned long long)((tsec > 0.001 ? tsec : 30.0) * 1e9);
  TRY(pycs_field_ptr(h, PYCS_F_Q, &mg->alloc[0]));
  TRY(pycs_field_ptr(h, PYCS_F_Q_NEXT, &mg->alloc[1]));
  CK(cudaMalloc(&mg->sync, sizeof(MgSync)));
  CK(cudaMemset(mg->sync, 0, sizeof(MgSync)));
  CK(cudaMalloc(&mg->counter, sizeof(unsigned)));
  CK(cudaMemset(mg->counter, 0, sizeof(unsigned)));
  CK(cudaStreamSynchronize(h->stream));
  cudaIpcMemHandle_t hd[3];
  CK(cudaIpcGetMemHandle(&hd[0], mg->alloc[0]));
  CK(cudaIpcGetMemHandle(&hd[1], mg->alloc[1]));
  CK(cudaIpcGetMemHandle(&hd[2], mg->sync));
  memcpy(handles_out, hd, sizeof hd);
  pycs_mgpu_rows(h->g.N, world, rank, &h->row_lo, &h->row_hi);
  mg->gf_lo = h->row_lo - 3;
  mg->gf_hi = h->row_hi + 3;
  h->mg = mg;
  k_fused_reset_grid(h);
  return 0;
}

int k_mg_replan(pycs_handle h) {
  MgpuState* mg = h->mg;
  if (!mg || !mg->connected) return 0;
  pycs_mgpu_plan_rects(h->g.N, mg->world, mg->rank, h->kmin_host, h->order, mg->rects);
  // one CTA per rectangle: cut the wide ones (halo rows) so that no CTA copies more than ~12 KB
  {
    std::vector<MgRect> cut;
    const int wmax = 512;
    for (const MgRect& r : *mg->rects)
      for (int j = r.j0; j < r.j1; j += wmax) cut.push_back(MgRect{r.peer, r.panel, r.i0, r.i1, j, j + wmax < r.j1 ? j + wmax : r.j1});
    mg->rects->swap(cut);
  }
  const int n = (int)mg->rects->size();
  std::vector<MgScatterJob> jobs(n > 0 ? n : 1);
  for (int idx = 0; idx < 2; ++idx) {
    for (int k = 0; k < n; ++k) {
      const MgRect& r = (*mg->rects)[k];
      jobs[k] = MgScatterJob{mg->peer_q[idx][r.peer], r.panel, r.i0, r.i1, r.j0, r.j1};
    }
    if (mg->jobs_dev[idx]) cudaFree(mg->jobs_dev[idx]);
    mg->jobs_dev[idx] = nullptr;
    CK(cudaMalloc(&mg->jobs_dev[idx], sizeof(MgScatterJob) * jobs.size()));
    CK(cudaMemcpy(mg->jobs_dev[idx], jobs.data(), sizeof(MgScatterJob) * jobs.size(), cudaMemcpyHostToDevice));
  }
  mg->njobs = n;
  k_fused_replan_exchange(h);
  return 0;
}

int k_mg_connect(pycs_handle h, const unsigned char* all_handles) {
  MgpuState* mg = h->mg;
  if (!mg) {
    pycs_set_error("pycs_mgpu_connect before pycs_mgpu_init");
    return PYCS_ERR_STATE;
  }
  for (int d = 0; d < mg->world; ++d) {
    if (d == mg->rank) {
      mg->peer_q[0][d] = mg->alloc[0];
      mg->peer_q[1][d] = mg->alloc[1];
      mg->peer_sync[d] = mg->sync;
      continue;
    }
    cudaIpcMemHandle_t hd[3];
    memcpy(hd, all_handles + (size_t)d * sizeof hd, sizeof hd);
    void* p[3];
    for (int k = 0; k < 3; ++k) CK(cudaIpcOpenMemHandle(&p[k], hd[k], cudaIpcMemLazyEnablePeerAccess));
    mg->peer_q[0][d] = (double*)p[0];
    mg->peer_q[1][d] = (double*)p[1];
    mg->peer_sync[d] = (MgSync*)p[2];
  }
  mg->connected = 1;
  return k_mg_replan(h);
}

void k_mg_release(pycs_handle h) {
  MgpuState* mg = h->mg;
  if (!mg) return;
  for (int d = 0; d < mg->world; ++d) {
    if (d == mg->rank || !mg->connected) continue;
    cudaIpcCloseMemHandle(mg->peer_q[0][d]);
    cudaIpcCloseMemHandle(mg->peer_q[1][d]);
    cudaIpcCloseMemHandle(mg->peer_sync[d]);
  }
  if (mg->sync) cudaFree(mg->sync);
  if (mg->counter) cudaFree(mg->counter);
  for (int idx = 0; idx < 2; ++idx)
    if (mg->jobs_dev[idx]) cudaFree(mg->jobs_dev[idx]);
  delete mg->rects;
  delete mg;
  h->mg = nullptr;
}

// Host side of the bounded wait: call after the stream has been synchronised.
int k_mg_check(pycs_handle h) {
  MgpuState* mg = h->mg;
  if (!mg) return 0;
  int e = 0;
  CK(cudaMemcpy(&e, &mg->sync->err, sizeof(int), cudaMemcpyDeviceToHost));
  if (e) {
    pycs_set_error("multi-GPU step: a peer's flag did not arrive within PYCS_MG_TIMEOUT_S (a rank is gone or "
                   "never entered the run); the state of this handle is undefined");
    return PYCS_ERR_STATE;
  }
  return 0;
}

int k_mg_exchange(pycs_handle h, const double* qnext, StepCtl* ctl, cudaStream_t st) {
  MgpuState* mg = h->mg;
  if (!mg->connected) {
    pycs_set_error("multi-GPU step before pycs_mgpu_connect");
    return PYCS_ERR_STATE;
  }
  const int idx = (qnext == mg->alloc[0]) ? 0 : 1;
  PeerSync ps;
  for (int d = 0; d < MG_MAX_WORLD; ++d) ps.s[d] = d < mg->world ? mg->peer_sync[d] : nullptr;
  // (a rank always has something to send: at least the halo rows of its slab neighbours)
  mg_exchange_kernel<<<mg->njobs, 256, 0, st>>>(h->g, mg->jobs_dev[idx], qnext, ps, mg->world, mg->rank, ctl, mg->counter);
  CKL(h);
  return 0;
}

int k_mg_device_barrier(pycs_handle h, cudaStream_t st) {
  MgpuState* mg = h->mg;
  if (!mg || !mg->connected) return 0;
  PeerSync ps;
  for (int d = 0; d < MG_MAX_WORLD; ++d) ps.s[d] = d < mg->world ? mg->peer_sync[d] : nullptr;
  mg->bcount += 1;
  mg_barrier_kernel<<<1, 32, 0, st>>>(ps, mg->sync, mg->world, mg->rank, mg->bcount, mg->timeout_ns);
  CKL(h);
  return 0;
}
int k_mg_quiesce_raise(pycs_handle h, cudaStream_t st) {
  MgpuState* mg = h->mg;
  if (!mg->connected) return 0;
  PeerSync ps;
  for (int d = 0; d < MG_MAX_WORLD; ++d) ps.s[d] = d < mg->world ? mg->peer_sync[d] : nullptr;
  mg->qcount += 1;
  mg_quiesce_raise_kernel<<<1, 32, 0, st>>>(ps, mg->world, mg->rank, mg->qcount);
  CKL(h);
  return 0;
}
int k_mg_quiesce_wait(pycs_handle h, cudaStream_t st) {
  MgpuState* mg = h->mg;
  mg_quiesce_wait_kernel<<<1, 32, 0, st>>>(mg->sync, mg->world, mg->qcount, mg->timeout_ns);
  CKL(h);
  return 0;
}

// stream-ordered wait until every rank has completed as many steps as this one (before the pending
// projection term of the last step is flushed: it needs every rank's sum)
extern StepCtl* k_fused_ctl(pycs_handle h);
int k_mg_wait_steps(pycs_handle h, cudaStream_t st) {
  MgpuState* mg = h->mg;
  mg_wait_steps_kernel<<<1, 32, 0, st>>>(mg->sync, k_fused_ctl(h), mg->world, mg->timeout_ns);
  CKL(h);
  return 0;
}
