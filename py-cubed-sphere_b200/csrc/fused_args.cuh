// Arguments of the fused step kernels (fused2b.cu: the production kernel; fused.cu: the first-generation
// kernel that still serves the limited reconstructions PPM-CW84 / PPM-L04).
#pragma once
#include <vector>
#include "pycs_common.cuh"

struct MgSync;

// Device-side control block of the fused step loop (one per handle, device memory).  Everything that
// changes from step to step lives here instead of in kernel parameters, so that the launches of a step
// are identical from one step to the next and can be replayed from a CUDA graph:
//   * steps      -- fused steps completed; read by every CTA of a step at its start, advanced by the
//                   step's last CTA.  It indexes the separable-wind time factors (ws_tab) and is the
//                   epoch of the multi-GPU MF-PR sums (flag `sflag`, mgpu.cuh);
//   * xcount     -- multi-GPU halo exchanges delivered by this rank (epoch of `dflag`);
//   * pend       -- the MF-PR sum(s) of the last step still have to be applied to what it wrote;
//   * corr_applied -- the projection coefficient the last step applied to what it read (ring restore);
//   * sum        -- single GPU: total of the last step's partial sums of pxdF + pydF.
struct StepCtl {
  long long steps;
  long long xcount;
  int pend;
  int pad_;
  double corr_applied;
  double sum;
};

// Multi-GPU: where the last CTA of a step publishes this rank's MF-PR sum and raises `sflag`.
struct FusedMgPub {
  int world, rank;            // world <= 1: single GPU
  MgSync* peer_sync[8];       // peer_sync[rank] is this rank's own block
};

struct FusedArgs {
  Geo g;
  const double* q;
  double* qn;
  const double *ua, *va;      // U_pu.ucontra_averaged, U_pv.vcontra_averaged (MASK & 2: the t = 0 winds)
  const double *um, *vm;      // mask sources (U_pu.ucontra, U_pv.vcontra)
  const double *sgc, *rgc, *sgu, *sgv;
  double* part;               // per-CTA partial sums of pxdF + pydF
  unsigned* counter;          // CTAs that have written their partial; reset by the last one
  StepCtl* ctl;
  const double* ws_tab;       // separable wind (MASK & 2): U(t_k) = U(0) * ws_tab[steps & ws_mask]
  int ws_mask;
  // pending MF-PR projection term of the previous step, added to Q as it is loaded:
  //   corr_ptr != null: coefficient written by the ghost-fill kernel of this step (serial path);
  //   else formed here from the previous step's sum(s): -sum / a2 (same expression, same bits)
  int apply_corr;             // the scheme has MF-PR
  const double* corr_ptr;
  double inv_a2;
  const double* gs;           // GH = 1: ghost(sqrtg) for the projection term on ghost cells (raw ghost fill)
  int rows_per_chunk, nstrips, wcols;
  int row_lo, row_hi;         // rows this launch updates (the whole interior, or this rank's slab)
  double cdx, cdy;            // dt/dx, dt/dy
  // split step: a step is two launches over disjoint CTA sets -- boundary CTAs (everything a ghost cell or
  // a peer reads, launched first on a high-priority stream, followed by the exchange and the next ghost
  // fill) and interior CTAs.  The CTAs are listed in a table (boundary CTAs first): CTA blockIdx.x of a
  // launch is entry cta_off + blockIdx.x = (first row, end row, strip, panel).  The boundary CTAs march
  // few rows each, so that they are done -- and the exchange under way -- long before the interior
  // CTAs.  Both launches share part[], the ticket counter and nblk_total: whichever finishes last
  // closes the step.
  const int4* cta_tab;        // nullptr: one launch over the uniform grid strips x chunks x 6
  int cta_off;
  int nblk_total;
  // several GPUs: the peers' sums (and the right to overwrite their buffers) arrive with sflag >= steps
  const long long* wait_flags;
  int wait_world;
  int* mg_err;                // bounded wait: see mg_wait_flag (mgpu.cuh)
  unsigned long long mg_timeout_ns;
  FusedMgPub pub;
  // several GPUs, boundary launch (GH = 1): the exchange is part of the kernel.  Each boundary CTA, once its
  // rows are written, stores the pieces of them that peers read (xjobs[xjob_off[cta] .. xjob_off[cta+1]),
  // rectangles of its own panel) straight into the peers' output arrays; the last of the n_boundary CTAs
  // raises dflag on every rank.  No exchange kernel, no launch between the boundary CTAs and the peers.
  const int* xjob_off;        // nullptr: no in-kernel exchange
  const int4* xjobs;          // (peer | i0 << 4, i1, j0, j1)
  double* xpeer_q[8];         // the peers' arrays that correspond to qn
  unsigned* xcounter;
  int n_boundary;
  // MF-AF (GH = 1 flavour): the outer fluxes on the four edge lines of every panel are recorded here,
  // [panel][W, E, S, N][N], for the cube-edge averaging that follows the launch (stepper.cu: mf_af_patch_kernel)
  double* edge_flux;
  int timing;                 // roofline timing launches: leave the control block alone
  int pdl;                    // launch with the programmatic-dependent-launch attribute (pycs_common.cuh)
  // GH = 2 (one-kernel step): the ghost cells a CTA's march reads are filled by the CTA itself before its first
  // row copy -- the Lagrange fill of ghost_core.cuh, raw (the projection term is added on load as for GH = 1) --
  // so a step is ONE launch: no ghost-fill kernel, no second stream.  Several GPUs: the first n_boundary CTAs of
  // the table (everything that reads a ghost cell or a peer's row) first wait for the peers' exchange flags.
  HaloMaps gf_maps;
  const int* gf_kmin;
  const double* gf_w;
  int gf_order;
  const long long* gf_flags;  // the peers' dflag (nullptr on one GPU)
};

// CTA table of a split step over the rows [row_lo, row_hi) of a panel of N x N cells cut into nstrips column
// strips: boundary CTAs first (row bands of `band` rows at both ends of the slab over all strips, then the
// first and last strip in chunks of `edge_rows` rows), then the interior CTAs (strips 1 .. nstrips-2 in
// chunks of `rows` rows).  Interior CTAs stage neither a ghost cell nor a row of another slab (band >= 3).
// Entries are (r0, r1, strip, panel).  Returns the number of boundary CTAs; a slab or panel too small to
// have an interior yields boundary CTAs only.
struct CtaDesc { int r0, r1, strip, panel; };
int pycs_plan_split_ctas(int row_lo, int row_hi, int nstrips, int band, int edge_rows, int rows,
                         std::vector<CtaDesc>* out);

#ifdef __CUDACC__
// The writer of the last partial of a launch adds them all up: consumers of the MF-PR sum
// (ghost fill, flush, multi-GPU exchange) then read one scalar instead of reducing the list.
__device__ __forceinline__ bool fused_last_writer(unsigned* counter, unsigned total, bool sys = false) {
  if (sys) __threadfence_system();           // peer stores of this CTA before the ticket
  else __threadfence();
  return atomicAdd(counter, 1u) == total - 1u;
}
// one full warp; the result is valid in lane 0
__device__ __forceinline__ double fused_warp_sum(const double* part, int n, int lane) {
  __threadfence();
  double v = 0.0;
  for (int k = lane; k < n; k += 32) v += __ldcg(part + k);
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
#endif

// v2b launcher (fused2b.cu): gh = 1 launches the flavour that adds the projection term on ghost cells itself
cudaError_t pycs_launch_fused2b(const FusedArgs& a, int recon, int split, int mask, int gh, int nblocks, cudaStream_t st);
// resident CTAs per SM of that instantiation (occupancy API; also sets its shared-memory attribute, which
// must not happen inside a stream capture); < 0 on error
int pycs_fused2b_resident(int recon, int split, int mask, int gh = 0);
bool pycs_fused2b_has(int recon, int split);
int pycs_fused2b_threads();
// v2 launcher (fused.cu): limited reconstructions, single GPU, serial path only
cudaError_t pycs_launch_fused_v2(const FusedArgs& a, int recon, int split, int mask, int nblocks, cudaStream_t st);
int pycs_fused_v2_threads();
int pycs_fused_v2_resident();
