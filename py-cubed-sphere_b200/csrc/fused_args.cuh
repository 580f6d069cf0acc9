// Arguments of the fused step kernels (fused.cu: v2 block-synchronous kernel;
// fused3.cu: v3 warp-autonomous kernel).
#pragma once
#include "pycs_common.cuh"

// Multi-GPU part of a fused launch (csrc/mgpu.cu): with world > 1 the step kernel itself stores
// the cells its peers need (boundary strips, halo rows) into their Q arrays over NVLink and its
// last CTA publishes the MF-PR sum and raises the flags -- compute and exchange in ONE kernel.
struct MgSync;
struct FusedMg {
  int world, rank, parity;
  long long epoch;
  double* peer_qn[8];         // the peers' output arrays (same layout, same positions)
  MgSync* peer_sync[8];
};

// Ghost fill fused into the step kernel (fused2b.cu, single GPU): every CTA computes the ghost
// cells its own rows / columns touch before it stages them, so the stand-alone ghost-fill launch
// disappears from the step.
struct FusedGhost {
  int enable, order, nsums;
  const int* kminE;
  const double* wE;
  const double* gs;           // ghost(sqrtg), see dg_fill_fused_kernel
  const double* sums;         // MF-PR sums of the previous step (nsums of them)
  double inv_a2;
  double* corr_out;           // the projection coefficient of this launch, for the ring restore
  HaloMaps maps;
};

struct FusedArgs {
  Geo g;
  const double* q;
  double* qn;
  const double *ua, *va;      // U_pu.ucontra_averaged, U_pv.vcontra_averaged
  const double *um, *vm;      // mask sources (U_pu.ucontra, U_pv.vcontra)
  const double *sgc, *rgc, *sgu, *sgv;
  double* part;               // per-CTA (v3: per-warp) partial sums of pxdF + pydF
  double* sum_out;            // their total, written by the last CTA of the launch (fixed order)
  unsigned* counter;          // CTAs (warps) that have written their partial; reset by the last one
  const double* corr;         // device scalar: pending projection coefficient -sum(s)/a2
  int rows_per_chunk, nstrips, wcols, apply_corr;
  int row_lo, row_hi;         // rows this launch updates (the whole interior, or this rank's slab)
  double cdx, cdy;            // dt/dx, dt/dy
  double ws;                  // separable wind: U(t) = U(0) * ws (MASK & 2)
  FusedMg mg;                 // world <= 1: single GPU
  int pdl;                    // launch with programmatic stream serialization (v2b)
  FusedGhost gf;
  // split launches (v2b, PYCS_SPLIT=1): a step is two launches over disjoint CTA sets -- the
  // interior CTAs, which read no ghost cell, start at once; the boundary CTAs follow the ghost
  // fill on a second stream.  blk_map[blockIdx.x] = CTA index in the full grid; both launches
  // share part[], the ticket counter and nblk_total, so the later one totals the MF-PR sum.
  const int* blk_map;         // nullptr: one launch over the whole grid
  int nblk_total;
  // several GPUs: the peers' MF-PR sums (gf.sums) are valid once every flag has reached wait_epoch; the
  // interior launch waits for them itself (the boundary launch follows the ghost fill, which has waited)
  const long long* wait_flags;
  int wait_world;
  long long wait_epoch;
  int* mg_err;                       // bounded wait: see mg_wait_flag (mgpu.cuh)
  unsigned long long mg_timeout_ns;
};

// CTA sets of a split step: strips x chunks x 6 panels, CTA = (chunk * nstrips + strip) * 6 + panel.
// Interior: strips 1 .. nstrips-2 and chunks 1 .. nchunks-2 (their staged rows / columns hold no
// ghost cell and no row of another slab).  Returns the number of interior CTAs (0: no split).
int pycs_split_sets(int nstrips, int nchunks, int* interior, int* boundary);

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

// v3 launcher: nw consumer warps, pf rows in flight, minb = register cap as CTAs per SM
cudaError_t pycs_launch_fused3(const FusedArgs& a, int recon, int split, int mask, int nw, int pf, int minb,
                               int nblocks, cudaStream_t st);
// resident CTAs per SM of the v3 kernel for this template point (occupancy API); < 0 on error
int pycs_fused3_resident(int recon, int split, int mask, int nw, int pf, int minb);
bool pycs_fused3_has(int recon, int split, int nw, int pf, int minb);
// v2b launcher (fused2b.cu): tb threads per CTA, pf rows in flight, minb = register cap as CTAs per SM
cudaError_t pycs_launch_fused2b(const FusedArgs& a, int recon, int split, int mask, int tb, int pf, int minb,
                                int nblocks, cudaStream_t st);
int pycs_fused2b_resident(int recon, int split, int mask, int tb, int pf, int minb);
bool pycs_fused2b_has(int recon, int split, int tb, int pf, int minb);
