// Fused advection step, v3: warp-autonomous column strips fed by a TMA producer warp.
//
// Same arithmetic as the v2 kernel in fused.cu (one kernel = src/discrete_operators.py:18-101
// + src/advection_timestep.py:43 for the ET-DG schemes), re-organised around what ncu showed
// v2 to be bound by -- instruction issue (336 instructions per thread-row), two block
// barriers per row and fixed-latency FP64 dependency stalls (profiles/r1_fused_v2_ncu.md):
//
//   * a CTA = NW consumer warps + 1 producer warp.  The producer streams rows of the seven
//     input arrays (Q, v, u, sqrtg at centres / x-edges / y-edges, 1/sqrtg) into a D-deep
//     shared-memory ring with cp.async.bulk (TMA, mbarrier complete_tx), adds the pending
//     MF-PR projection term to the Q row, and hands the slot to the consumers through a
//     second mbarrier.  Consumers give slots back through a third one.  No __syncthreads in
//     the row loop; warps drift freely and hide each other's FP64 latency;
//   * a consumer warp marches its own 64 columns down the chunk, each lane owning two
//     adjacent columns (16-byte shared / global accesses, two independent FP64 chains).  The
//     x-sweeps (stencil across rows) live in registers as rolling 5-row windows; the
//     y-sweeps read the staged Q row and a warp-private Qx row; neighbouring lanes exchange
//     the flux at the shared edge with a warp shuffle.  Warps overlap by 6-7 columns (the
//     7x7 dependence box of the split scheme), which are recomputed, not exchanged;
//   * the flux is evaluated in weight form (fused3_core.cuh): ~100 FP64 instructions per
//     cell instead of ~127, and no per-row ramp-up predicates: the first rows run on
//     zero-initialised windows and only the stores are predicated.
//
// Supported: PPM-0 / PPM-PL07 (linear reconstructions), every splitting, RK1 / RK2 masks,
// separable wind.  PPM-CW84 / PPM-L04 stay on the v2 kernel.
#include "fused_args.cuh"
#include "fused3_core.cuh"

namespace {

using namespace f3;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_row(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// A protocol error must not hang the GPU: give up after ~2^24 polls (seconds) and trap.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t n = 0; !mbar_try_wait(bar, parity); ++n)
    if (n > (1u << 24)) __trap();
}

template <int RECON, int SPLIT, int MASK, int NW, int D>
__global__ void __launch_bounds__((NW + 1) * 32, (NW <= 3 ? 4 : (NW == 4 ? 3 : 1))) fused3_kernel(FusedArgs a) {
  constexpr int RW = RowWidth<NW>::value;
  constexpr int NARR = (MASK & 1) ? 9 : 7;
  constexpr int SLOT = NARR * RW;             // doubles per ring slot
  constexpr int PF = D - 4;                   // rows in flight ahead of the slowest consumer
  static_assert(D >= 5, "rows r-3..r are live: the ring needs at least 5 slots");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* ring = reinterpret_cast<double*>(smem_raw);              // [D][NARR][RW]
  double* sxall = ring + D * SLOT;                                 // [NW][SXW]
  uint64_t* full = reinterpret_cast<uint64_t*>(sxall + NW * SXW);  // [D] TMA landed
  uint64_t* ready = full + D;                                      // [D] row patched, consumers may read
  uint64_t* empty = ready + D;                                     // [D] all consumers done with the slot

  const Geo& g = a.g;
  int b = blockIdx.x;
  const int p = b % 6;
  b /= 6;
  const int strip = b % a.nstrips, chunk = b / a.nstrips;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int js0 = g.lo + strip * a.wcols;
  const int js1 = min(js0 + a.wcols, g.hi);
  const int r0 = g.lo + chunk * a.rows_per_chunk;
  const int r1 = min(r0 + a.rows_per_chunk, g.hi);
  const int rfirst = r0 - 3, rlast = r1 + 2;
  const int c0 = ((js0 - 3) & ~1) - 4;        // first staged column (16-byte aligned: JOFF, c0 even)
  const int len = min(RW, g.ld - PYCS_JOFF - c0) & ~1;

  for (int k = tid; k < D * SLOT + NW * SXW; k += (NW + 1) * 32) ring[k] = 0.0;
  if (tid == 0) {
    for (int s = 0; s < D; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&ready[s], 1);
      mbar_init(&empty[s], NW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == NW) {
    // ------------------------------------------------------------------ producer warp
    const uint32_t row_bytes = (uint32_t)len * 8u;
    const long long colb = (long long)p * g.ps + PYCS_JOFF + c0;   // per-panel arrays
    const long long colm = PYCS_JOFF + c0;                         // panel-independent metric arrays
    const double corr = a.apply_corr ? *a.corr : 0.0;
    auto issue = [&](int r) {                                      // lane 0 only
      const int s = (r - rfirst) % D;
      double* dst = ring + s * SLOT;
      const long long rr = (long long)r * g.ld;
      mbar_expect_tx(&full[s], row_bytes * NARR);
      tma_row(dst + A_Q * RW, a.q + colb + rr, row_bytes, &full[s]);
      tma_row(dst + A_V * RW, a.va + colb + rr, row_bytes, &full[s]);
      tma_row(dst + A_SGC * RW, a.sgc + colm + rr, row_bytes, &full[s]);
      tma_row(dst + A_SGV * RW, a.sgv + colm + rr, row_bytes, &full[s]);
      tma_row(dst + A_RGC * RW, a.rgc + colm + rr, row_bytes, &full[s]);
      tma_row(dst + A_SGU * RW, a.sgu + colm + rr, row_bytes, &full[s]);
      tma_row(dst + A_U * RW, a.ua + colb + rr, row_bytes, &full[s]);
      if (MASK & 1) {
        tma_row(dst + A_VM * RW, a.vm + colb + rr, row_bytes, &full[s]);
        tma_row(dst + A_UM * RW, a.um + colb + rr, row_bytes, &full[s]);
      }
    };
    if (lane == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // zero fill above -> TMA writes
      for (int r = rfirst; r < rfirst + PF && r <= rlast; ++r) issue(r);
    }
    for (int r = rfirst; r <= rlast; ++r) {
      const int rn = r + PF;
      if (rn <= rlast) {
        const int n = (rn - rfirst) / D, s = (rn - rfirst) % D;
        if (n > 0) mbar_wait(&empty[s], (uint32_t)(n - 1) & 1u);
        if (lane == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          issue(rn);
        }
      }
      const int s = (r - rfirst) % D;
      mbar_wait(&full[s], (uint32_t)((r - rfirst) / D) & 1u);
      if (a.apply_corr && r >= g.lo && r < g.hi) {
        // pending MF-PR term of the previous step on the interior cells of this row
        double* qrow = ring + s * SLOT + A_Q * RW;
        const double* srow = ring + s * SLOT + A_SGC * RW;
        for (int k = lane; k < len; k += 32) {
          const int j = c0 + k;
          if (j >= g.lo && j < g.hi) qrow[k] = fma(srow[k], corr, qrow[k]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&ready[s]);
    }
    return;
  }

  // -------------------------------------------------------------------- consumer warps
  int cw0, us, ue;
  warp_columns(js0, js1, warp, cw0, us, ue);
  const int ca = cw0 - c0 + NC * lane;               // lane's first column inside a staged row
  const int col = cw0 + NC * lane;                   // ... and inside the panel
  double* sx = sxall + warp * SXW + 4 + NC * lane;   // own pair of the private Qx row
  const bool use0 = col >= us && col < ue, use1 = col + 1 >= us && col + 1 < ue;
  const double cdx = a.cdx, cdy = a.cdy, ws = a.ws;
  double* __restrict__ QN = a.qn + (long long)p * g.ps + PYCS_JOFF + min(col, g.P - 2) + (long long)r0 * g.ld;

  Lane L;
  lane_init(L);
  int o0 = 0, o1 = (D - 1) * SLOT, o2 = (D - 2) * SLOT, o3 = (D - 3) * SLOT;   // slots of rows r .. r-3
  int s0 = 0, s3 = D - 3;
  uint32_t par = 0;
#pragma unroll 1
  for (int r = rfirst; r <= rlast; ++r) {
    mbar_wait(&ready[s0], par);
    const double *R0 = ring + o0, *R1 = ring + o1, *R2 = ring + o2, *R3 = ring + o3;
    XEdge X;
    double qx[NC];
    phase_x_inner<RECON, SPLIT, MASK, RW>(L, X, R0, R1, R2, R3, ca, cdx, ws, qx);
    st2(sx, qx[0], qx[1]);
    __syncwarp();
    double F[NC + 1], G[NC + 1], CF[NC + 1], CG[NC];
    yflux_pair<RECON, SPLIT, MASK, RW>(R0, ca, R0 + A_Q * RW + ca, cdy, ws, F, CF);
    yflux_pair<RECON, SPLIT, MASK, RW>(R3, ca, sx, cdy, ws, G, CG);
    F[NC] = __shfl_down_sync(0xffffffffu, F[0], 1);
    G[NC] = __shfl_down_sync(0xffffffffu, G[0], 1);
    if (SPLIT != 1) CF[NC] = __shfl_down_sync(0xffffffffu, CF[0], 1);
    else CF[0] = CF[1] = CF[2] = 0.0;
    double out[NC], sdiv[NC];
    phase_x_outer<RECON, SPLIT, RW>(L, X, R0, ca, F, G, CF, out, sdiv);
    if (r >= r0 + 3) {                               // output row r-3
      if (use0 && use1) st2(QN, out[0], out[1]);
      else if (use0) QN[0] = out[0];
      else if (use1) QN[1] = out[1];
      if (use0) L.psum += sdiv[0];
      if (use1) L.psum += sdiv[1];
      QN += g.ld;
    }
    if (lane == 0 && r >= rfirst + 3) mbar_arrive(&empty[s3]);   // row r-3 is no longer needed
    o3 = o2; o2 = o1; o1 = o0;
    o0 = (o0 + SLOT == D * SLOT) ? 0 : o0 + SLOT;
    s3 = (s3 + 1 == D) ? 0 : s3 + 1;
    s0 = (s0 + 1 == D) ? 0 : s0 + 1;
    if (s0 == 0) par ^= 1u;
  }
  // per-warp partial of sum(pxdF + pydF) over its outputs (MF-PR), fixed order
  double v = L.psum;
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if (lane == 0) a.part[(long long)blockIdx.x * NW + warp] = v;
}

template <int NW, int D, int MASK>
constexpr size_t smem_bytes() {
  return sizeof(double) * ((size_t)D * ((MASK & 1) ? 9 : 7) * RowWidth<NW>::value + (size_t)NW * SXW) +
         sizeof(uint64_t) * 3 * D + 16;
}

template <int RECON, int SPLIT, int MASK, int NW, int D>
cudaError_t launch_one(const FusedArgs& a, int nblocks, cudaStream_t st, int* resident) {
  static bool configured = false;
  const size_t smem = smem_bytes<NW, D, MASK>();
  auto kern = fused3_kernel<RECON, SPLIT, MASK, NW, D>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (resident) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(resident, kern, (NW + 1) * 32, smem);
  kern<<<nblocks, (NW + 1) * 32, smem, st>>>(a);
  return cudaSuccess;
}

template <int RECON, int SPLIT, int NW, int D>
cudaError_t launch_mask(const FusedArgs& a, int mask, int nblocks, cudaStream_t st, int* resident) {
  if (mask == 1) return launch_one<RECON, SPLIT, 1, NW, D>(a, nblocks, st, resident);
  if (mask == 2) return launch_one<RECON, SPLIT, 2, NW, D>(a, nblocks, st, resident);
  return launch_one<RECON, SPLIT, 0, NW, D>(a, nblocks, st, resident);
}

// The par-default scheme (PPM-PL07 / SP-AVLT) is instantiated for every tuning point
// (consumer warps x ring depth); the other tuples use 3 consumer warps, depth 5.
cudaError_t dispatch(const FusedArgs& a, int recon, int split, int mask, int nw, int depth, int nblocks,
                     cudaStream_t st, int* resident) {
  if (recon == 3 && split == 1) {
#define TUNE(W, DD) \
  if (nw == W && depth == DD) return launch_mask<3, 1, W, DD>(a, mask, nblocks, st, resident)
    TUNE(3, 5); TUNE(3, 6); TUNE(3, 7);
    TUNE(4, 5); TUNE(4, 6);
    TUNE(7, 5); TUNE(7, 6);
#undef TUNE
    return cudaErrorInvalidValue;
  }
  if (nw != 3 || depth != 5) return cudaErrorInvalidValue;
#define CASE(R, S) \
  if (recon == R && split == S) return launch_mask<R, S, 3, 5>(a, mask, nblocks, st, resident)
  CASE(3, 2); CASE(3, 3); CASE(1, 1); CASE(1, 2); CASE(1, 3);
#undef CASE
  return cudaErrorInvalidValue;
}

}  // namespace

bool pycs_fused3_has(int recon, int split, int nw, int depth) {
  if (recon != 1 && recon != 3) return false;
  if (recon == 3 && split == 1)
    return ((nw == 3 && depth >= 5 && depth <= 7) || ((nw == 4 || nw == 7) && (depth == 5 || depth == 6)));
  return nw == 3 && depth == 5;
}

cudaError_t pycs_launch_fused3(const FusedArgs& a, int recon, int split, int mask, int nw, int depth, int nblocks,
                               cudaStream_t st) {
  return dispatch(a, recon, split, mask, nw, depth, nblocks, st, nullptr);
}

int pycs_fused3_resident(int recon, int split, int mask, int nw, int depth) {
  FusedArgs a{};
  int n = 0;
  if (dispatch(a, recon, split, mask, nw, depth, 0, nullptr, &n) != cudaSuccess) return -1;
  return n;
}
