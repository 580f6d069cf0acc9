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

template <int RECON, int SPLIT, int MASK, int NW, int PF, int MINB>
__global__ void __launch_bounds__((NW + 1) * 32, MINB % 10) fused3_kernel(FusedArgs a) {
  // MINB >= 10: same register cap (MINB - 10 CTAs/SM), march loop unrolled by the window length
  constexpr int RW = RowWidth<NW>::value;
  constexpr int NS = (MASK & 1) ? 4 : 3, NL = (MASK & 1) ? 5 : 4;   // arrays per short / long slot
  constexpr int DS = PF + 1, DL = PF + 4, ND = PF + 2;              // ring depths, "iteration done" barriers
  constexpr int SSLOT = NS * RW, LSLOT = NL * RW;                   // doubles per slot
  static_assert(PF >= 1, "at least one row in flight");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* ringS = reinterpret_cast<double*>(smem_raw);             // [DS][NS][RW]
  double* ringL = ringS + DS * SSLOT;                              // [DL][NL][RW]
  double* sxall = ringL + DL * LSLOT;                              // [NW][SXW]
  uint64_t* full = reinterpret_cast<uint64_t*>(sxall + NW * SXW);  // [DL] TMA of a row landed
  uint64_t* ready = full + DL;                                     // [DL] row patched (MF-PR pending)
  uint64_t* done = ready + DL;                                     // [ND] a march step finished by all consumers

  const Geo& g = a.g;
  int b = blockIdx.x;
  const int p = b % 6;
  b /= 6;
  const int strip = b % a.nstrips, chunk = b / a.nstrips;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int js0 = g.lo + strip * a.wcols;
  const int js1 = min(js0 + a.wcols, g.hi);
  const int r0 = a.row_lo + chunk * a.rows_per_chunk;
  const int r1 = min(r0 + a.rows_per_chunk, a.row_hi);
  const int rfirst = r0 - 3, rlast = r1 + 2;
  const int c0 = strip_c0(js0);               // first staged column (16-byte aligned: JOFF, c0 even)
  const int len = min(RW, g.ld - PYCS_JOFF - c0) & ~1;

  for (int k = tid; k < DS * SSLOT + DL * LSLOT + NW * SXW; k += (NW + 1) * 32) ringS[k] = 0.0;
  if (tid == 0) {
    for (int s = 0; s < DL; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&ready[s], 1);
    }
    for (int s = 0; s < ND; ++s) mbar_init(&done[s], NW);
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
      const int k = r - rfirst;
      double* dS = ringS + (k % DS) * SSLOT;
      double* dL = ringL + (k % DL) * LSLOT;
      uint64_t* bar = &full[k % DL];
      const long long rr = (long long)r * g.ld, r1_ = (long long)max(r - 1, 0) * g.ld,
                      r2_ = (long long)max(r - 2, 0) * g.ld;
      mbar_expect_tx(bar, row_bytes * (NS + NL));
      tma_row(dS + S_Q * RW, a.q + colb + rr, row_bytes, bar);
      tma_row(dL + L_V * RW, a.va + colb + rr, row_bytes, bar);
      tma_row(dL + L_SGC * RW, a.sgc + colm + rr, row_bytes, bar);
      tma_row(dL + L_SGV * RW, a.sgv + colm + rr, row_bytes, bar);
      tma_row(dL + L_RGC * RW, a.rgc + colm + rr, row_bytes, bar);
      tma_row(dS + S_SGU * RW, a.sgu + colm + r1_, row_bytes, bar);
      tma_row(dS + S_U * RW, a.ua + colb + r2_, row_bytes, bar);
      if (MASK & 1) {
        tma_row(dL + L_VM * RW, a.vm + colb + rr, row_bytes, bar);
        tma_row(dS + S_UM * RW, a.um + colb + r2_, row_bytes, bar);
      }
    };
    if (lane == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // zero fill above -> TMA writes
      for (int r = rfirst; r < rfirst + PF && r <= rlast; ++r) issue(r);
    }
    for (int r = rfirst; r <= rlast; ++r) {
      const int rn = r + PF;
      if (rn <= rlast) {
        // the slots of row rn were last read in march step rn - PF - 1
        const int m = rn - PF - 1 - rfirst;
        if (m >= 0) mbar_wait(&done[m % ND], (uint32_t)(m / ND) & 1u);
        if (lane == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          issue(rn);
        }
      }
      if (a.apply_corr) {
        // pending MF-PR term of the previous step on the interior cells of the Q row
        const int k = r - rfirst;
        mbar_wait(&full[k % DL], (uint32_t)(k / DL) & 1u);
        if (r >= g.lo && r < g.hi) {
          double* qrow = ringS + (k % DS) * SSLOT + S_Q * RW;
          const double* srow = ringL + (k % DL) * LSLOT + L_SGC * RW;
          for (int i = lane; i < len; i += 32) {
            const int j = c0 + i;
            if (j >= g.lo && j < g.hi) qrow[i] = fma(srow[i], corr, qrow[i]);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready[k % DL]);
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumer warps
  int cw0, us, ue;
  warp_columns(js0, js1, warp, cw0, us, ue);
  const int ca = cw0 - c0 + lane;                    // lane's first column inside a staged row
  const int col = cw0 + lane;                        // ... and inside the panel (second: + CSTEP)
  double* sx = sxall + warp * SXW + 4 + lane;        // own columns of the private Qx row
  const bool use0 = col >= us && col < ue, use1 = col + CSTEP >= us && col + CSTEP < ue;
  const double cdx = a.cdx * ((MASK & 2) ? a.ws : 1.0), cdy = a.cdy * ((MASK & 2) ? a.ws : 1.0);   // time factor folded in
  uint64_t* const rowbar = a.apply_corr ? ready : full;
  // row r0 of the output panel; own columns (clamped for lanes that own no output)
  double* __restrict__ QN = a.qn + (long long)p * g.ps + PYCS_JOFF + (long long)r0 * g.ld;
  const int qo0 = min(col, g.P - 1), qo1 = min(col + CSTEP, g.P - 1);

  Lane L;
  lane_init(L);
  int oS = 0;                                                                       // short slot of row r
  int oL0 = 0, oL1 = (DL - 1) * LSLOT, oL2 = (DL - 2) * LSLOT, oL3 = (DL - 3) * LSLOT;   // long slots of rows r .. r-3
  int sb = 0, sd = 0;                                                               // barrier indices of this step
  uint32_t parb = 0, pard = 0;
#pragma unroll(MINB >= 10 ? 5 : 1)
  for (int r = rfirst; r <= rlast; ++r) {
    mbar_wait(&rowbar[sb], parb);
    RowPtrs R;
    R.q = ringS + oS + S_Q * RW + ca;
    R.u = ringS + oS + S_U * RW + ca;
    R.um = ringS + oS + S_UM * RW + ca;
    R.su1 = ringS + oS + S_SGU * RW + ca;
    R.v0 = ringL + oL0 + L_V * RW + ca;
    R.vm0 = ringL + oL0 + L_VM * RW + ca;
    R.sgv0 = ringL + oL0 + L_SGV * RW + ca;
    R.sgc0 = ringL + oL0 + L_SGC * RW + ca;
    R.rg0 = ringL + oL0 + L_RGC * RW + ca;
    R.sgc2 = ringL + oL2 + L_SGC * RW + ca;
    R.v3 = ringL + oL3 + L_V * RW + ca;
    R.vm3 = ringL + oL3 + L_VM * RW + ca;
    R.sgv3 = ringL + oL3 + L_SGV * RW + ca;
    R.sgc3 = ringL + oL3 + L_SGC * RW + ca;
    R.rg3 = ringL + oL3 + L_RGC * RW + ca;
    XEdge X;
    double qx[NC];
    const double qnew[NC] = {R.q[0], R.q[CSTEP]};
    phase_x_inner<RECON, SPLIT, MASK>(L, X, R, qnew, cdx, qx);
    sx[0] = qx[0];
    sx[CSTEP] = qx[1];
    __syncwarp();
    double F[NC], G[NC], CF[NC], CG[NC], Fn[NC], Gn[NC], CFn[NC];
    CF[0] = CF[1] = 0.0;
    yflux_pair<RECON, SPLIT, MASK>(R.v0, R.vm0, R.sgv0, R.sgc0, R.q, cdy, F, CF);
    yflux_pair<RECON, SPLIT, MASK>(R.v3, R.vm3, R.sgv3, R.sgc3, sx, cdy, G, CG);
    // flux at the column right of each own column: next lane's same column; for lane 31 the
    // right neighbour of column cw0+31 is lane 0's second column (its own second column's
    // neighbour lies outside the warp and is never needed)
    {
      const double f0 = __shfl_down_sync(0xffffffffu, F[0], 1), f1 = __shfl_down_sync(0xffffffffu, F[1], 1);
      const double fw = __shfl_sync(0xffffffffu, F[1], 0);
      const double g0 = __shfl_down_sync(0xffffffffu, G[0], 1), g1 = __shfl_down_sync(0xffffffffu, G[1], 1);
      const double gw = __shfl_sync(0xffffffffu, G[1], 0);
      Fn[0] = lane == 31 ? fw : f0; Fn[1] = f1;
      Gn[0] = lane == 31 ? gw : g0; Gn[1] = g1;
      if (SPLIT != 1) {
        const double c0_ = __shfl_down_sync(0xffffffffu, CF[0], 1), c1_ = __shfl_down_sync(0xffffffffu, CF[1], 1);
        const double cw = __shfl_sync(0xffffffffu, CF[1], 0);
        CFn[0] = lane == 31 ? cw : c0_; CFn[1] = c1_;
      } else {
        CFn[0] = CFn[1] = 0.0;
      }
    }
    double out[NC], sdiv[NC];
    phase_x_outer<RECON, SPLIT>(L, X, R, F, Fn, G, Gn, CF, CFn, out, sdiv);
    if (r >= r0 + 3) {                               // output row r-3
      if (use0) { QN[qo0] = out[0]; L.psum += sdiv[0]; }
      if (use1) { QN[qo1] = out[1]; L.psum += sdiv[1]; }
      QN += g.ld;
    }
    if (lane == 0) mbar_arrive(&done[sd]);           // this march step no longer needs its oldest rows
    oS = (oS + SSLOT == DS * SSLOT) ? 0 : oS + SSLOT;
    oL3 = oL2; oL2 = oL1; oL1 = oL0;
    oL0 = (oL0 + LSLOT == DL * LSLOT) ? 0 : oL0 + LSLOT;
    if (++sb == DL) { sb = 0; parb ^= 1u; }
    if (++sd == ND) { sd = 0; pard ^= 1u; }
  }
  (void)pard;
  // per-warp partial of sum(pxdF + pydF) over its outputs (MF-PR), fixed order
  double v = L.psum;
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  int last = 0;
  if (lane == 0) {
    a.part[(long long)blockIdx.x * NW + warp] = v;
    last = fused_last_writer(a.counter, gridDim.x * NW) ? 1 : 0;
  }
  last = __shfl_sync(0xffffffffu, last, 0);
  if (last) {                                  // last warp of the launch: total in a fixed order
    const double tot = fused_warp_sum(a.part, (int)(gridDim.x * NW), lane);
    if (lane == 0) {
      *a.sum_out = tot;
      *a.counter = 0u;
    }
  }
}

template <int NW, int PF, int MASK>
constexpr size_t smem_bytes() {
  return sizeof(double) * ((size_t)RowWidth<NW>::value * ((PF + 1) * ((MASK & 1) ? 4 : 3) + (PF + 4) * ((MASK & 1) ? 5 : 4)) +
                           (size_t)NW * SXW) +
         sizeof(uint64_t) * (2 * (PF + 4) + PF + 2) + 16;
}

template <int RECON, int SPLIT, int MASK, int NW, int PF, int MINB>
cudaError_t launch_one(const FusedArgs& a, int nblocks, cudaStream_t st, int* resident) {
  static bool configured = false;
  const size_t smem = smem_bytes<NW, PF, MASK>();
  auto kern = fused3_kernel<RECON, SPLIT, MASK, NW, PF, MINB>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (resident) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(resident, kern, (NW + 1) * 32, smem);
  kern<<<nblocks, (NW + 1) * 32, smem, st>>>(a);
  return cudaSuccess;
}

template <int RECON, int SPLIT, int NW, int PF, int MINB>
cudaError_t launch_mask(const FusedArgs& a, int mask, int nblocks, cudaStream_t st, int* resident) {
  if (mask == 1) return launch_one<RECON, SPLIT, 1, NW, PF, MINB>(a, nblocks, st, resident);
  if (mask == 2) return launch_one<RECON, SPLIT, 2, NW, PF, MINB>(a, nblocks, st, resident);
  return launch_one<RECON, SPLIT, 0, NW, PF, MINB>(a, nblocks, st, resident);
}

// Tuning points: consumer warps NW x rows in flight PF x register cap (MINB CTAs per SM).
// The par-default scheme (PPM-PL07 / SP-AVLT) has all of them, the other tuples the default.
#define F3_DEFAULT_NW 3
#define F3_DEFAULT_PF 3
#define F3_DEFAULT_MINB 3
cudaError_t dispatch(const FusedArgs& a, int recon, int split, int mask, int nw, int pf, int minb, int nblocks,
                     cudaStream_t st, int* resident) {
  if (recon == 3 && split == 1) {
#define TUNE(W, P, M) \
  if (nw == W && pf == P && minb == M) return launch_mask<3, 1, W, P, M>(a, mask, nblocks, st, resident)
    TUNE(3, 2, 3); TUNE(3, 3, 3); TUNE(3, 4, 3); TUNE(3, 2, 4); TUNE(3, 3, 4); TUNE(3, 3, 13); TUNE(3, 4, 13);
    TUNE(2, 3, 4); TUNE(2, 4, 4); TUNE(2, 4, 5);
    TUNE(4, 2, 2); TUNE(4, 3, 2); TUNE(4, 2, 3);
#undef TUNE
    return cudaErrorInvalidValue;
  }
  if (nw != F3_DEFAULT_NW || pf != F3_DEFAULT_PF || minb != F3_DEFAULT_MINB) return cudaErrorInvalidValue;
#define CASE(R, S) \
  if (recon == R && split == S)  \
    return launch_mask<R, S, F3_DEFAULT_NW, F3_DEFAULT_PF, F3_DEFAULT_MINB>(a, mask, nblocks, st, resident)
  CASE(3, 2); CASE(3, 3); CASE(1, 1); CASE(1, 2); CASE(1, 3);
#undef CASE
  return cudaErrorInvalidValue;
}

}  // namespace

bool pycs_fused3_has(int recon, int split, int nw, int pf, int minb) {
  if (recon != 1 && recon != 3) return false;
  if (recon == 3 && split == 1) {
    const int t[][3] = {{3, 2, 3}, {3, 3, 3}, {3, 4, 3}, {3, 2, 4}, {3, 3, 4}, {3, 3, 13}, {3, 4, 13}, {2, 3, 4}, {2, 4, 4}, {2, 4, 5},
                        {4, 2, 2}, {4, 3, 2}, {4, 2, 3}};
    for (auto& x : t)
      if (x[0] == nw && x[1] == pf && x[2] == minb) return true;
    return false;
  }
  return nw == F3_DEFAULT_NW && pf == F3_DEFAULT_PF && minb == F3_DEFAULT_MINB;
}

cudaError_t pycs_launch_fused3(const FusedArgs& a, int recon, int split, int mask, int nw, int pf, int minb,
                               int nblocks, cudaStream_t st) {
  return dispatch(a, recon, split, mask, nw, pf, minb, nblocks, st, nullptr);
}

int pycs_fused3_resident(int recon, int split, int mask, int nw, int pf, int minb) {
  FusedArgs a{};
  int n = 0;
  if (dispatch(a, recon, split, mask, nw, pf, minb, 0, nullptr, &n) != cudaSuccess) return -1;
  return n;
}
