// Fused advection step, v2b: the block-synchronous march of fused.cu (one column per thread,
// many warps per SM) running the lean arithmetic of fused3_core.cuh.
//
// ncu on v2 (profiles/r1_fused_v2_ncu.md) showed an instruction-issue bound kernel (336
// instructions per thread-row); v3 (fused3.cu) cut the instruction count but its two columns
// per lane cost so many registers that only 9 consumer warps fit on an SM
// (profiles/r1_fused_v3_ncu.md).  v2b keeps v2's occupancy (one column per thread, ~90
// registers, 4-5 CTAs of 160 threads per SM) and takes from v3:
//   * the weight form of the flux and the per-edge weights shared by inner / outer x-flux;
//   * no ramp-up predicates: the first rows of a chunk run on zero-initialised windows, only
//     the stores are predicated;
//   * the two-level row ring (Q, u, sqrtg_pu: one row + prefetch; v, sqrtg_pv, sqrtg_pc,
//     1/sqrtg_pc: four rows + prefetch), filled by thread 0 with cp.async.bulk (TMA) row
//     copies signalled on one mbarrier per row.
// Per marched row: phase 1 (own column: new row, inner x-flux, Qx) | barrier | phase 2
// (y-fluxes of Q row r and of Qx row r-3 at the thread's edge) | barrier | phase 3 (Qy, outer
// x-flux, output row r-3).
#include "fused_args.cuh"
#include "mgpu.cuh"
#include "ghost_core.cuh"
// PPM-PL07 edge-value coefficients as constant-bank operands (const-slot march only)
__constant__ double f2b_ppm_coef[5] = {2.0 / 60.0, -13.0 / 60.0, 47.0 / 60.0, 27.0 / 60.0, -3.0 / 60.0};
#define F3_COEF_BANK f2b_ppm_coef
#define F3_NAMESPACE f1
#define F3_NC 1
#define F3_CSTEP 0
#include "fused3_core.cuh"
#undef F3_COEF_BANK

#define F2B_CS_MINB 34     // the default march (const-slot, 4 CTAs/SM)

namespace {

using namespace f1;

template <int K> struct IC { static constexpr int value = K; };   // compile-time row phase

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
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

// one lane of a converged warp (the CUTLASS idiom: the compiler keeps the TMA operands in
// uniform registers instead of uniformising per-thread values with a loop per instruction)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}

// MG = 1: multi-GPU build -- the kernel stores boundary cells to the peers and publishes sum + flags
// (kept out of the single-GPU instantiation: the extra code in the unrolled march loop costs ~7 %).
// MG = 2: split-launch build (FusedArgs::blk_map): CTA index through the map, projection
// coefficient formed from the previous step's sum (no dependence on the ghost-fill kernel),
// partials / ticket shared with the other launch of the step.
template <int TB, int RECON, int SPLIT, int MASK, int PF, int MINB, int MG>
__global__ void __launch_bounds__(TB, MINB % 10) fused2b_kernel(FusedArgs a) {
  // MINB >= 10: register cap of MINB % 10 CTAs/SM and the march loop unrolled by the window length;
  // MINB >= 30: the march runs in groups of DL rows with every ring slot a compile-time constant
  // (all staged-row loads become [one base register + immediate]) and the TMA issue path keeps
  // running byte offsets instead of recomputing row * ld (see profiles/r1_sass_static.md)
  constexpr bool CS = ((MINB >= 30 && MINB < 50) || (MINB >= 60 && MINB < 70)) && MG != 1;
  // MINB >= 60 (const-slot march only; parity-checked at N=50 with the last GPU seconds of round 1, NOT yet timed):
  // barrier B of a row replaced by producer/consumer named barriers between neighbouring warps --
  // warp w only needs lane 0 of warp w+1 for F/G[e+1] -- with the work rows double-buffered by row
  // parity; barrier A (block-wide, once per row) still bounds the skew between warps to one row
  constexpr bool PB = CS && (MINB >= 60);
  constexpr int NWARP = TB / 32;
  // MINB >= 50: two rows per pair of block barriers (rings of 4 and 8 slots, windows of 8 registers)
  constexpr bool PAIR = (MINB >= 50) && (MINB < 60) && MG != 1;
  constexpr int RW = TB + 6;                 // staged row: columns jbase-6 .. jbase+TB-1
  constexpr int NS = (MASK & 1) ? 4 : 3, NL = (MASK & 1) ? 5 : 4;
  constexpr int DS = PAIR ? 4 : PF + 1, DL = PAIR ? 8 : PF + 4;
  constexpr int NWORK = (PAIR || PB) ? 8 : 4;        // work rows: Qx, inner / outer y-flux, sqrtg_pv*cy (per row of a pair)
  constexpr int SSLOT = NS * RW, LSLOT = NL * RW;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* ringS = reinterpret_cast<double*>(smem_raw);           // [DS][NS][RW]
  double* ringL = ringS + DS * SSLOT;                            // [DL][NL][RW]
  double* sX = ringL + DL * LSLOT;                               // Qx row r-3
  double* sF = sX + RW;                                          // inner y-flux, row r
  double* sG = sF + RW;                                          // outer y-flux, row r-3
  double* sC = sG + RW;                                          // sqrtg_pv*cy (SPLIT != 1)
  uint64_t* full = reinterpret_cast<uint64_t*>(sX + NWORK * RW);   // [DL]

  const Geo& g = a.g;
  int b = (MG == 2) ? a.blk_map[blockIdx.x] : (int)blockIdx.x;   // CTA index in the full grid
  const int p = b % 6;
  b /= 6;
  const int strip = b % a.nstrips, chunk = b / a.nstrips;
  const int tid = threadIdx.x;
  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp index, provably warp-uniform
  const int e = tid + 3;                     // own element in a staged row (the y-stencil reaches e-3 .. e+2)
  const int jbase = g.lo + strip * a.wcols;
  const int jend = min(jbase + a.wcols, g.hi);
  const int j = jbase - 3 + tid;
  const int r0 = a.row_lo + chunk * a.rows_per_chunk;
  const int r1 = min(r0 + a.rows_per_chunk, a.row_hi);
  const int rfirst = r0 - 3, rlast = r1 + 2;
  const bool out_lane = (tid >= 3) && (j < jend);
  const bool jint = (j >= g.lo) && (j < g.hi);
  const double cdx = a.cdx * ((MASK & 2) ? a.ws : 1.0), cdy = a.cdy * ((MASK & 2) ? a.ws : 1.0);   // time factor folded in
  const int mgw = (MG == 1) ? a.mg.world : 0;
  const int c0 = jbase - 6;                  // 16-byte aligned: JOFF and wcols are even
  const int len = min(RW, g.ld - PYCS_JOFF - c0) & ~1;
  const uint32_t row_bytes = (uint32_t)len * 8u;

  for (int k = tid; k < DS * SSLOT + DL * LSLOT + NWORK * RW; k += TB) ringS[k] = 0.0;
  if (tid == 0) {
    for (int s = 0; s < DL; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // programmatic dependent launch: everything above overlaps the tail of the previous kernel
  // (the ghost fill); its results (ghost cells, corr) are read only below.  No-ops otherwise.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  double corr = 0.0;
  if (MG == 2) {
    if (a.wait_flags) {         // same wait as dg_fill_fused_kernel (csrc/fused.cu)
      if (tid < a.wait_world) {
        mg_wait_flag(a.wait_flags + tid, a.wait_epoch, a.mg_err, a.mg_timeout_ns);   // bounded, see mgpu.cuh
        __threadfence_system();
      }
      __syncthreads();
    }
    if (a.apply_corr) {         // single GPU: same expression as the ghost-fill kernel, same bits
      double sm = 0.0;
      for (int k = 0; k < a.gf.nsums; ++k) sm += a.gf.sums[k];
      corr = -sm * a.gf.inv_a2;
    }
  } else if (a.gf.enable) {
    // ---- ghost prologue: the Lagrange ghost cells of the rectangle this CTA stages (rows
    // rfirst..rlast, columns jbase-3..jend+2), written in place before the TMA copies read them.
    // Neighbouring CTAs compute the ghost cells they share; the values are identical.
    if (a.apply_corr) {
      double sm = 0.0;
      for (int k = 0; k < a.gf.nsums; ++k) sm += a.gf.sums[k];
      corr = -sm * a.gf.inv_a2;
      if (blockIdx.x == 0 && tid == 0) *a.gf.corr_out = corr;
    }
    const int cl = jbase - 3, cr = min(jend + 2, g.P - 1);
    if (rfirst < g.lo || rlast >= g.hi || cl < g.lo || cr >= g.hi) {
      const int nC = cr - cl + 1, ncell = (rlast - rfirst + 1) * nC;
      double* qw = const_cast<double*>(a.q);
      for (int t = tid; t < ncell; t += TB) {
        const int i = rfirst + t / nC, jj = cl + t % nC;
        if (i >= g.lo && i < g.hi && jj >= g.lo && jj < g.hi) continue;
        double v = dg_ghost_cell(g, a.gf.maps, a.q, a.gf.kminE, a.gf.wE, a.gf.order, p, i, jj);
        const long long id = gidx(g, p, i, jj);
        if (a.apply_corr) v = fma(a.gf.gs[id], corr, v);
        qw[id] = v;
      }
      asm volatile("fence.proxy.async.global;" ::: "memory");   // generic stores -> TMA (async proxy) reads
      __threadfence();
    }
    __syncthreads();
  } else if (a.apply_corr) {
    corr = *a.corr;
  }

  // TMA row copies are issued by one elected lane of warp 0 (one row per marched row, PF rows ahead).  Its slot
  // offsets rotate in registers and the global addresses are base + row * ld: the issue path
  // sits between barrier A and barrier B of warp 0, so it is kept short.
  const uint32_t ringS_a = smem_u32(ringS), ringL_a = smem_u32(ringL), full_a = smem_u32(full);
  const double* const gq = a.q + (long long)p * g.ps + PYCS_JOFF + c0;
  const double* const gv = a.va + (long long)p * g.ps + PYCS_JOFF + c0;
  const double* const gu = a.ua + (long long)p * g.ps + PYCS_JOFF + c0;
  const double* const gvm = a.vm + (long long)p * g.ps + PYCS_JOFF + c0;
  const double* const gum = a.um + (long long)p * g.ps + PYCS_JOFF + c0;
  const double* const gsgc = a.sgc + PYCS_JOFF + c0;
  const double* const gsgv = a.sgv + PYCS_JOFF + c0;
  const double* const grgc = a.rgc + PYCS_JOFF + c0;
  const double* const gsgu = a.sgu + PYCS_JOFF + c0;
  int iS = 0, iL = 0, ib = 0;                // slot byte offsets / barrier index of the next row to issue
  auto tma = [&](uint32_t dst, const double* src, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(row_bytes), "r"(bar)
                 : "memory");
  };
  auto issue = [&](int r) {                  // elected lane of warp 0 (shifting-window march)
    const uint32_t dS = ringS_a + (uint32_t)iS, dL = ringL_a + (uint32_t)iL, bar = full_a + 8u * (uint32_t)ib;
    const long long rr = (long long)r * g.ld, r1_ = (long long)max(r - 1, 0) * g.ld,
                    r2_ = (long long)max(r - 2, 0) * g.ld;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(row_bytes * (NS + NL))
                 : "memory");
    tma(dS + 8u * S_Q * RW, gq + rr, bar);
    tma(dL + 8u * L_V * RW, gv + rr, bar);
    tma(dL + 8u * L_SGC * RW, gsgc + rr, bar);
    tma(dL + 8u * L_SGV * RW, gsgv + rr, bar);
    if (MASK & 1) tma(dL + 8u * L_VM * RW, gvm + rr, bar);
    tma(dL + 8u * L_RGC * RW, grgc + rr, bar);
    tma(dS + 8u * S_SGU * RW, gsgu + r1_, bar);
    tma(dS + 8u * S_U * RW, gu + r2_, bar);
    if (MASK & 1) tma(dS + 8u * S_UM * RW, gum + r2_, bar);
  };
  // pointers to the staged rows of the row with (compile-time) phase k, for the const-slot marches
  const double* const eS = ringS + e;          // the thread's element in slot 0 of either ring
  const double* const eL = ringL + e;
  auto rowptrs = [&](auto kc) {
    constexpr int k = decltype(kc)::value;
    constexpr int kS = (k % DS) * SSLOT, kL0 = k * LSLOT, kL2 = ((k + DL - 2) % DL) * LSLOT,
                  kL3 = ((k + DL - 3) % DL) * LSLOT;
    RowPtrs R;
    R.q = eS + kS + S_Q * RW;
    R.u = eS + kS + S_U * RW;
    R.um = eS + kS + S_UM * RW;
    R.su1 = eS + kS + S_SGU * RW;
    R.v0 = eL + kL0 + L_V * RW;
    R.vm0 = eL + kL0 + L_VM * RW;
    R.sgv0 = eL + kL0 + L_SGV * RW;
    R.sgc0 = eL + kL0 + L_SGC * RW;
    R.rg0 = eL + kL0 + L_RGC * RW;
    R.sgc2 = eL + kL2 + L_SGC * RW;
    R.v3 = eL + kL3 + L_V * RW;
    R.vm3 = eL + kL3 + L_VM * RW;
    R.sgv3 = eL + kL3 + L_SGV * RW;
    R.sgc3 = eL + kL3 + L_SGC * RW;
    R.rg3 = eL + kL3 + L_RGC * RW;
    return R;
  };
  const long long ld8 = (long long)g.ld * 8;
  long long o0 = (long long)rfirst * ld8;    // byte offset of the next row to issue
  auto at = [](const double* base, long long off) {
    return reinterpret_cast<const double*>(reinterpret_cast<const char*>(base) + off);
  };
  constexpr int NCOPY = NS + NL;
  // the TMA copies of the row with phase k (elected lane of warp 0, after its fence.proxy.async); o0, o1, o2:
  // byte offsets of that row and of the rows staged one (sqrtg_pu) and two (u) rows late
  auto issue_k = [&](auto kc, long long o1, long long o2) {
    constexpr int k = decltype(kc)::value;
    const uint32_t dS = ringS_a + 8u * (uint32_t)((k % DS) * SSLOT), dL = ringL_a + 8u * (uint32_t)(k * LSLOT),
                   bar = full_a + 8u * (uint32_t)k;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(row_bytes * NCOPY)
                 : "memory");
    tma(dS + 8u * S_Q * RW, at(gq, o0), bar);
    tma(dL + 8u * L_V * RW, at(gv, o0), bar);
    tma(dL + 8u * L_SGC * RW, at(gsgc, o0), bar);
    tma(dL + 8u * L_SGV * RW, at(gsgv, o0), bar);
    tma(dL + 8u * L_RGC * RW, at(grgc, o0), bar);
    tma(dS + 8u * S_SGU * RW, at(gsgu, o1), bar);
    tma(dS + 8u * S_U * RW, at(gu, o2), bar);
    if (MASK & 1) {
      tma(dL + 8u * L_VM * RW, at(gvm, o0), bar);
      tma(dS + 8u * S_UM * RW, at(gum, o2), bar);
    }
  };
  double psum_cta = 0.0;
  if constexpr (PAIR) {
    // ---- two rows per pair of barriers.  Step s marches rows (ra, rb) = rfirst + 2s, + 1:
    //   phase 1 of ra then rb | barrier A | the four y-fluxes | barrier B | TMA issue of step s+2 |
    //   phase 3 of ra then rb.
    // Half the barriers per row and two independent rows of FP64 work between them; costs ~16
    // more registers (two XEdge sets, two flux pairs across barrier B) -> 3 CTAs/SM.
    // After barrier B of step s nobody reads the ring slots of step s (short ring) and of rows
    // ra-4, ra-3 (long ring) any more: exactly the slots of step s+2, so its copies start there
    // (2.7 rows ahead of their first use).  Row phase k = (row - rfirst) % 8 is compile time.
    static_assert(DS == 4 && DL == WMAX, "two-row march: rings of 4 and 8 slots, windows of 8 registers");
    constexpr int W = WMAX;
    double* sXb = sX + 4 * RW;                 // second row of a pair: Qx, F, G, C
    double *sFb = sXb + RW, *sGb = sFb + RW, *sCb = sGb + RW;
    const int rlastp = rlast + ((rlast - rfirst + 1) & 1);   // even number of rows (row r1+3 <= P-1 exists; its output is dropped)
    if (warp_u == 0) {                         // steps 0 and 1: rows rfirst .. rfirst+3 (rfirst >= 1: only row -1 is clamped)
      const bool el = elect_one();
      if (el) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue_k(IC<0>{}, o0 - ld8, (long long)max(rfirst - 2, 0) * ld8);
      }
      o0 += ld8;
      if (el) issue_k(IC<1>{}, o0 - ld8, o0 - 2 * ld8);
      o0 += ld8;
      if (rfirst + 2 <= rlastp) {
        if (el) issue_k(IC<2>{}, o0 - ld8, o0 - 2 * ld8);
        o0 += ld8;
        if (el) issue_k(IC<3>{}, o0 - ld8, o0 - 2 * ld8);
        o0 += ld8;
      }
    }
    double* __restrict__ QN = a.qn + (long long)p * g.ps + PYCS_JOFF + min(j, g.P - 1) + (long long)r0 * g.ld;
    Lane L;
    lane_init(L);
    uint32_t parb = 0;
    // one step = two marched rows; false once the chunk is finished
    auto step = [&](auto kc, int ra) -> bool {
      constexpr int k = decltype(kc)::value;         // phase of row ra (even); rb has k + 1
      if (ra > rlastp) return false;
      const int rb = ra + 1;
      while (!mbar_try_wait(&full[k], parb)) {}
      while (!mbar_try_wait(&full[k + 1], parb)) {}
      const RowPtrs Ra = rowptrs(IC<k>{}), Rb = rowptrs(IC<k + 1>{});
      double qa[1] = {Ra.q[0]}, qb[1] = {Rb.q[0]};
      if (a.apply_corr && jint) {                    // pending MF-PR term on the own interior cells
        if (ra >= g.lo && ra < g.hi) {
          qa[0] = fma(Ra.sgc0[0], corr, qa[0]);
          ringS[(k % DS) * SSLOT + S_Q * RW + e] = qa[0];
        }
        if (rb >= g.lo && rb < g.hi) {
          qb[0] = fma(Rb.sgc0[0], corr, qb[0]);
          ringS[((k + 1) % DS) * SSLOT + S_Q * RW + e] = qb[0];
        }
      }
      // ---------------- phase 1: own column, both rows
      XEdge Xa, Xb;
      double qxa[1], qxb[1];
      phase_x_inner<RECON, SPLIT, MASK, k, W>(L, Xa, Ra, qa, cdx, qxa);
      phase_x_inner<RECON, SPLIT, MASK, k + 1, W>(L, Xb, Rb, qb, cdx, qxb);
      sX[e] = qxa[0];
      sXb[e] = qxb[0];
      __syncthreads();                                   // barrier A
      // ---------------- phase 2: y-fluxes at edge j: inner on Q rows ra, rb, outer on Qx rows ra-3, rb-3
      double Fa[1], Ga[1], CFa[1] = {0.0}, CGa[1], Fb[1], Gb[1], CFb[1] = {0.0}, CGb[1];
      yflux_pair<RECON, SPLIT, MASK, k>(Ra.v0, Ra.vm0, Ra.sgv0, Ra.sgc0, Ra.q, cdy, Fa, CFa);
      yflux_pair<RECON, SPLIT, MASK, k>(Rb.v0, Rb.vm0, Rb.sgv0, Rb.sgc0, Rb.q, cdy, Fb, CFb);
      yflux_pair<RECON, SPLIT, MASK, k>(Ra.v3, Ra.vm3, Ra.sgv3, Ra.sgc3, sX + e, cdy, Ga, CGa);
      yflux_pair<RECON, SPLIT, MASK, k>(Rb.v3, Rb.vm3, Rb.sgv3, Rb.sgc3, sXb + e, cdy, Gb, CGb);
      sF[e] = Fa[0];
      sG[e] = Ga[0];
      sFb[e] = Fb[0];
      sGb[e] = Gb[0];
      if (SPLIT != 1) { sC[e] = CFa[0]; sCb[e] = CFb[0]; }
      __syncthreads();                                   // barrier B
      if (warp_u == 0 && ra + 4 <= rlastp) {             // the TMA copies of step s+2 (rows ra+4, ra+5)
        if (elect_one()) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          issue_k(IC<(k + 4) % DL>{}, o0 - ld8, o0 - 2 * ld8);
        }
        o0 += ld8;
        if (elect_one()) issue_k(IC<(k + 5) % DL>{}, o0 - ld8, o0 - 2 * ld8);
        o0 += ld8;
      }
      // ---------------- phase 3: Qy rows ra, rb, outer x-fluxes, output rows ra-3, rb-3
      double Fn[1], Gn[1], CFn[1] = {0.0}, out[1], sdiv[1];
      Fn[0] = sF[e + 1];
      Gn[0] = sG[e + 1];
      if (SPLIT != 1) CFn[0] = sC[e + 1];
      phase_x_outer<RECON, SPLIT, k, W>(L, Xa, Ra, Fa, Fn, Ga, Gn, CFa, CFn, out, sdiv);
      if (ra >= r0 + 3) {
        if (out_lane) {
          *QN = out[0];
          L.psum += sdiv[0];
        }
        QN += g.ld;
      }
      Fn[0] = sFb[e + 1];
      Gn[0] = sGb[e + 1];
      if (SPLIT != 1) CFn[0] = sCb[e + 1];
      phase_x_outer<RECON, SPLIT, k + 1, W>(L, Xb, Rb, Fb, Fn, Gb, Gn, CFb, CFn, out, sdiv);
      if (rb >= r0 + 3 && rb <= rlast) {
        if (out_lane) {
          *QN = out[0];
          L.psum += sdiv[0];
        }
        QN += g.ld;
      }
      return true;
    };
    for (int rb0 = rfirst;; rb0 += DL, parb ^= 1u) {
      if (!step(IC<0>{}, rb0)) break;
      if (!step(IC<2>{}, rb0 + 2)) break;
      if (!step(IC<4>{}, rb0 + 4)) break;
      if (!step(IC<6>{}, rb0 + 6)) break;
    }
    psum_cta = L.psum;
  } else if constexpr (CS) {
    static_assert(PF == 2, "const-slot march: rings of 3 and 6 rows, windows of 6 registers");
    static_assert(DL == WLEN && DL % DS == 0, "const-slot march: one period for rings and windows");
    if (warp_u == 0) {                         // rows rfirst, rfirst + 1 (rfirst >= 1: only row -1 is clamped)
      if (elect_one()) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue_k(IC<0>{}, o0 - ld8, (long long)max(rfirst - 2, 0) * ld8);
      }
      o0 += ld8;
      if (rfirst + 1 <= rlast) {
        if (elect_one()) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          issue_k(IC<1>{}, o0 - ld8, o0 - 2 * ld8);
        }
        o0 += ld8;
      }
    }
    double* __restrict__ QN = a.qn + (long long)p * g.ps + PYCS_JOFF + min(j, g.P - 1) + (long long)r0 * g.ld;
    Lane L;
    lane_init(L);
    uint32_t parb = 0;
    // one marched row; false once the chunk is finished
    auto row = [&](auto kc, int r) -> bool {
      constexpr int k = decltype(kc)::value;
      if (r > rlast) return false;
      constexpr int kS = (k % DS) * SSLOT;
      constexpr int WO = PB ? (k & 1) * 4 * RW : 0;      // work rows of this row's parity
      while (!mbar_try_wait(&full[k], parb)) {}
      const RowPtrs R = rowptrs(IC<k>{});
      double qnew[1] = {R.q[0]};
      if (a.apply_corr && jint && r >= g.lo && r < g.hi) {
        qnew[0] = fma(R.sgc0[0], corr, qnew[0]);
        ringS[kS + S_Q * RW + e] = qnew[0];
      }
      // ---------------- phase 1: own column
      XEdge X;
      double qx[1];
      phase_x_inner<RECON, SPLIT, MASK, k>(L, X, R, qnew, cdx, qx);
      sX[WO + e] = qx[0];
      __syncthreads();                                   // barrier A
      if (warp_u == 0 && r + PF <= rlast) {              // warp 0 issues the TMA copies of row r+PF
        if (elect_one()) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          issue_k(IC<(k + PF) % DL>{}, o0 - ld8, o0 - 2 * ld8);
        }
        o0 += ld8;
      }
      // ---------------- phase 2: y-fluxes at edge j: inner on Q row r, outer on Qx row r-3
      double F[1], G[1], CF[1] = {0.0}, CG[1];
      yflux_pair<RECON, SPLIT, MASK, k>(R.v0, R.vm0, R.sgv0, R.sgc0, R.q, cdy, F, CF);
      yflux_pair<RECON, SPLIT, MASK, k>(R.v3, R.vm3, R.sgv3, R.sgc3, sX + WO + e, cdy, G, CG);
      sF[WO + e] = F[0];
      sG[WO + e] = G[0];
      if (SPLIT != 1) sC[WO + e] = CF[0];
      if constexpr (PB) {
        // named barrier w+1 pairs warp w (consumer: bar.sync) with warp w+1 (producer: bar.arrive)
        if (warp_u > 0) asm volatile("bar.arrive %0, 64;" ::"r"(warp_u) : "memory");
        if (warp_u < NWARP - 1) asm volatile("bar.sync %0, 64;" ::"r"(warp_u + 1) : "memory");
        else __syncwarp();
      } else {
        __syncthreads();                                 // barrier B
      }
      // ---------------- phase 3: Qy row r, outer x-flux on Qy, output row r-3
      double Fn[1], Gn[1], CFn[1] = {0.0}, out[1], sdiv[1];
      Fn[0] = sF[WO + e + 1];
      Gn[0] = sG[WO + e + 1];
      if (SPLIT != 1) CFn[0] = sC[WO + e + 1];
      phase_x_outer<RECON, SPLIT, k>(L, X, R, F, Fn, G, Gn, CF, CFn, out, sdiv);
      if (r >= r0 + 3) {
        if (out_lane) {
          *QN = out[0];
          L.psum += sdiv[0];
        }
        QN += g.ld;
      }
      return true;
    };
    for (int rb = rfirst;; rb += DL, parb ^= 1u) {
      if (!row(IC<0>{}, rb)) break;
      if (!row(IC<1>{}, rb + 1)) break;
      if (!row(IC<2>{}, rb + 2)) break;
      if (!row(IC<3>{}, rb + 3)) break;
      if (!row(IC<4>{}, rb + 4)) break;
      if (!row(IC<5>{}, rb + 5)) break;
    }
    psum_cta = L.psum;
  } else {
    if (warp_u == 0) {
      for (int r = rfirst; r < rfirst + PF && r <= rlast; ++r) {
        if (elect_one()) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          issue(r);
        }
        iS = (iS + 8 * SSLOT == 8 * DS * SSLOT) ? 0 : iS + 8 * SSLOT;
        iL = (iL + 8 * LSLOT == 8 * DL * LSLOT) ? 0 : iL + 8 * LSLOT;
        ib = (ib + 1 == DL) ? 0 : ib + 1;
      }
    }

    double* __restrict__ QN = a.qn + (long long)p * g.ps + PYCS_JOFF + min(j, g.P - 1) + (long long)r0 * g.ld;
    Lane L;
    lane_init(L);
    int oS = 0;
    int oL0 = 0, oL1 = (DL - 1) * LSLOT, oL2 = (DL - 2) * LSLOT, oL3 = (DL - 3) * LSLOT;
    int sb = 0;
    uint32_t parb = 0;
  #pragma unroll((MINB >= 10 && MG != 1) ? 5 : 1)   // MINB 1x / 2x: unrolled by the window length
    for (int r = rfirst; r <= rlast; ++r) {
      while (!mbar_try_wait(&full[sb], parb)) {}
      RowPtrs R;
      R.q = ringS + oS + S_Q * RW + e;
      R.u = ringS + oS + S_U * RW + e;
      R.um = ringS + oS + S_UM * RW + e;
      R.su1 = ringS + oS + S_SGU * RW + e;
      R.v0 = ringL + oL0 + L_V * RW + e;
      R.vm0 = ringL + oL0 + L_VM * RW + e;
      R.sgv0 = ringL + oL0 + L_SGV * RW + e;
      R.sgc0 = ringL + oL0 + L_SGC * RW + e;
      R.rg0 = ringL + oL0 + L_RGC * RW + e;
      R.sgc2 = ringL + oL2 + L_SGC * RW + e;
      R.v3 = ringL + oL3 + L_V * RW + e;
      R.vm3 = ringL + oL3 + L_VM * RW + e;
      R.sgv3 = ringL + oL3 + L_SGV * RW + e;
      R.sgc3 = ringL + oL3 + L_SGC * RW + e;
      R.rg3 = ringL + oL3 + L_RGC * RW + e;
      // pending MF-PR term on the own interior cell of the new row (neighbours read it after barrier A)
      double qnew[1] = {R.q[0]};
      if (a.apply_corr && jint && r >= g.lo && r < g.hi) {
        qnew[0] = fma(R.sgc0[0], corr, qnew[0]);
        ringS[oS + S_Q * RW + e] = qnew[0];
      }
      // ---------------- phase 1: own column
      XEdge X;
      double qx[1];
      phase_x_inner<RECON, SPLIT, MASK>(L, X, R, qnew, cdx, qx);
      sX[e] = qx[0];
      __syncthreads();                                   // barrier A
      if (warp_u == 0 && r + PF <= rlast) {              // warp 0 issues the TMA copies of row r+PF
        if (elect_one()) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          issue(r + PF);
        }
        iS = (iS + 8 * SSLOT == 8 * DS * SSLOT) ? 0 : iS + 8 * SSLOT;
        iL = (iL + 8 * LSLOT == 8 * DL * LSLOT) ? 0 : iL + 8 * LSLOT;
        ib = (ib + 1 == DL) ? 0 : ib + 1;
      }
      // ---------------- phase 2: y-fluxes at edge j: inner on Q row r, outer on Qx row r-3
      double F[1], G[1], CF[1] = {0.0}, CG[1];
      yflux_pair<RECON, SPLIT, MASK>(R.v0, R.vm0, R.sgv0, R.sgc0, R.q, cdy, F, CF);
      yflux_pair<RECON, SPLIT, MASK>(R.v3, R.vm3, R.sgv3, R.sgc3, sX + e, cdy, G, CG);
      sF[e] = F[0];
      sG[e] = G[0];
      if (SPLIT != 1) sC[e] = CF[0];
      __syncthreads();                                   // barrier B
      // ---------------- phase 3: Qy row r, outer x-flux on Qy, output row r-3
      double Fn[1], Gn[1], CFn[1] = {0.0}, out[1], sdiv[1];
      Fn[0] = sF[e + 1];
      Gn[0] = sG[e + 1];
      if (SPLIT != 1) CFn[0] = sC[e + 1];
      phase_x_outer<RECON, SPLIT>(L, X, R, F, Fn, G, Gn, CF, CFn, out, sdiv);
      if (r >= r0 + 3) {
        if (out_lane) {
          *QN = out[0];
          L.psum += sdiv[0];
          if (MG && mgw > 1) {
            // fused exchange: the peers need this cell if it lies in a 4-wide boundary strip of the
            // panel (sources of their ghost fill) or in the 3 rows next to a neighbour's slab
            const int ro = r - 3;
            const long long off = QN - a.qn;
            const bool strip = j < g.lo + PYCS_NG || j >= g.hi - PYCS_NG || ro < g.lo + PYCS_NG || ro >= g.hi - PYCS_NG;
            if (strip) {
              for (int d = 0; d < mgw; ++d)
                if (d != a.mg.rank) a.mg.peer_qn[d][off] = out[0];
            } else {
              if (ro < a.row_lo + 3 && a.mg.rank > 0) a.mg.peer_qn[a.mg.rank - 1][off] = out[0];
              if (ro >= a.row_hi - 3 && a.mg.rank < mgw - 1) a.mg.peer_qn[a.mg.rank + 1][off] = out[0];
            }
          }
        }
        QN += g.ld;
      }
      oS = (oS + SSLOT == DS * SSLOT) ? 0 : oS + SSLOT;
      oL3 = oL2; oL2 = oL1; oL1 = oL0;
      oL0 = (oL0 + LSLOT == DL * LSLOT) ? 0 : oL0 + LSLOT;
      if (++sb == DL) { sb = 0; parb ^= 1u; }
    }
    psum_cta = L.psum;
  }
  // per-CTA partial of sum(pxdF + pydF) over its outputs (MF-PR), fixed order
  __syncthreads();
  double v = psum_cta;
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((tid & 31) == 0) sF[tid >> 5] = v;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < TB / 32; ++w) t += sF[w];
    a.part[(MG == 2) ? a.blk_map[blockIdx.x] : (int)blockIdx.x] = t;   // re-read: not worth a register across the march
    sF[40] = fused_last_writer(a.counter, (MG == 2) ? (unsigned)a.nblk_total : gridDim.x, mgw > 1) ? 1.0 : 0.0;
  }
  __syncthreads();
  if (sF[40] != 0.0 && tid < 32) {            // last CTA of the launch: total in a fixed order
    double tot = fused_warp_sum(a.part, (MG == 2) ? a.nblk_total : (int)gridDim.x, tid);
    if (tid == 0) {
      *a.sum_out = tot;
      *a.counter = 0u;
    }
    if (MG == 1 && mgw > 1) {                 // publish this rank's sum, then raise its flag everywhere
      tot = __shfl_sync(0xffffffffu, tot, 0);
      if (tid < mgw) {
        MgSync* sy = a.mg.peer_sync[tid];
        sy->psum[a.mg.parity][a.mg.rank] = tot;
        __threadfence_system();
        *((volatile long long*)&sy->flag[a.mg.rank]) = a.mg.epoch;
      }
    }
  }
}

template <int TB, int PF, int MASK, int MINB>
constexpr size_t smem_bytes() {
  constexpr bool PAIR = (MINB >= 50) && (MINB < 60);
  constexpr int DS = PAIR ? 4 : PF + 1, DL = PAIR ? 8 : PF + 4, NWORK = (MINB >= 50) ? 8 : 4;
  return sizeof(double) * (size_t)(TB + 6) * (DS * ((MASK & 1) ? 4 : 3) + DL * ((MASK & 1) ? 5 : 4) + NWORK) +
         sizeof(uint64_t) * DL + 16;
}

template <int TB, int RECON, int SPLIT, int MASK, int PF, int MINB, int MG>
cudaError_t launch_mg(const FusedArgs& a, int nblocks, cudaStream_t st, int* resident) {
  static bool configured = false;
  const size_t smem = smem_bytes<TB, PF, MASK, (MG == 1) ? 0 : MINB>();
  auto kern = fused2b_kernel<TB, RECON, SPLIT, MASK, PF, MINB, MG>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (resident) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(resident, kern, TB, smem);
  if (a.pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nblocks);
    cfg.blockDim = dim3(TB);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, a);
  }
  kern<<<nblocks, TB, smem, st>>>(a);
  return cudaSuccess;
}

template <int TB, int RECON, int SPLIT, int MASK, int PF, int MINB>
cudaError_t launch_one(const FusedArgs& a, int nblocks, cudaStream_t st, int* resident) {
  // the multi-GPU build is never unrolled: one instantiation per register cap
  if (a.mg.world > 1) return launch_mg<TB, RECON, SPLIT, MASK, PF, MINB % 10, 1>(a, nblocks, st, resident);
  if (a.blk_map) {               // split launches: built for the default march only
    if constexpr (TB == 160 && PF == 2 && MINB == F2B_CS_MINB)
      return launch_mg<TB, RECON, SPLIT, MASK, PF, MINB, 2>(a, nblocks, st, resident);
    else
      return cudaErrorInvalidValue;
  }
  return launch_mg<TB, RECON, SPLIT, MASK, PF, MINB, 0>(a, nblocks, st, resident);
}

template <int TB, int RECON, int SPLIT, int PF, int MINB>
cudaError_t launch_mask(const FusedArgs& a, int mask, int nblocks, cudaStream_t st, int* resident) {
  if (mask == 1) return launch_one<TB, RECON, SPLIT, 1, PF, MINB>(a, nblocks, st, resident);
  if (mask == 2) return launch_one<TB, RECON, SPLIT, 2, PF, MINB>(a, nblocks, st, resident);
  return launch_one<TB, RECON, SPLIT, 0, PF, MINB>(a, nblocks, st, resident);
}

#define F2B_DEFAULT_TB 160
#define F2B_DEFAULT_PF 2
#define F2B_DEFAULT_MINB 4
cudaError_t dispatch(const FusedArgs& a, int recon, int split, int mask, int tb, int pf, int minb, int nblocks,
                     cudaStream_t st, int* resident) {
  if (recon == 3 && split == 1) {
#define TUNE(T, P, M) \
  if (tb == T && pf == P && minb == M) return launch_mask<T, 3, 1, P, M>(a, mask, nblocks, st, resident)
    // (threads, rows in flight, MINB): the default, the earlier marches it is tested against and the
    // candidates of the next sweep; measured-and-dropped points are in profiles/r1_sweep_v2b*.log
    TUNE(160, 2, 14); TUNE(160, 2, 4); TUNE(160, 2, 34); TUNE(128, 2, 35); TUNE(160, 2, 53); TUNE(128, 2, 54); TUNE(160, 2, 64); TUNE(160, 2, 33);
#undef TUNE
    return cudaErrorInvalidValue;
  }
  // the other schemes: the plain march (MINB 4) and the const-slot march (MINB 34)
  if (tb != F2B_DEFAULT_TB || pf != F2B_DEFAULT_PF || (minb != F2B_DEFAULT_MINB && minb != F2B_CS_MINB))
    return cudaErrorInvalidValue;
#define CASE(R, S) \
  if (recon == R && split == S) \
    return minb == F2B_CS_MINB \
               ? launch_mask<F2B_DEFAULT_TB, R, S, F2B_DEFAULT_PF, F2B_CS_MINB>(a, mask, nblocks, st, resident) \
               : launch_mask<F2B_DEFAULT_TB, R, S, F2B_DEFAULT_PF, F2B_DEFAULT_MINB>(a, mask, nblocks, st, resident)
  CASE(3, 2); CASE(3, 3); CASE(1, 1); CASE(1, 2); CASE(1, 3);
#undef CASE
  return cudaErrorInvalidValue;
}

}  // namespace

int pycs_split_sets(int nstrips, int nchunks, int* interior, int* boundary) {
  if (nstrips < 3 || nchunks < 3) return 0;
  int ni = 0, nb = 0;
  for (int chunk = 0; chunk < nchunks; ++chunk)
    for (int strip = 0; strip < nstrips; ++strip) {
      const bool in = strip >= 1 && strip <= nstrips - 2 && chunk >= 1 && chunk <= nchunks - 2;
      for (int p = 0; p < 6; ++p) {
        const int b = (chunk * nstrips + strip) * 6 + p;
        if (in) interior[ni++] = b;
        else boundary[nb++] = b;
      }
    }
  return ni;
}

bool pycs_fused2b_has(int recon, int split, int tb, int pf, int minb) {
  if (recon != 1 && recon != 3) return false;
  if (recon == 3 && split == 1) {
    const int t[][3] = {{160, 2, 14}, {160, 2, 4}, {160, 2, 34}, {128, 2, 35}, {160, 2, 53}, {128, 2, 54}, {160, 2, 64}, {160, 2, 33}};
    for (auto& x : t)
      if (x[0] == tb && x[1] == pf && x[2] == minb) return true;
    return false;
  }
  return tb == F2B_DEFAULT_TB && pf == F2B_DEFAULT_PF && (minb == F2B_DEFAULT_MINB || minb == F2B_CS_MINB);
}

cudaError_t pycs_launch_fused2b(const FusedArgs& a, int recon, int split, int mask, int tb, int pf, int minb,
                                int nblocks, cudaStream_t st) {
  return dispatch(a, recon, split, mask, tb, pf, minb, nblocks, st, nullptr);
}

int pycs_fused2b_resident(int recon, int split, int mask, int tb, int pf, int minb) {
  FusedArgs a{};
  int n = 0;
  if (dispatch(a, recon, split, mask, tb, pf, minb, 0, nullptr, &n) != cudaSuccess) return -1;
  return n;
}
