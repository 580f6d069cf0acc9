// Fused advection step, v2b -- the production step kernel.
//
// One launch = divergence + Q update of src/discrete_operators.py:18-101 and
// src/advection_timestep.py:43 for the duo-grid schemes (ET-DG): inner x/y PPM fluxes, the
// splitting update (Qx, Qy), outer fluxes on (Qy, Qx), the metric-weighted divergence and
// Q -= dt * div.  Per cell it reads Q, the two winds and sqrt(g) (40 B algorithmic) and writes Q;
// q_L / q_R / fluxes / Qx / Qy never leave the SM.  Arithmetic: fused3_core.cuh (weight form of the
// CW84 flux, shared verbatim with the host emulator of tests/emul).
//
// Decomposition: a CTA = (panel, strip of TB-6 columns, chunk of rows); a thread owns one column
// and marches the rows.  The x-sweeps (stencil across rows) live in registers as circular five-row
// windows; the y-sweeps (stencil along the row) go through shared memory.  Per marched row:
//   phase 1 (own column: new row, inner x-flux, Qx) | barrier A | phase 2 (y-fluxes of Q row r and
//   of Qx row r-3 at the thread's edge) | barrier B | phase 3 (Qy, outer x-flux, output row r-3).
// Rows are staged by one elected lane of warp 0 with cp.async.bulk (TMA 1-D row copies,
// mbarrier complete_tx), two rows ahead (the first two before the CTA looks at the control block), into a two-level ring: Q, u, sqrtg_pu are needed by one
// march step (3 slots); v, sqrtg_pv, sqrtg_pc, 1/sqrtg_pc by steps r..r+3 (6 slots).  The march
// runs in groups of six rows -- the common period of both rings and of the register windows --
// so that every ring slot and every window register is a compile-time constant
// (profiles/r1_sass_static.md).
//
// What changes from step to step (separable-wind time factor, MF-PR sums, multi-GPU epochs) is read
// from the device-side control block (StepCtl, fused_args.cuh), not from kernel parameters: the
// launches of a step are replayable from a CUDA graph (stepper.cu).
//
// GH = 0: the CTA's staged cells carry the pending MF-PR term already on their ghost cells (serial
//         path: the ghost fill folded it in) or have no ghost cells at all (interior CTAs of a
//         split step);
// GH = 1: boundary CTAs of a split step: the ghost fill ran early and raw, the projection term of
//         the ghost cells (corr * ghost(sqrtg)) is added here as the rows are loaded.
// GH = 2: one-kernel step: as GH = 1, and every CTA first fills the ghost cells its own march reads (Lagrange fill,
//         ghost_core.cuh; on several GPUs after waiting for the peers' exchange flags) -- no ghost-fill kernel.
#include <cstdlib>
#include "fused_args.cuh"
#include "mgpu.cuh"
#include "ghost_core.cuh"
// PPM-PL07 edge-value coefficients as constant-bank operands
__constant__ double f2b_ppm_coef[5] = {2.0 / 60.0, -13.0 / 60.0, 47.0 / 60.0, 27.0 / 60.0, -3.0 / 60.0};
#define F3_COEF_BANK f2b_ppm_coef
#define F3_NAMESPACE f1
#define F3_NC 1
#define F3_CSTEP 0
#include "fused3_core.cuh"
#undef F3_COEF_BANK

#define F2B_TB 160         // threads per CTA: 154 output columns; 96 registers, 49 KB -> 4 CTAs / SM
#define F2B_PF 2           // rows in flight ahead of the march
#define F2B_MINB 4         // register cap as CTAs per SM
// Variant bits of the march (template parameter VAR; PYCS_VARIANT selects among the instantiated ones, experiment knob):
//   1  the TMA copies of row r+3 are issued after barrier B of row r (default form: row r+2 after barrier A) -- the
//      ring slots are free at that point already, so the copies get ~0.7 row times more lead at no shared memory
//   2  L2 prefetch (cp.async.bulk.prefetch.L2) of the HBM-sourced rows (Q, u, v) 4 rows beyond the row being issued
//   4  the same, 8 rows beyond
//   8  y-stencils take the thread's own cell from its register (4 LDS per flux instead of 5)
//  16  x-edge weights: one dynamically addressed load of sqrtg at the upwind centre instead of both candidates
//  32  the CFL numbers of the two y-fluxes (wind row x dt/dy) are formed before barrier A
// Measured at N=1536 (profiles/r2_march_variants.log): default 0.1640 ms; 1: 0.1659; 2 / 4 (+1): 0.1659-0.1664;
// 8: 0.1680; 16: 0.1660; 32: 0.1637 (the default since).  A probe with barrier B left out (wrong results) ran at
// 0.1603 ms: the barriers cost 2 %, and the longer lead of the copies / the L2 prefetch buy nothing -- the march is
// bound by per-warp dependency stalls at 20 warps per SM, not by the staging.
// To time a variant: -DF2B_VARIANTS="F2B_V(1); F2B_V(8);" instantiates it for the par-default scheme.
#define F2B_VAR_DEFAULT 32

namespace {

using namespace f1;

template <int K> struct IC { static constexpr int value = K; };   // compile-time row phase
template <bool B> struct BC { static constexpr bool value = B; };

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


// ---- GH = 2: the ghost cells of the rectangle rows [Ra, Rb) x columns [Ca, Cb) of panel p, written to Q (raw
// Lagrange fill, same operations in the same order as ghost_core.cuh).  This runs once per CTA at a panel edge and
// is pure latency, so it is written for memory-level parallelism in little code: ONE loop over the CTA's ghost
// cells, two cells per trip, and per cell the stencil position first, then all source cells and weights at once
// (up to eight, predicated), then the sum.  History (N = 1536, one GPU, one-kernel step against the serial step's
// 0.170 ms): dg_ghost_cell inlined in two loops -- a dependent load per stencil point, 2 600 instructions: 0.180;
// four cells per trip with the corner path inlined four times -- 14 500 instructions: 0.192; the plain loop in a
// non-inlined function that reads its arguments through a generic pointer: 0.190 (profiles/r2_mgpu_onekernel.md).
__device__ __noinline__ double ghost_corner_nl(Geo g, const HaloMaps* maps, const double* q, const int* kmin,
                                               const double* wE, int order, int p, int i, int j) {
  return dg_corner_value(g, *maps, q, kmin, wE, order, p, i >= g.hi ? SIDE_E : SIDE_W, i >= g.hi ? i - g.hi : i, j);
}
struct GhostCell { int i, j, s, gl, k, ge, km; bool ok, corner; };
__device__ __forceinline__ void ghost_prologue(const FusedArgs& a, int p, int Ra, int Rb, int Ca, int Cb, int tid, int tb) {
  const Geo& g = a.g;
  double* __restrict__ qw = const_cast<double*>(a.q);
  const int nL = max(0, g.lo - Ca), nR = max(0, Cb - g.hi), nT = max(0, g.lo - Ra), nB = max(0, Rb - g.hi);
  const int ncg = nL + nR, ncol = ncg * (Rb - Ra);       // ghost columns over all rows of the rectangle (corners included)
  const int ja = max(Ca, g.lo), w = min(Cb, g.hi) - ja, nrg = nT + nB;   // ghost rows over the interior columns
  const int n = ncol + nrg * w;
  auto locate = [&](int t) {
    GhostCell c;
    c.ok = t < n;
    c.i = c.j = c.s = c.gl = c.k = c.ge = c.km = 0;
    c.corner = false;
    if (c.ok) {
      if (t < ncol) {
        const int cc = t % ncg;
        c.i = Ra + t / ncg;
        c.j = cc < nL ? Ca + cc : g.hi + (cc - nL);
      } else {
        const int u = t - ncol, rr = u / w;
        c.j = ja + u % w;
        c.i = rr < nT ? Ra + rr : g.hi + (rr - nT);
      }
      const bool ii = c.i >= g.lo && c.i < g.hi, jj = c.j >= g.lo && c.j < g.hi;
      c.corner = !ii && !jj;
      if (ii) { c.s = c.j >= g.hi ? SIDE_N : SIDE_S; c.gl = c.j >= g.hi ? c.j - g.hi : c.j; c.k = c.i; }
      else { c.s = c.i >= g.hi ? SIDE_E : SIDE_W; c.gl = c.i >= g.hi ? c.i - g.hi : c.i; c.k = c.j; }
      c.ge = (c.s == SIDE_E || c.s == SIDE_N) ? c.gl : PYCS_NG - 1 - c.gl;
      if (!c.corner) c.km = a.gf_kmin[c.ge * g.P + c.k];
    }
    return c;
  };
  auto gather = [&](const GhostCell& c, double (&v)[8], double (&wt)[8]) {
    const SideMap& m = a.gf_maps.m[p][c.s];
    // source cell of stencil point l: halo_src(gl, km + l) on the E / W sides, (km + l, gl) on N / S: affine in l
    const int a0 = c.s < 2 ? c.gl : c.km, b0 = c.s < 2 ? c.km : c.gl, da = c.s < 2 ? 0 : 1, db = 1 - da;
    const long long base = gidx(g, m.nb, m.ci + m.ai * a0 + m.bi * b0, m.cj + m.aj * a0 + m.bj * b0);
    const long long step = (long long)(m.ai * da + m.bi * db) * g.ld + (m.aj * da + m.bj * db);
    const double* __restrict__ wp = a.gf_w + ((long long)c.ge * g.P + c.k) * a.gf_order;
#pragma unroll
    for (int l = 0; l < 8; ++l) {
      const bool on = c.ok && !c.corner && l < a.gf_order;
      v[l] = on ? a.q[base + l * step] : 0.0;
      wt[l] = on ? wp[l] : 0.0;
    }
  };
  auto finish = [&](const GhostCell& c, const double (&v)[8], const double (&wt)[8]) {
    if (!c.ok) return;
    double acc = 0.0;
    if (c.corner) {
      acc = ghost_corner_nl(g, &a.gf_maps, a.q, a.gf_kmin, a.gf_w, a.gf_order, p, c.i, c.j);
    } else {
#pragma unroll
      for (int l = 0; l < 8; ++l)
        if (l < a.gf_order) acc = __dadd_rn(acc, __dmul_rn(v[l], wt[l]));
    }
    qw[gidx(g, p, c.i, c.j)] = acc;
  };
#pragma unroll 1
  for (int t = tid; t < n; t += 2 * tb) {
    const GhostCell c0 = locate(t), c1 = locate(t + tb);
    double v0[8], w0[8], v1[8], w1[8];
    gather(c0, v0, w0);
    gather(c1, v1, w1);
    finish(c0, v0, w0);
    finish(c1, v1, w1);
  }
}

template <int TB, int RECON, int SPLIT, int MASK, int GH, int VAR = 0>
__global__ void __launch_bounds__(TB, F2B_MINB) fused2b_kernel(const __grid_constant__ FusedArgs a) {
  constexpr int PF = F2B_PF;
  constexpr int RW = TB + 6;                 // staged row: columns jbase-6 .. jbase+TB-1
  constexpr int NS = (MASK & 1) ? 4 : 3, NL = (MASK & 1) ? 5 : 4;
  constexpr int DS = PF + 1, DL = PF + 4;
  constexpr int NWORK = 4;                   // work rows: Qx, inner / outer y-flux, sqrtg_pv*cy
  constexpr int SSLOT = NS * RW, LSLOT = NL * RW;
  static_assert(PF == 2 && DL == WLEN && DL % DS == 0, "const-slot march: rings of 3 and 6 rows, windows of 6 registers");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* ringS = reinterpret_cast<double*>(smem_raw);           // [DS][NS][RW]
  double* ringL = ringS + DS * SSLOT;                            // [DL][NL][RW]
  double* sX = ringL + DL * LSLOT;                               // Qx row r-3
  double* sF = sX + RW;                                          // inner y-flux, row r
  double* sG = sF + RW;                                          // outer y-flux, row r-3
  double* sC = sG + RW;                                          // sqrtg_pv*cy (SPLIT != 1)
  uint64_t* full = reinterpret_cast<uint64_t*>(sX + NWORK * RW); // [DL]

  const Geo& g = a.g;
  int cta, p, strip, r0, r1;
  if (a.cta_tab) {                           // split step: the CTA's rows and strip come from the table
    cta = a.cta_off + (int)blockIdx.x;
    const int4 d = a.cta_tab[cta];
    r0 = d.x; r1 = d.y; strip = d.z; p = d.w;
  } else {                                   // uniform grid: strips x chunks x 6, panel fastest
    cta = (int)blockIdx.x;
    int b = cta;
    p = b % 6;
    b /= 6;
    strip = b % a.nstrips;
    const int chunk = b / a.nstrips;
    r0 = a.row_lo + chunk * a.rows_per_chunk;
    r1 = min(r0 + a.rows_per_chunk, a.row_hi);
  }
  const int tid = threadIdx.x;
  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp index, provably warp-uniform
  const int e = tid + 3;                     // own element in a staged row (the y-stencil reaches e-3 .. e+2)
  const int jbase = g.lo + strip * a.wcols;
  const int jend = min(jbase + a.wcols, g.hi);
  const int j = jbase - 3 + tid;
  const int rfirst = r0 - 3, rlast = r1 + 2;
  const bool out_lane = (tid >= 3) && (j < jend);
  const bool jint = (j >= g.lo) && (j < g.hi);
  const int c0 = jbase - 6;                  // 16-byte aligned: JOFF and wcols are even
  const int len = min(RW, g.ld - PYCS_JOFF - c0) & ~1;
  const uint32_t row_bytes = (uint32_t)len * 8u;

  pdl_trigger();                             // the next kernel of the stream may be scheduled behind this grid's last CTAs
  for (int k = tid; k < DS * SSLOT + DL * LSLOT + NWORK * RW; k += TB) ringS[k] = 0.0;
  if (tid == 0) {
    for (int s = 0; s < DL; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_wait();                                // everything above overlaps the predecessor (ghost fill / wind kernels)

  // ---- which CTAs stage cells that this launch itself has to produce or wait for (GH = 2) ----------------------
  // gcta: the rectangle the CTA stages holds ghost cells (GH = 1: the table says so already); peer_rows: rows that
  // peers deliver (the first n_boundary CTAs of a sharded handle's table).  Everybody else can issue its first row
  // copies before it looks at the control block and the flags.
  const int Ra = (r0 - 3 < g.lo) ? 0 : r0 - 3, Rb = (r1 + 3 > g.hi) ? g.P : r1 + 3;
  const int Ca = (jbase - 3 < g.lo) ? 0 : jbase - 3, Cb = (jend + 3 > g.hi) ? g.P : jend + 3;
  const int nL = max(0, g.lo - Ca), nR = max(0, Cb - g.hi), nT = max(0, g.lo - Ra), nB = max(0, Rb - g.hi);
  const bool gcta = GH == 1 || (GH == 2 && __shfl_sync(0xffffffffu, (nL | nR | nT | nB) != 0 ? 1 : 0, 0) != 0);   // warp-uniform
  const bool peer_rows = GH == 2 && a.gf_flags && cta < a.n_boundary;
  const bool late_issue = GH == 2 && (gcta || peer_rows);
  // TMA row copies are issued by one elected lane of warp 0 (one row per marched row, PF rows ahead).
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
  auto tma = [&](uint32_t dst, const double* src, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(row_bytes), "r"(bar)
                 : "memory");
  };
  // pointers to the staged rows of the row with (compile-time) phase k
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
  // Source pointers of the next row to issue, one per staged array, each advanced by one row pitch per issue:
  // the elected lane's path between the two barriers of a row is two uniform instructions per address (it was
  // five with base + offset - constant: profiles/r2_fused_v2b_ncu.md, "issue path").  Q, v and the centre
  // metrics are staged at the marched row R, sqrtg_pu one row late (R-1), u two rows late (R-2).
  auto rowp = [&](const double* base, int row) {
    return reinterpret_cast<const char*>(base) + (long long)row * ld8;
  };
  const char *pQ = rowp(gq, rfirst), *pV = rowp(gv, rfirst), *pSGC = rowp(gsgc, rfirst), *pSGV = rowp(gsgv, rfirst),
             *pRGC = rowp(grgc, rfirst), *pSGU = rowp(gsgu, rfirst - 1), *pU = rowp(gu, max(rfirst - 2, 0)),
             *pVM = rowp(gvm, rfirst), *pUM = rowp(gum, max(rfirst - 2, 0));
  constexpr int AHEAD = (VAR & 1) ? 3 : PF;                     // rows between the marched row and the row issued
  constexpr int PFD = (VAR & 2) ? 4 : ((VAR & 4) ? 8 : 0);       // L2 prefetch distance beyond the issued row
  auto l2pf = [&](const char* src) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(row_bytes) : "memory");
  };
  constexpr int NCOPY = NS + NL;
  // the TMA copies of the row with phase k (elected lane of warp 0, after its fence.proxy.async)
  auto issue_k = [&](auto kc) {
    constexpr int k = decltype(kc)::value;
    const uint32_t dS = ringS_a + 8u * (uint32_t)((k % DS) * SSLOT), dL = ringL_a + 8u * (uint32_t)(k * LSLOT),
                   bar = full_a + 8u * (uint32_t)k;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(row_bytes * NCOPY)
                 : "memory");
    auto P = [](const char* p) { return reinterpret_cast<const double*>(p); };
    tma(dS + 8u * S_Q * RW, P(pQ), bar);
    tma(dL + 8u * L_V * RW, P(pV), bar);
    tma(dL + 8u * L_SGC * RW, P(pSGC), bar);
    tma(dL + 8u * L_SGV * RW, P(pSGV), bar);
    tma(dL + 8u * L_RGC * RW, P(pRGC), bar);
    tma(dS + 8u * S_SGU * RW, P(pSGU), bar);
    tma(dS + 8u * S_U * RW, P(pU), bar);
    if (MASK & 1) {
      tma(dL + 8u * L_VM * RW, P(pVM), bar);
      tma(dS + 8u * S_UM * RW, P(pUM), bar);
    }
  };
  // L2 prefetch of the HBM-sourced rows PFD rows beyond the row whose copies were just issued (row `ri`)
  auto prefetch_k = [&](int ri) {
    if (PFD > 0 && ri + PFD <= rlast) {
      l2pf(pQ + PFD * ld8);
      l2pf(pV + PFD * ld8);
      l2pf(pU + PFD * ld8);
      if (MASK & 1) { l2pf(pVM + PFD * ld8); l2pf(pUM + PFD * ld8); }
    }
  };
  auto advance = [&]() {                     // warp-uniform
    pQ += ld8; pV += ld8; pSGC += ld8; pSGV += ld8; pRGC += ld8; pSGU += ld8; pU += ld8;
    if (MASK & 1) { pVM += ld8; pUM += ld8; }
  };

  auto first_issue = [&]() {
    if (warp_u == 0) {                         // rows rfirst .. rfirst + AHEAD - 1 (rfirst >= 1: only row -1 is clamped)
      if (elect_one()) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue_k(IC<0>{});
      }
      advance();
      pU = rowp(gu, rfirst - 1);               // the clamp of the first row does not carry over
      pUM = rowp(gum, rfirst - 1);
      if (rfirst + 1 <= rlast) {
        if (elect_one()) issue_k(IC<1>{});
        advance();
      }
      if (AHEAD == 3 && rfirst + 2 <= rlast) {
        if (elect_one()) issue_k(IC<2>{});
        advance();
      }
      if (PFD > 0 && elect_one()) {            // the pointers stand at row rfirst + AHEAD: rows up to PFD - 1 beyond it
        for (int d = 0; d < PFD; ++d)
          if (rfirst + AHEAD + d <= rlast) {
            l2pf(pQ + d * ld8);
            l2pf(pV + d * ld8);
            l2pf(pU + d * ld8);
            if (MASK & 1) { l2pf(pVM + d * ld8); l2pf(pUM + d * ld8); }
          }
      }
    }
  };
  if (!late_issue) first_issue();

  // ---- per-step state from the control block ------------------------------------------------
  const long long step = *((const volatile long long*)&a.ctl->steps);
  if (a.wait_flags) {           // several GPUs: every rank has finished step - 1 (its sum is here, and
    if (tid < a.wait_world) {   // nobody reads the buffers this step's exchange will write any more)
      mg_wait_flag(a.wait_flags + tid, step, a.mg_err, a.mg_timeout_ns);     // acquire: see mgpu.cuh
    }
    __syncthreads();
  }
  double corr = 0.0;
  if (a.apply_corr) {
    if (a.corr_ptr) {
      corr = *a.corr_ptr;
    } else if (*((const volatile int*)&a.ctl->pend)) {
      double sm = 0.0;
      if (a.pub.world > 1) {
        const volatile double* ps = a.pub.peer_sync[a.pub.rank]->psum[step & 1];
        for (int k = 0; k < a.pub.world; ++k) sm += ps[k];
      } else {
        sm = *((const volatile double*)&a.ctl->sum);
      }
      corr = -sm * a.inv_a2;
    }
  }
  const double ws = (MASK & 2) ? a.ws_tab[step & a.ws_mask] : 1.0;
  const double cdx = a.cdx * ws, cdy = a.cdy * ws;       // time factor of a separable wind folded in

  // ---- GH = 2: ghost cells of the rectangle this CTA stages (src/interpolation.py:154-314), written to Q itself
  // before the first row copy reads them.  Neighbouring CTAs whose rectangles overlap write the same bits.  At a
  // panel edge all four ghost layers (and the 4 x 4 corners) are filled although the march reads three, so that the
  // ring of the array is complete for whoever restores it after the run (stepper.cu: copy_ring_kernel).
  if (GH == 2) {
    if (peer_rows) {                         // the peers' pieces of this rank's halo rows and ghost-fill sources
      if (tid < a.wait_world) {
        const long long xc = *((const volatile long long*)&a.ctl->xcount);
        mg_wait_flag(a.gf_flags + tid, xc, a.mg_err, a.mg_timeout_ns);
      }
      __syncthreads();
    }
    if (gcta) {
      ghost_prologue(a, p, Ra, Rb, Ca, Cb, tid, TB);
      // the row copies below read these cells through the async proxy
      asm volatile("fence.proxy.async;" ::: "memory");
      __threadfence();
      __syncthreads();
    }
  }

  if (late_issue) first_issue();
  double* __restrict__ QN = a.qn + (long long)p * g.ps + PYCS_JOFF + min(j, g.P - 1) + (long long)r0 * g.ld;
  // GH: ghost(sqrtg) of the thread's column, one row ahead of the march (an L2 hit that must not sit on
  // the critical path of phase 1); only threads that own a ghost cell in that row load it
  const double* __restrict__ GS = nullptr;
  double gs_next = 0.0;
  if (GH) {
    GS = a.gs + (long long)p * g.ps + PYCS_JOFF + min(j, g.P - 1) + (long long)rfirst * g.ld;
    if (gcta && a.apply_corr && !(jint && rfirst >= g.lo && rfirst < g.hi)) gs_next = __ldg(GS);
  }
  Lane L;
  lane_init(L);
  uint32_t parb = 0;
  // warp 0: TMA copies of row r + AHEAD (the generic-proxy reads of the slots it overwrites were ordered before this
  // point by the barrier the caller has just passed)
  auto issue_ahead = [&](auto kc, int r) {
    constexpr int k = decltype(kc)::value;
    if (warp_u == 0 && r + AHEAD <= rlast) {
      if (elect_one()) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue_k(IC<(k + AHEAD) % DL>{});
        prefetch_k(r + AHEAD);
      }
      advance();
    }
  };
  // one marched row; false once the chunk is finished
  // GC (compile time): the CTA stages ghost cells and adds their projection term itself (GH = 1 always, GH = 2 for
  // the CTAs at a panel edge -- those run their own copy of the march, so that the loop of the others is the GH = 0
  // loop: one loop with both forms ran 8 % slower, 28 KB of code per trip against 21.6 KB)
  auto row = [&](auto kc, auto gc, int r) -> bool {
    constexpr int k = decltype(kc)::value;
    constexpr bool GC = decltype(gc)::value;
    if (r > rlast) return false;
    constexpr int kS = (k % DS) * SSLOT;
    while (!mbar_try_wait(&full[k], parb)) {}
    const RowPtrs R = rowptrs(IC<k>{});
    double qnew[1] = {R.q[0]};
    if (a.apply_corr) {
      // pending MF-PR term on the own cell of the new row (neighbours read it after barrier A)
      const bool inner = jint && r >= g.lo && r < g.hi;
      if (GC) {
        qnew[0] = fma(inner ? R.sgc0[0] : gs_next, corr, qnew[0]);
        ringS[kS + S_Q * RW + e] = qnew[0];
        GS += g.ld;
        if (r < rlast && !(jint && r + 1 >= g.lo && r + 1 < g.hi)) gs_next = __ldg(GS);
      } else if (inner) {
        qnew[0] = fma(R.sgc0[0], corr, qnew[0]);
        ringS[kS + S_Q * RW + e] = qnew[0];
      }
    }
    // ---------------- phase 1: own column
    XEdge X;
    double qx[1];
    phase_x_inner<RECON, SPLIT, MASK, k, WLEN, (VAR & 16) != 0>(L, X, R, qnew, cdx, qx);
    sX[e] = qx[0];
    double cch0[1], cch3[1];
    if (VAR & 32) { cch0[0] = R.v0[0] * cdy; cch3[0] = R.v3[0] * cdy; }
    __syncthreads();                                   // barrier A
    if (!(VAR & 1)) issue_ahead(IC<k>{}, r);           // warp 0 issues the TMA copies of row r+PF
    // ---------------- phase 2: y-fluxes at edge j: inner on Q row r, outer on Qx row r-3
    double F[1], G[1], CF[1] = {0.0}, CG[1];
    yflux_pair<RECON, SPLIT, MASK, k, (VAR & 8) != 0, (VAR & 32) != 0>(R.v0, R.vm0, R.sgv0, R.sgc0, R.q, cdy, F, CF, qnew, cch0);
    yflux_pair<RECON, SPLIT, MASK, k, (VAR & 8) != 0, (VAR & 32) != 0>(R.v3, R.vm3, R.sgv3, R.sgc3, sX + e, cdy, G, CG, qx, cch3);
    sF[e] = F[0];
    sG[e] = G[0];
    if (SPLIT != 1) sC[e] = CF[0];
    __syncthreads();                                   // barrier B
    if (VAR & 1) issue_ahead(IC<k>{}, r);              // rows r (short ring) and r-3 (long ring) are dead: row r+3
    // ---------------- phase 3: Qy row r, outer x-flux on Qy, output row r-3
    double Fn[1], Gn[1], CFn[1] = {0.0}, out[1], sdiv[1];
    Fn[0] = sF[e + 1];
    Gn[0] = sG[e + 1];
    if (SPLIT != 1) CFn[0] = sC[e + 1];
    phase_x_outer<RECON, SPLIT, k>(L, X, R, F, Fn, G, Gn, CF, CFn, out, sdiv);
    if (r >= r0 + 3) {
      if (out_lane) {
        *QN = out[0];
        L.psum += sdiv[0];
      }
      QN += g.ld;
    }
    if (GH == 1 && a.edge_flux) {
      // MF-AF: the outer fluxes (times dt/dx, as the kernel carries them) on the panel's edge lines, interior extent
      const int ex = r - 2, ry = r - 3;                 // x-edge just completed; row of the outer y-fluxes
      if ((ex == g.lo || ex == g.hi) && out_lane)
        a.edge_flux[((long long)p * 4 + (ex == g.lo ? 0 : 1)) * g.N + (j - g.lo)] = L.fout_prev[0];
      if ((j == g.lo || j == g.hi) && ry >= r0 && ry < r1)
        a.edge_flux[((long long)p * 4 + (j == g.lo ? 2 : 3)) * g.N + (ry - g.lo)] = G[0];
    }
    return true;
  };
  auto march = [&](auto gc) {
    for (int rb = rfirst;; rb += DL, parb ^= 1u) {
      if (!row(IC<0>{}, gc, rb)) break;
      if (!row(IC<1>{}, gc, rb + 1)) break;
      if (!row(IC<2>{}, gc, rb + 2)) break;
      if (!row(IC<3>{}, gc, rb + 3)) break;
      if (!row(IC<4>{}, gc, rb + 4)) break;
      if (!row(IC<5>{}, gc, rb + 5)) break;
    }
  };
  if (GH == 2) {
    if (gcta) march(BC<true>{});
    else march(BC<false>{});
  } else {
    march(BC<GH == 1>{});
  }

  // ---- several GPUs, boundary CTAs: ship what the peers read of this CTA's rows (see FusedArgs::xjobs)
  if (GH && a.xjob_off && cta < a.n_boundary) {
    __syncthreads();                           // the CTA's output rows are visible to all its threads
    const long long xc = *((const volatile long long*)&a.ctl->xcount);
    const int jb0 = a.xjob_off[cta], jb1 = a.xjob_off[cta + 1];
    for (int jb = jb0; jb < jb1; ++jb) {
      const int4 x = a.xjobs[jb];
      const int peer = x.x & 15, i0 = x.x >> 4, w = x.w - x.z, n = (x.y - i0) * w;
      double* __restrict__ dst = a.xpeer_q[peer];
      for (int t = tid; t < n; t += TB) {
        const long long id = gidx(g, p, i0 + t / w, x.z + t % w);
        dst[id] = __ldcg(a.qn + id);
      }
    }
    __syncthreads();
    if (tid == 0) {
      __threadfence_system();                  // cumulative over the stores the barrier ordered before it
      sF[41] = (atomicAdd(a.xcounter, 1u) == (unsigned)a.n_boundary - 1u) ? 1.0 : 0.0;
    }
    __syncthreads();
    if (sF[41] != 0.0 && tid < 32) {           // last boundary CTA: every piece of this rank's exchange is out
      if (tid == 0) {
        *a.xcounter = 0u;
        a.ctl->xcount = xc + 1;
      }
      if (tid < a.pub.world) {
        __threadfence_system();
        *((volatile long long*)&a.pub.peer_sync[tid]->dflag[a.pub.rank]) = xc + 1;
      }
    }
  }

  // ---- per-CTA partial of sum(pxdF + pydF) over its outputs (MF-PR), fixed order; the last CTA of
  // the step (of both launches of a split step) totals them and closes the step
  __syncthreads();
  double v = L.psum;
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((tid & 31) == 0) sF[tid >> 5] = v;
  __syncthreads();
  const int ntot = a.cta_tab ? a.nblk_total : (int)gridDim.x;
  const int mgw = a.pub.world;
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < TB / 32; ++w) t += sF[w];
    a.part[cta] = t;
    // device scope suffices here: the ticket orders the partial sums; peer stores (in-kernel exchange) were
    // fenced at system scope before their own ticket, and the publication below fences its own stores
    sF[40] = fused_last_writer(a.counter, (unsigned)ntot, false) ? 1.0 : 0.0;
  }
  __syncthreads();
  if (sF[40] != 0.0) {                       // CTA-uniform: the last CTA of the step
    // total of the partials in a fixed order, by the whole CTA: thread t adds part[t], part[t + TB], ... (all loads
    // in flight at once), then warps, then the five warp sums -- ~1 us where one warp took ~5 (36 dependent L2 trips
    // at 1140 CTAs), and that time is serial: nothing of the next step can start before it
    __threadfence();
    double pv[8];
    double tsum = 0.0;
    for (int k0 = tid; k0 < ntot; k0 += 8 * TB) {
#pragma unroll
      for (int u = 0; u < 8; ++u) pv[u] = (k0 + u * TB < ntot) ? __ldcg(a.part + k0 + u * TB) : 0.0;
#pragma unroll
      for (int u = 0; u < 8; ++u) tsum += pv[u];
    }
    for (int o = 16; o > 0; o >>= 1) tsum += __shfl_down_sync(0xffffffffu, tsum, o);
    if ((tid & 31) == 0) sG[tid >> 5] = tsum;
    __syncthreads();
  }
  if (sF[40] != 0.0 && tid < 32) {
    double tot = 0.0;
    for (int w = 0; w < TB / 32; ++w) tot += sG[w];
    if (tid == 0) *a.counter = 0u;
    if (!a.timing) {
      if (tid == 0) {
        a.ctl->sum = tot;
        a.ctl->corr_applied = corr;
        a.ctl->pend = a.apply_corr;
        __threadfence();
        a.ctl->steps = step + 1;
      }
      if (mgw > 1 && tid < mgw) {             // publish this rank's sum, then raise its flag on every rank
        MgSync* sy = a.pub.peer_sync[tid];
        sy->psum[(step + 1) & 1][a.pub.rank] = tot;
        __threadfence_system();
        *((volatile long long*)&sy->sflag[a.pub.rank]) = step + 1;
      }
    }
  }
}

template <int TB, int MASK>
constexpr size_t smem_bytes() {
  constexpr int DS = F2B_PF + 1, DL = F2B_PF + 4, NWORK = 4;
  return sizeof(double) * (size_t)(TB + 6) * (DS * ((MASK & 1) ? 4 : 3) + DL * ((MASK & 1) ? 5 : 4) + NWORK) +
         sizeof(uint64_t) * DL + 16;
}

template <int TB, int RECON, int SPLIT, int MASK, int GH, int VAR>
cudaError_t launch_var(const FusedArgs& a, int nblocks, cudaStream_t st, int* resident) {
  static bool configured = false;
  const size_t smem = smem_bytes<TB, MASK>();
  auto kern = fused2b_kernel<TB, RECON, SPLIT, MASK, GH, VAR>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (resident) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(resident, kern, TB, smem);
  if (a.pdl) return pycs_launch_pdl(kern, dim3(nblocks), dim3(TB), smem, st, true, a);
  kern<<<nblocks, TB, smem, st>>>(a);
  return cudaSuccess;
}

#ifdef F2B_VARIANTS
// PYCS_VARIANT selects among the instantiated march variants of the par-default scheme's serial-step kernels
// (experiment knob; anything not instantiated falls back to the default)
static int march_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PYCS_VARIANT");
    v = e ? atoi(e) : F2B_VAR_DEFAULT;
  }
  return v;
}
#endif
template <int TB, int RECON, int SPLIT, int MASK, int GH>
cudaError_t launch_one(const FusedArgs& a, int nblocks, cudaStream_t st, int* resident) {
#ifdef F2B_VARIANTS
  if constexpr (RECON == 3 && SPLIT == 1 && GH == 0 && MASK != 1) {
    const int v = march_variant();
#define F2B_V(V) if (v == V && V != F2B_VAR_DEFAULT) return launch_var<TB, RECON, SPLIT, MASK, GH, V>(a, nblocks, st, resident)
    F2B_VARIANTS
#undef F2B_V
  }
#endif
  return launch_var<TB, RECON, SPLIT, MASK, GH, F2B_VAR_DEFAULT>(a, nblocks, st, resident);
}

template <int RECON, int SPLIT>
cudaError_t launch_mask(const FusedArgs& a, int mask, int gh, int nblocks, cudaStream_t st, int* resident) {
#define F2B_CASE(M, G) \
  if (mask == M && gh == G) return launch_one<F2B_TB, RECON, SPLIT, M, G>(a, nblocks, st, resident)
  F2B_CASE(0, 0); F2B_CASE(1, 0); F2B_CASE(2, 0);
  F2B_CASE(0, 1); F2B_CASE(1, 1); F2B_CASE(2, 1);
  F2B_CASE(0, 2); F2B_CASE(1, 2); F2B_CASE(2, 2);
#undef F2B_CASE
  return cudaErrorInvalidValue;
}

cudaError_t dispatch(const FusedArgs& a, int recon, int split, int mask, int gh, int nblocks, cudaStream_t st,
                     int* resident) {
#define F2B_SCHEME(R, S) \
  if (recon == R && split == S) return launch_mask<R, S>(a, mask, gh, nblocks, st, resident)
  F2B_SCHEME(3, 1); F2B_SCHEME(3, 2); F2B_SCHEME(3, 3);
  F2B_SCHEME(1, 1); F2B_SCHEME(1, 2); F2B_SCHEME(1, 3);
#undef F2B_SCHEME
  return cudaErrorInvalidValue;
}

}  // namespace

int pycs_plan_split_ctas(int row_lo, int row_hi, int nstrips, int band, int edge_rows, int rows,
                         std::vector<CtaDesc>* out) {
  out->clear();
  const int nrows = row_hi - row_lo;
  if (band < 3) band = 3;
  if (edge_rows < 4) edge_rows = 4;
  if (rows < 4) rows = 4;
  auto chunks = [&](int a, int b, int len, int s0, int s1) {        // rows [a, b) in chunks of <= len, strips [s0, s1)
    if (b <= a) return;
    const int n = (b - a + len - 1) / len, per = (b - a + n - 1) / n;
    for (int c = 0; c < n; ++c) {
      const int r0 = a + c * per, r1 = r0 + per < b ? r0 + per : b;
      if (r0 >= r1) continue;
      for (int st = s0; st < s1; ++st)
        for (int p = 0; p < 6; ++p) out->push_back(CtaDesc{r0, r1, st, p});   // panel fastest: shared metric rows
    }
  };
  if (nstrips < 3 || nrows < 2 * band + 8) {           // no interior: everything is boundary
    chunks(row_lo, row_hi, edge_rows > band ? edge_rows : band, 0, nstrips);
    return (int)out->size();
  }
  chunks(row_lo, row_lo + band, band, 0, nstrips);                                  // lower band
  chunks(row_hi - band, row_hi, band, 0, nstrips);                                  // upper band
  chunks(row_lo + band, row_hi - band, edge_rows, 0, 1);                            // first strip
  chunks(row_lo + band, row_hi - band, edge_rows, nstrips - 1, nstrips);            // last strip
  const int nb = (int)out->size();
  chunks(row_lo + band, row_hi - band, rows, 1, nstrips - 1);                       // interior
  return nb;
}

bool pycs_fused2b_has(int recon, int split) { return (recon == 1 || recon == 3) && split >= 1 && split <= 3; }
int pycs_fused2b_threads() { return F2B_TB; }

cudaError_t pycs_launch_fused2b(const FusedArgs& a, int recon, int split, int mask, int gh, int nblocks,
                                cudaStream_t st) {
  return dispatch(a, recon, split, mask, gh, nblocks, st, nullptr);
}

int pycs_fused2b_resident(int recon, int split, int mask, int gh) {
  FusedArgs a{};
  int n = 0;
  if (dispatch(a, recon, split, mask, gh, 0, nullptr, &n) != cudaSuccess) return -1;
  return n;
}
