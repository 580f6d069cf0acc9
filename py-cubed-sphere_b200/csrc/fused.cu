// Fused advection step for the duo-grid (ET-DG) schemes: host-side driver of the production
// path (step sequencing, ghost fill, deferred MF-PR projection, separable wind, multi-GPU
// hooks) and the first-generation step kernel (v2).  The default step kernel is v2b in
// fused2b.cu (same march, leaner arithmetic from fused3_core.cuh); v3 is in fused3.cu;
// PYCS_FUSED_IMPL = 2 | 3 | 4 selects.  v2 remains the kernel of the limited reconstructions
// (PPM-CW84 / PPM-L04).
//
// One kernel does what src/discrete_operators.py:18-101 + src/advection_timestep.py:43
// do in ~60 whole-array passes: inner x/y PPM fluxes, the splitting update
// (Qx, Qy), outer fluxes on (Qy, Qx), the metric-weighted divergence and the
// Q update.  Per cell it reads Q, the two time-averaged winds and sqrt(g)
// (40 B algorithmic) and writes Q; q_L/q_R/flux/Qx/Qy never
// leave the SM.
//
// Decomposition: a CTA owns a strip of TB-6 output columns (j, the contiguous
// axis) of one panel and marches along i over a chunk of rows.  Each thread
// owns one column:
//   * x-direction sweeps (stencil across rows) live in registers: a rolling
//     5-row window of Q and of Qy per thread, one new parabola per row;
//   * y-direction sweeps (stencil along the row) go through four TB-wide
//     shared-memory rows (Q row, Qx row, two flux rows), two __syncthreads per
//     marched row.
// Row r entering the window produces the output of row r-3 (the 7x7 dependence
// box of the split scheme).  Ghost cells come from the Lagrange fill that runs
// before the kernel; with ET-DG nothing else crosses a panel edge.
//
// MF-PR (src/discrete_operators.py:98-101) needs a global sum: the kernel writes
// Q - dt*div and per-CTA partial sums; the projection term sqrtg*m0/a2 is added
// when the next consumer loads Q (the next step's fill + fused kernel, or the
// flush kernel before anything else reads Q).  Algebraically identical, one
// rounding apart from the reference order.
#include <cstdio>
#include <map>
#include <vector>
#include "pycs_common.cuh"
#include "fused_args.cuh"
#include "fused3_core.cuh"
#include "mgpu.cuh"
#include "ghost_core.cuh"
static int f3_strip_capacity(int nw) { return f3::strip_capacity(nw); }

namespace {


// ---- PPM edge values of one cell from its 5-point neighbourhood -------------
// src/reconstruction_1d.py:36-192 (q3 is the cell itself)
template <int RECON>
__device__ __forceinline__ void edge_values(double q1, double q2, double q3, double q4, double q5,
                                            double& l, double& r) {
  if (RECON == 3) {
    const double a1 = 2.0 / 60.0, a2 = -13.0 / 60.0, a3 = 47.0 / 60.0, a4 = 27.0 / 60.0, a5 = -3.0 / 60.0;
    r = fma(a5, q5, fma(a4, q4, fma(a3, q3, fma(a2, q2, a1 * q1))));
    l = fma(a1, q5, fma(a2, q4, fma(a3, q3, fma(a4, q2, a5 * q1))));
  } else if (RECON == 1) {
    const double c7 = 7.0 / 12.0, c1 = 1.0 / 12.0;
    l = c7 * (q3 + q2) - c1 * (q4 + q1);
    r = c7 * (q4 + q3) - c1 * (q5 + q2);
  } else if (RECON == 2) {
    double qq[5] = {q1, q2, q3, q4, q5};
    double dQ[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double qm = qq[k], q0 = qq[k + 1], qp = qq[k + 2];
      double d0 = 0.5 * (qp - qm), d1 = 2.0 * (qp - q0), d2 = 2.0 * (q0 - qm);
      double d = fmin(fmin(fabs(d0), fabs(d1)), fabs(d2));
      d = (d0 > 0.0) ? d : ((d0 < 0.0) ? -d : 0.0);
      dQ[k] = ((qp - q0) * (q0 - qm) > 0.0) ? d : 0.0;
    }
    l = 0.5 * (q3 + q2) - (dQ[1] - dQ[0]) / 6.0;
    r = 0.5 * (q4 + q3) - (dQ[2] - dQ[1]) / 6.0;
    double dq = r - l, q6 = 6.0 * q3 - 3.0 * (r + l);
    if ((r - q3) * (q3 - l) <= 0.0) { r = q3; l = q3; }
    bool over = fabs(dq) < fabs(q6);
    double rl = r - l, mid = q3 - 0.5 * (r + l);
    bool left = rl * mid > (rl * rl) / 6.0;
    bool right = -(rl * rl) / 6.0 > rl * mid;
    if (over && left) l = 3.0 * q3 - 2.0 * r;
    if (over && right) r = 3.0 * q3 - 2.0 * l;
  } else {
    double qq[5] = {q1, q2, q3, q4, q5};
    double mono[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double qm = qq[k], q0 = qq[k + 1], qp = qq[k + 2];
      double d = 0.25 * (qp - qm);
      double dmin = fmax(fmax(qm, q0), qp) - q0;
      double dmax = q0 - fmin(fmin(qm, q0), qp);
      double m = fmin(fmin(fabs(d), dmin), dmax);
      mono[k] = (d > 0.0) ? m : ((d < 0.0) ? -m : 0.0);
    }
    l = 0.5 * (q3 + q2) - (mono[1] - mono[0]) / 3.0;
    r = 0.5 * (q4 + q3) - (mono[2] - mono[1]) / 3.0;
    double m = mono[1], am = 2.0 * fabs(m);
    double s = (m > 0.0) ? 1.0 : ((m < 0.0) ? -1.0 : 0.0);
    l = q3 - fmin(am, fabs(l - q3)) * s;
    r = q3 + fmin(am, fabs(r - q3)) * s;
  }
}

// Parabola of a cell in flux form (src/flux.py:27-44): metric-weighted edge values
// (MT-0) and the coefficients dq, q6.
template <int RECON, int MT>
__device__ __forceinline__ void parabola(double q1, double q2, double q3, double q4, double q5,
                                         double gl, double gr, double gcc, double& eL, double& eR,
                                         double& q6, double& dq) {
  double l, r;
  edge_values<RECON>(q1, q2, q3, q4, q5, l, r);
  double q = q3;
  if (MT == 1) { l *= gl; r *= gr; q *= gcc; }
  eL = l;
  eR = r;
  dq = r - l;
  q6 = 3.0 * fma(2.0, q, -(r + l));
}

// CW84 eq. 1.12 as written at src/flux.py:48-61:
//   u >= 0: f = q_R + c/2 (q6 - dq) - q6 c^2/3     (s = +1, e = q_R of the left cell)
//   u <  0: f = q_L - c/2 (q6 + dq) - q6 c^2/3     (s = -1, e = q_L of the right cell)
__device__ __forceinline__ double ppm_flux(double e, double q6, double dq, double s, double c) {
  double t = fma(s, q6, -dq);
  double f = fma(0.5 * c, t, e);
  return fma(-(q6 * c), c * (1.0 / 3.0), f);
}

// ---- TMA / mbarrier helpers (cp.async.bulk 1-D row copies into a shared-memory ring) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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

// Arrays staged per marched row (slot of row r): Q[r], V[r], SGC[r], SGV[r], RGC[r],
// SGU[r-1], U[r-2] (+ mask sources VM[r], UM[r-2] for RK2).
enum { A_Q = 0, A_V = 1, A_SGC = 2, A_SGV = 3, A_RGC = 4, A_SGU = 5, A_U = 6, A_VM = 7, A_UM = 8 };

template <int TB, int RECON, int SPLIT, int MASK, int D>
__global__ void __launch_bounds__(TB) fused_step_kernel(FusedArgs a) {
  constexpr int MT = (SPLIT == 3) ? 2 : 1;
  constexpr int NARR = (MASK & 1) ? 9 : 7;
  constexpr int TBW = TB + 2;               // staged row: columns jbase-4 .. jbase+TB-3
  constexpr int PF = D - 4;                 // rows in flight ahead of the consumer
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* ring = reinterpret_cast<double*>(smem_raw);            // [D][NARR][TBW]
  double* sX = ring + D * NARR * TBW;                            // Qx row r-3
  double* sF = sX + TBW;                                         // inner y-flux, row r
  double* sG = sF + TBW;                                         // outer y-flux, row r-3
  double* sC = sG + TBW;                                         // sqrtg_pv*cy (SPLIT != 1)
  uint64_t* full = reinterpret_cast<uint64_t*>(sC + TBW);        // [D]

  const Geo& g = a.g;
  int b = blockIdx.x;
  const int p = b % 6;
  b /= 6;
  const int strip = b % a.nstrips, chunk = b / a.nstrips;
  const int tid = threadIdx.x;
  const int e = tid + 1;                    // own element in a staged row
  const int jbase = g.lo + strip * a.wcols;
  const int jend = min(jbase + a.wcols, g.hi);
  const int j = jbase - 3 + tid;
  const int r0 = a.row_lo + chunk * a.rows_per_chunk;
  const int r1 = min(r0 + a.rows_per_chunk, a.row_hi);
  const int rfirst = r0 - 3, rlast = r1 + 2;
  const bool out_lane = (tid >= 3) && (j < jend);
  const bool jint = (j >= g.lo) && (j < g.hi);
  const long long L = g.ld;
  const double corr = a.apply_corr ? *a.corr : 0.0;
  const double cdx = a.cdx, cdy = a.cdy;
  // staged segment: starts at column jbase-4 (16-byte aligned: JOFF and wcols are even)
  const int c0 = jbase - 4;
  int len = min(TBW, g.ld - PYCS_JOFF - c0) & ~1;
  const uint32_t row_bytes = (uint32_t)len * 8u;

  auto issue = [&](int r) {                 // thread 0 only
    const int s = (r - rfirst) % D;
    double* dst = ring + (size_t)s * NARR * TBW;
    const long long colb = (long long)p * g.ps + PYCS_JOFF + c0;
    const long long colm = PYCS_JOFF + c0;
    const long long rr = (long long)r * L, ru = (long long)max(r - 1, 0) * L, r2 = (long long)max(r - 2, 0) * L;
    mbar_expect_tx(&full[s], row_bytes * NARR);
    tma_row(dst + A_Q * TBW, a.q + colb + rr, row_bytes, &full[s]);
    tma_row(dst + A_V * TBW, a.va + colb + rr, row_bytes, &full[s]);
    tma_row(dst + A_SGC * TBW, a.sgc + colm + rr, row_bytes, &full[s]);
    tma_row(dst + A_SGV * TBW, a.sgv + colm + rr, row_bytes, &full[s]);
    tma_row(dst + A_RGC * TBW, a.rgc + colm + rr, row_bytes, &full[s]);
    tma_row(dst + A_SGU * TBW, a.sgu + colm + ru, row_bytes, &full[s]);
    tma_row(dst + A_U * TBW, a.ua + colb + r2, row_bytes, &full[s]);
    if (MASK & 1) {
      tma_row(dst + A_VM * TBW, a.vm + colb + rr, row_bytes, &full[s]);
      tma_row(dst + A_UM * TBW, a.um + colb + r2, row_bytes, &full[s]);
    }
  };

  if (tid == 0) {
    for (int s = 0; s < D; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int r = rfirst; r < rfirst + PF && r <= rlast; ++r) issue(r);
  }

  double* __restrict__ QN = a.qn + (long long)p * g.ps + PYCS_JOFF + min(j, g.P - 1) + (long long)r0 * L;

  // rolling state, all for column j
  double q0 = 0, q1 = 0, q2 = 0, q3 = 0, q4 = 0;       // Q rows r-4 .. r
  double y0 = 0, y1 = 0, y2 = 0, y3 = 0, y4 = 0;       // Qy rows r-4 .. r
  double pqR = 0, pq6 = 0, pdq = 0;                    // parabola of Q, cell r-3 (u >= 0 side)
  double pyR = 0, py6 = 0, pyd = 0;                    // parabola of Qy, cell r-3
  double fin_prev = 0, fout_prev = 0;                  // x-fluxes at edge r-3
  double cm_prev = 0;                                  // sqrtg_pu * cx at edge r-3 (SPLIT != 1)
  double gu_prev = 0;                                  // sqrtg_pu at row r-2 (carried from the previous slot)
  double psum = 0;
  int s = 0, s2 = D - 2, s3 = D - 3;                   // slots of rows r, r-2, r-3
  uint32_t par = 0;

  for (int r = rfirst; r <= rlast; ++r) {
    const double* S = ring + (size_t)s * NARR * TBW;
    const double* S2 = ring + (size_t)s2 * NARR * TBW;
    const double* S3 = ring + (size_t)s3 * NARR * TBW;
    while (!mbar_try_wait(&full[s], par)) {}

    const bool have_cell = (r >= r0 + 1);   // cell r-2 has its 5 rows
    const bool have_edge = (r >= r0 + 2);   // edge r-2: cells r-3 and r-2 both have parabolas
    const bool outp = (r >= r0 + 3);        // output row r-3

    // ---------------- phase 1: own column -- new row, inner x-flux at edge r-2, Qx row r-3
    q0 = q1; q1 = q2; q2 = q3; q3 = q4;
    y0 = y1; y1 = y2; y2 = y3; y3 = y4;
    q4 = S[A_Q * TBW + e];
    if (a.apply_corr) {
      if (jint && r >= g.lo && r < g.hi) {
        q4 = fma(S[A_SGC * TBW + e], corr, q4);
        const_cast<double*>(S)[A_Q * TBW + e] = q4;      // neighbours read the corrected value
      }
    }
    const double gu_cur = S[A_SGU * TBW + e];            // sqrtg_pu at row r-1
    double nqL = 0, nqR = 0, nq6 = 0, ndq = 0;           // parabola of Q, cell r-2
    double u = 0, cxe = 0, sup = 1.0;
    bool up = true;
    double fin = 0, gcc2 = 0;
    if (have_cell) {
      gcc2 = S2[A_SGC * TBW + e];
      parabola<RECON, MT>(q0, q1, q2, q3, q4, gu_prev, gu_cur, gcc2, nqL, nqR, nq6, ndq);
      if (have_edge) {
        u = S[A_U * TBW + e];
        if (MASK & 2) u *= a.ws;
        const double um = (MASK & 1) ? S[A_UM * TBW + e] : u;
        up = um >= 0;
        sup = up ? 1.0 : -1.0;
        cxe = u * cdx;
        const double ee = up ? pqR : nqL, s6 = up ? pq6 : nq6, sd = up ? pdq : ndq;
        fin = ppm_flux(ee, s6, sd, sup, cxe) * u;
        if (MT == 2) fin *= gu_prev;
      }
    }
    if (outp) {
      const double dFx = -(fin - fin_prev) * cdx;
      const double rg = S3[A_RGC * TBW + e];
      double qx;
      if (SPLIT == 1) qx = fma(0.5 * dFx, rg, q1);
      else {
        const double cd = gu_prev * cxe - cm_prev;        // c1x - c2x
        if (SPLIT == 2) qx = fma(0.5 * fma(cd, q1, dFx), rg, q1);
        else qx = 0.5 * (q1 + (q1 + dFx) / (1.0 - cd));
      }
      sX[e] = qx;
    }
    __syncthreads();                                     // barrier A
    if (tid == 0 && r + PF <= rlast) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(r + PF);
    }

    // ---------------- phase 2: y-fluxes at edge j: inner on row r, outer on row r-3
    {
      double v = S[A_V * TBW + e];
      if (MASK & 2) v *= a.ws;
      const double vm = (MASK & 1) ? S[A_VM * TBW + e] : v;
      const bool vp = vm >= 0;
      int cu = vp ? e - 1 : e;
      cu = max(2, min(cu, TBW - 3));
      const double gl = S[A_SGV * TBW + cu], gr = S[A_SGV * TBW + cu + 1], gcc = S[A_SGC * TBW + cu];
      double eL, eR, s6, sd;
      const double* qr = S + A_Q * TBW + cu;
      parabola<RECON, MT>(qr[-2], qr[-1], qr[0], qr[1], qr[2], gl, gr, gcc, eL, eR, s6, sd);
      const double cy = v * cdy;
      double f = ppm_flux(vp ? eR : eL, s6, sd, vp ? 1.0 : -1.0, cy) * v;
      if (MT == 2 || SPLIT != 1) {
        const double gvo = S[A_SGV * TBW + e];
        if (MT == 2) f *= gvo;
        if (SPLIT != 1) sC[e] = gvo * cy;
      }
      sF[e] = f;
    }
    if (outp) {
      double v3 = S3[A_V * TBW + e];
      if (MASK & 2) v3 *= a.ws;
      const double vm = (MASK & 1) ? S3[A_VM * TBW + e] : v3;
      const bool vp = vm >= 0;
      int cu = vp ? e - 1 : e;
      cu = max(2, min(cu, TBW - 3));
      const double gl = S3[A_SGV * TBW + cu], gr = S3[A_SGV * TBW + cu + 1], gcc = S3[A_SGC * TBW + cu];
      double eL, eR, s6, sd;
      const double* xr = sX + cu;
      parabola<RECON, MT>(xr[-2], xr[-1], xr[0], xr[1], xr[2], gl, gr, gcc, eL, eR, s6, sd);
      const double cy = v3 * cdy;
      double f = ppm_flux(vp ? eR : eL, s6, sd, vp ? 1.0 : -1.0, cy) * v3;
      if (MT == 2) f *= S3[A_SGV * TBW + e];
      sG[e] = f;
    }
    __syncthreads();                                     // barrier B

    // ---------------- phase 3: Qy row r, outer x-flux at edge r-2 on Qy, output row r-3
    {
      const double dFy = -(sF[e + 1] - sF[e]) * cdy;
      const double rg = S[A_RGC * TBW + e];
      if (SPLIT == 1) y4 = fma(0.5 * dFy, rg, q4);
      else {
        const double cd = sC[e + 1] - sC[e];              // c1y - c2y
        if (SPLIT == 2) y4 = fma(0.5 * fma(cd, q4, dFy), rg, q4);
        else y4 = 0.5 * (q4 + (q4 + dFy) / (1.0 - cd));
      }
    }
    double fout = 0;
    if (have_cell) {
      double nyL, nyR, ny6, nyd;
      parabola<RECON, MT>(y0, y1, y2, y3, y4, gu_prev, gu_cur, gcc2, nyL, nyR, ny6, nyd);
      if (have_edge) {
        const double ee = up ? pyR : nyL, s6 = up ? py6 : ny6, sd = up ? pyd : nyd;
        fout = ppm_flux(ee, s6, sd, sup, cxe) * u;
        if (MT == 2) fout *= gu_prev;
      }
      pyR = nyR; py6 = ny6; pyd = nyd;
    }
    if (outp) {
      const double sdiv = -(fout - fout_prev) * cdx - (sG[e + 1] - sG[e]) * cdy;   // pxdF + pydF
      if (out_lane) {
        // Q - dt*div with div = -(pxdF+pydF)/(dt*sqrtg)  (src/discrete_operators.py:95,
        // src/advection_timestep.py:43)
        *QN = fma(sdiv, S3[A_RGC * TBW + e], q1);
        psum += sdiv;
      }
      QN += L;
    }
    pqR = nqR; pq6 = nq6; pdq = ndq;
    fin_prev = fin;
    fout_prev = fout;
    if (SPLIT != 1) cm_prev = gu_prev * cxe;
    gu_prev = gu_cur;
    // advance the ring: row r+1 -> slot s+1 (rows r-1, r-2 follow in lock step)
    s = (s + 1 == D) ? 0 : s + 1;
    s2 = (s2 + 1 == D) ? 0 : s2 + 1;
    s3 = (s3 + 1 == D) ? 0 : s3 + 1;
    if (s == 0) par ^= 1;
  }

  // per-CTA partial of sum(pxdF + pydF) over its interior outputs (for MF-PR)
  __syncthreads();
  double v = psum;
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((tid & 31) == 0) sF[tid >> 5] = v;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < TB / 32; ++w) t += sF[w];
    a.part[blockIdx.x] = t;
    sF[40] = fused_last_writer(a.counter, gridDim.x) ? 1.0 : 0.0;
  }
  __syncthreads();
  if (sF[40] != 0.0 && tid < 32) {            // last CTA of the launch: total in a fixed order
    const double tot = fused_warp_sum(a.part, (int)gridDim.x, tid);
    if (tid == 0) {
      *a.sum_out = tot;
      *a.counter = 0u;
    }
  }
}


// ---- Lagrange ghost fill that honours the pending projection term ------------
// The fill is linear, so ghost(Q + corr*sqrtg) = ghost(Q) + corr * ghost(sqrtg): the second
// factor is static and precomputed once (gs: the two-phase fill applied to the sqrtg field
// itself, kept in the ghost cells of a 6-panel array).  The kernel therefore gathers only Q.
// (device functions in ghost_core.cuh)

// sum of n partials in a fixed order (every CTA gets the same bits)
__device__ double reduce_partials(const double* __restrict__ part, int n, double* sh) {
  double v = 0.0;
  for (int k = threadIdx.x; k < n; k += blockDim.x) v += part[k];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
  __syncthreads();
  return t;
}

// The whole Lagrange ghost fill in ONE launch (it sits on the per-step critical path, and on
// several GPUs between the peers' flags and the step kernel):
//   * multi-GPU: every CTA first waits for the flags of exchange `epoch` (null on one GPU);
//   * blocks [0, nb1): phase 1, one thread per edge ghost cell;
//   * last 3 blocks: the 12 x 32 corner cells of phase 2 (src/interpolation.py:250-314).  A
//     corner stencil reads the neighbour's strip, whose ends are that neighbour's phase-1
//     ghosts; instead of waiting for them they are recomputed in registers (same arithmetic,
//     same bits), so corners depend on interior cells only.
// Pending MF-PR term: ghost += corr * gs (see above); gs == nullptr computes the raw fill.
__global__ void dg_fill_fused_kernel(Geo g, HaloMaps maps, double* __restrict__ q, const int* __restrict__ kminE,
                                     const double* __restrict__ wE, int order, const double* __restrict__ gs,
                                     const double* __restrict__ part, int npart, double inv_a2,
                                     double* __restrict__ corr_out, const long long* __restrict__ flags, int world,
                                     long long epoch, int nbx, const double* __restrict__ corr_in, int* mg_err,
                                     unsigned long long mg_timeout_ns) {
  __shared__ double sh[32];
  // programmatic dependent launch (no-ops otherwise): this grid may start while the previous
  // step kernel drains; nothing of it is read before this point
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (flags) {
    if (threadIdx.x < world) {
      mg_wait_flag(flags + threadIdx.x, epoch, mg_err, mg_timeout_ns);     // bounded, see mgpu.cuh
      __threadfence_system();
    }
    __syncthreads();
  }
  double corr = 0.0;
  if (corr_in) {                       // projection coefficient already known (ring restore after a run)
    corr = *corr_in;
    npart = 1;
  } else if (npart > 0) {
    corr = -reduce_partials(part, npart, sh) * inv_a2;
    if (blockIdx.x == 0 && threadIdx.x == 0) *corr_out = corr;
  }
  const int nb1 = nbx * 4 * 24;
  if ((int)blockIdx.x < nb1) {
    const int bx = blockIdx.x % nbx, gl = (blockIdx.x / nbx) & 3, ps = blockIdx.x / (4 * nbx);
    const int k = g.lo + bx * blockDim.x + threadIdx.x;
    if (k >= g.hi) return;
    const int p = ps >> 2, s = ps & 3;
    double acc = dg_phase1_value(g, maps, q, kminE, wE, order, p, s, gl, k);
    int i, j;
    if (s == SIDE_E) { i = g.hi + gl; j = k; }
    else if (s == SIDE_W) { i = gl; j = k; }
    else if (s == SIDE_N) { i = k; j = g.hi + gl; }
    else { i = k; j = gl; }
    const long long id = gidx(g, p, i, j);
    if (npart > 0) acc = fma(gs[id], corr, acc);
    q[id] = acc;
    return;
  }
  // corners: 12 (panel, E|W) x 4 layers x 8 positions
  const int t = ((int)blockIdx.x - nb1) * blockDim.x + threadIdx.x;
  if (t >= 12 * 32) return;
  const int c32 = t & 31, pe = t >> 5;
  const int gl = c32 >> 3, c = c32 & 7;
  const int k = (c < 4) ? c : g.hi + (c - 4);
  const int p = pe >> 1, s = pe & 1;
  double acc = dg_corner_value(g, maps, q, kminE, wE, order, p, s, gl, k);
  const int i = (s == SIDE_E) ? g.hi + gl : gl;
  const long long id = gidx(g, p, i, k);
  if (npart > 0) acc = fma(gs[id], corr, acc);
  q[id] = acc;
}

// sqrtg of the single metric panel copied into the interior of all six panels of dst
__global__ void spread_metric_kernel(Geo g, const double* __restrict__ sgc, double* __restrict__ dst) {
  int j = g.lo + blockIdx.x * blockDim.x + threadIdx.x, i = g.lo + blockIdx.y, p = blockIdx.z;
  if (j >= g.hi) return;
  dst[gidx(g, p, i, j)] = sgc[gidx(g, 0, i, j)];
}

// add the pending projection term to the interior (before anything else reads Q)
__global__ void flush_corr_kernel(Geo g, double* __restrict__ q, const double* __restrict__ sgc,
                                  const double* __restrict__ part, int npart, double inv_a2) {
  __shared__ double sh[32];
  double corr = -reduce_partials(part, npart, sh) * inv_a2;
  int j = g.lo + blockIdx.x * blockDim.x + threadIdx.x, i = g.lo + blockIdx.y, p = blockIdx.z;
  if (j >= g.hi) return;
  long long id = gidx(g, p, i, j);
  q[id] = fma(sgc[gidx(g, 0, i, j)], corr, q[id]);
}

// ghost ring of the buffer the last step read (filled at the start of that step) ->
// the buffer it wrote, so that Q looks exactly like the reference's after the step
__global__ void copy_ring_kernel(Geo g, double* __restrict__ dst, const double* __restrict__ src) {
  int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, p = blockIdx.z;
  if (j >= g.P) return;
  if (i >= g.lo && i < g.hi && j >= g.lo && j < g.hi) return;
  long long id = gidx(g, p, i, j);
  dst[id] = src[id];
}

__global__ void recip_kernel(Geo g, const double* __restrict__ s, double* __restrict__ d) {
  int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j > g.P) return;
  long long id = gidx(g, 0, i, j);
  double v = s[id];
  d[id] = (v != 0.0) ? 1.0 / v : 0.0;
}

template <int TB, int MASK, int D>
size_t fused_smem_bytes() {
  const int narr = (MASK & 1) ? 9 : 7;
  return sizeof(double) * (size_t)(TB + 2) * (D * narr + 4) + sizeof(uint64_t) * D + 16;
}

template <int TB, int RECON, int SPLIT, int MASK, int D>
cudaError_t launch_kernel(const FusedArgs& a, int nblocks, cudaStream_t st) {
  static bool configured = false;
  const size_t smem = fused_smem_bytes<TB, MASK, D>();
  auto kern = fused_step_kernel<TB, RECON, SPLIT, MASK, D>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  kern<<<nblocks, TB, smem, st>>>(a);
  return cudaSuccess;
}

template <int TB, int RECON, int SPLIT, int D>
cudaError_t launch_variant(const FusedArgs& a, int mask, int nblocks, cudaStream_t st) {
  if (mask == 1) return launch_kernel<TB, RECON, SPLIT, 1, D>(a, nblocks, st);
  if (mask == 2) return launch_kernel<TB, RECON, SPLIT, 2, D>(a, nblocks, st);
  return launch_kernel<TB, RECON, SPLIT, 0, D>(a, nblocks, st);
}

// The par-default scheme (PPM-PL07 / SP-AVLT) is instantiated for every tuning point
// (threads per CTA x ring depth); the other scheme tuples use 160 threads, depth 6.
cudaError_t launch_fused(const FusedArgs& a, int recon, int split, int mask, int nblocks, int tb, int depth,
                         cudaStream_t st) {
  if (recon == 3 && split == 1) {
#define TUNE(T, DD) \
  if (tb == T && depth == DD) return launch_variant<T, 3, 1, DD>(a, mask, nblocks, st)
    TUNE(128, 5); TUNE(128, 6); TUNE(128, 7);
    TUNE(160, 5); TUNE(160, 6); TUNE(160, 7);
    TUNE(256, 5); TUNE(256, 6);
#undef TUNE
    return launch_variant<160, 3, 1, 6>(a, mask, nblocks, st);
  }
#define CASE(R, S) \
  if (recon == R && split == S) return launch_variant<160, R, S, 6>(a, mask, nblocks, st)
  CASE(3, 2); CASE(3, 3);
  CASE(1, 1); CASE(1, 2); CASE(1, 3);
  CASE(2, 1); CASE(2, 2); CASE(2, 3);
  CASE(4, 1); CASE(4, 2); CASE(4, 3);
#undef CASE
  return cudaErrorInvalidValue;
}

}  // namespace

// fused-path state kept next to the handle (one per handle, keyed by pointer)
struct FusedState {
  double* rgc = nullptr;       // 1/sqrtg_pc
  double* gs = nullptr;        // ghost cells: Lagrange fill of the sqrtg field (static)
  double* part = nullptr;
  unsigned* counter = nullptr; // last-writer ticket of the step kernels
  int ghost_fused = 0;         // PYCS_GHOST_FUSED=1: v2b computes the ghost cells itself (single GPU)
  int last_pend = 0;           // the last step applied a pending projection term (for the ring restore)
  int pdl = 0;                 // PYCS_PDL=1: ghost fill and step kernel launched with programmatic stream serialization
  int mg_fused = 0;            // multi-GPU: 1 = v2b stores to the peers itself (PYCS_MG_FUSED=1; measured slower), 0 = exchange kernel
  int prof = -1;               // PYCS_STEP_PROFILE: CUDA events around the kernels of every step
  std::vector<cudaEvent_t> ev; // 4 per profiled step: start, after ghost fill, after step kernel, after exchange
  int npart_cap = 0;
  double* bu = nullptr;        // separable wind: ucontra(t = 0) incl. ghost edges
  double* bv = nullptr;        //                 vcontra(t = 0)
  int base_valid = 0;
  int pending = 0;             // partials of the last step wait to be applied
  int ring_pending = 0;        // ghost ring of the current buffer is stale
  int npart = 0;
  int tb = 160, depth = 6, rows = 0, nstrips = 0, wcols = 0, nchunks = 0;
  int impl = 0;                // 2: block-synchronous kernel (this file), 3: warp-autonomous kernel (fused3.cu)
  int nw = 3, pf = 3, minb = 3; // v3: consumer warps per CTA, rows in flight, register cap (CTAs/SM)
  // PYCS_SPLIT=1 (single GPU, default v2b march, steps without wind kernels): a step = interior CTAs on
  // the handle's stream + ghost fill and boundary CTAs on a second stream (FusedArgs::blk_map)
  int split = 0;
  int* map_i = nullptr;        // CTA indices of the interior / boundary launch
  int* map_b = nullptr;
  int n_i = 0, n_b = 0;
  cudaStream_t s2 = nullptr;
  cudaEvent_t e_fork = nullptr, e_join = nullptr;
};

static std::map<pycs_handle, FusedState> g_fused;

int k_fused_supported(pycs_handle h) {
  // duo-grid ghost cells only (ET-S72/PL07 refill ghosts between the two stages and
  // ET-PL07 couples parabolas across panels); MF-AF couples fluxes across panels.
  return (h->prm.et == 3 && h->prm.mf != 2) ? 1 : 0;
}

static int fused_setup(pycs_handle h, FusedState& fs) {
  const Geo& g = h->g;
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
    const char* ei = getenv("PYCS_FUSED_IMPL");
    int impl = ei ? atoi(ei) : 4;     // v2b: fastest so far (profiles/r1_sweep_v2b.log); 2 = v2, 3 = v3
    const char* ew = getenv("PYCS_FUSED_NW");
    const char* ed = getenv("PYCS_FUSED_DEPTH");
    const char* er = getenv("PYCS_FUSED_ROWS");
    int rows = er ? atoi(er) : 0;
    const char* ep = getenv("PYCS_FUSED_PF");
    const char* em = getenv("PYCS_FUSED_MINB");
    int nw = ew ? atoi(ew) : 3, depth = ed ? atoi(ed) : 5;
    int pf = ep ? atoi(ep) : 3, minb = em ? atoi(em) : 3;
    if (impl == 3 && !pycs_fused3_has(h->prm.recon, h->prm.opsplit, nw, pf, minb)) { nw = 3; pf = 3; minb = 3; }
    if (impl == 3 && !pycs_fused3_has(h->prm.recon, h->prm.opsplit, nw, pf, minb)) impl = 2;   // limited PPM
    const char* etb = getenv("PYCS_FUSED_TB");
    // default: 160 threads, two rows in flight, const-slot march at 4 CTAs/SM (MINB 34:
    // profiles/r1_sweep_v2b_cs.log, 0.164 ms against 0.183 ms for the shifting-window march MINB 14)
    int tb4 = etb ? atoi(etb) : 160, pf4 = ep ? atoi(ep) : 2, minb4 = em ? atoi(em) : 34;
    if (impl == 4 && !pycs_fused2b_has(h->prm.recon, h->prm.opsplit, tb4, pf4, minb4)) { tb4 = 160; pf4 = 2; minb4 = 34; }
    if (impl == 4 && !pycs_fused2b_has(h->prm.recon, h->prm.opsplit, tb4, pf4, minb4)) impl = 2;   // limited PPM
    int resident, lag, cols;
    if (impl == 3) {
      // strips of up to 57 + 58 (NW-1) columns, one consumer warp per 57/58 of them
      fs.nw = nw;
      fs.pf = pf;
      fs.minb = minb;
      const int cap = f3_strip_capacity(nw);
      fs.nstrips = (g.N + cap - 1) / cap;
      fs.wcols = (g.N + fs.nstrips - 1) / fs.nstrips;
      fs.wcols += fs.wcols & 1;                  // even: column pairs stay 16-byte aligned
      const int mask = (h->prm.dp == 2) ? 1 : 0;
      int per_sm = pycs_fused3_resident(h->prm.recon, h->prm.opsplit, mask, nw, pf, minb);
      if (per_sm < 1) {
        pycs_set_error("fused3 kernel: occupancy query failed");
        return PYCS_ERR_CUDA;
      }
      resident = h->sm_count * per_sm;
      lag = 6;
    } else if (impl == 4) {
      // v2b: strips of TB-6 columns, one column per thread
      fs.tb = tb4;
      fs.pf = pf4;
      fs.minb = minb4;
      const int wmax = tb4 - 6;
      fs.nstrips = (g.N + wmax - 1) / wmax;
      fs.wcols = (g.N + fs.nstrips - 1) / fs.nstrips;
      fs.wcols += fs.wcols & 1;
      const int mask = (h->prm.dp == 2) ? 1 : 0;
      int per_sm = pycs_fused2b_resident(h->prm.recon, h->prm.opsplit, mask, tb4, pf4, minb4);
      if (per_sm < 1) {
        pycs_set_error("fused2b kernel: occupancy query failed");
        return PYCS_ERR_CUDA;
      }
      resident = h->sm_count * per_sm;
      lag = 6;
    } else {
      // strips of TB-6 columns; chunks sized so that the grid fills whole waves
      const char* e = getenv("PYCS_FUSED_TB");
      int tb = e ? atoi(e) : 160;
      if (tb != 128 && tb != 160 && tb != 256) tb = 160;
      if (depth < 5 || depth > 7 || (tb == 256 && depth > 6)) depth = 5;
      if (!(h->prm.recon == 3 && h->prm.opsplit == 1)) { tb = 160; depth = 6; }
      fs.tb = tb;
      fs.depth = depth;
      int wmax = tb - 6;
      fs.nstrips = (g.N + wmax - 1) / wmax;
      fs.wcols = (g.N + fs.nstrips - 1) / fs.nstrips;
      fs.wcols += fs.wcols & 1;                  // even: staged rows start 16-byte aligned
      resident = h->sm_count * (depth == 5 ? 4 : 3);
      lag = 4;
    }
    fs.impl = impl;
    if (const char* emf = getenv("PYCS_MG_FUSED")) fs.mg_fused = atoi(emf);
    if (const char* epd = getenv("PYCS_PDL")) fs.pdl = atoi(epd);
    if (const char* egf = getenv("PYCS_GHOST_FUSED")) fs.ghost_fused = atoi(egf);
    cols = 6 * fs.nstrips;
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
  }
  int nb = 6 * fs.nstrips * fs.nchunks * (fs.impl == 3 ? fs.nw : 1);   // MF-PR partial sums
  if (fs.npart_cap < nb) {
    if (fs.part) cudaFree(fs.part);
    CK(cudaMalloc(&fs.part, sizeof(double) * nb));
    fs.npart_cap = nb;
  }
  fs.npart = nb;
  if (fs.split == 0) {
    const char* es = getenv("PYCS_SPLIT");
    fs.split = (es && atoi(es) && fs.impl == 4 && fs.tb == 160 && fs.pf == 2 && fs.minb == 34) ? 1 : -1;
    if (fs.split == 1) {
      std::vector<int> in(nb), bd(nb);
      fs.n_i = pycs_split_sets(fs.nstrips, fs.nchunks, in.data(), bd.data());
      fs.n_b = nb - fs.n_i;
      if (fs.n_i == 0) {
        fs.split = -1;           // too few strips / chunks: nothing is ghost-free
      } else {
        CK(cudaMalloc(&fs.map_i, sizeof(int) * fs.n_i));
        CK(cudaMalloc(&fs.map_b, sizeof(int) * fs.n_b));
        CK(cudaMemcpy(fs.map_i, in.data(), sizeof(int) * fs.n_i, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(fs.map_b, bd.data(), sizeof(int) * fs.n_b, cudaMemcpyHostToDevice));
        if (!fs.s2) {
          CK(cudaStreamCreateWithFlags(&fs.s2, cudaStreamNonBlocking));
          CK(cudaEventCreateWithFlags(&fs.e_fork, cudaEventDisableTiming));
          CK(cudaEventCreateWithFlags(&fs.e_join, cudaEventDisableTiming));
        }
      }
    }
  }
  if (!fs.counter) {
    CK(cudaMalloc(&fs.counter, sizeof(unsigned)));
    CK(cudaMemsetAsync(fs.counter, 0, sizeof(unsigned), h->stream));
  }
  return 0;
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
  fprintf(stderr, "[pycs step profile] rank %d: %zu steps; ghost fill (+flag wait) %.2f us, winds+step kernel %.2f us, "
                  "exchange %.2f us, gap to next step %.2f us\n",
          h->mg ? h->mg->rank : 0, ns - skip, 1e3 * t[0] / n, 1e3 * t[1] / n, 1e3 * t[2] / n, 1e3 * t[3] / n);
  for (auto e : fs.ev) cudaEventDestroy(e);
  fs.ev.clear();
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
  it->second.pending = 0;
  it->second.ring_pending = 0;
  return 0;
}

// apply the pending projection term to the current Q so that every other code
// path (download, operator kernels, diagnostics) sees the reference's Q
int k_fused_flush(pycs_handle h) {
  auto it = g_fused.find(h);
  if (it == g_fused.end()) return 0;
  FusedState& fs = it->second;
  const Geo& g = h->g;
  double *sgc, *q, *qo;
  TRY(pycs_field_ptr(h, PYCS_F_SQRTG_PC, &sgc));
  TRY(pycs_field_ptr(h, h->qcur ? PYCS_F_Q_NEXT : PYCS_F_Q, &q));
  TRY(pycs_field_ptr(h, h->qcur ? PYCS_F_Q : PYCS_F_Q_NEXT, &qo));
  if (fs.pending) {
    const double* sums = h->red_out + 9;   // total of the last step kernel's partials
    int nsums = 1;
    if (h->mg) {                     // per-rank sums of the last exchange, once they have all arrived
      TRY(k_mg_wait(h));
      sums = k_mg_sums(h);
      nsums = h->mg->world;
    }
    flush_corr_kernel<<<dim3((g.N + 127) / 128, g.N, 6), 128, 0, h->stream>>>(g, q, sgc, sums, nsums,
                                                                             1.0 / h->a2);
    CKL(h);
    fs.pending = 0;
  }
  if (fs.ring_pending && fs.impl == 4 && fs.ghost_fused && !h->mg && h->kminE) {
    // the step kernel filled only the ghost cells it needed: rebuild the whole 4-wide ring of the
    // buffer the last step read (its interior is intact) before it is copied over
    const int nbx = (g.N + 127) / 128;
    dg_fill_fused_kernel<<<nbx * 4 * 24 + 3, 128, 0, h->stream>>>(
        g, h->maps, qo, h->kminE, h->wE, h->order, fs.last_pend ? fs.gs : nullptr, nullptr, 0, 0.0, h->red_out + 10,
        nullptr, 0, 0, nbx, fs.last_pend ? h->red_out + 8 : nullptr, nullptr, 0ull);
    CKL(h);
  }
  if (fs.ring_pending) {
    copy_ring_kernel<<<dim3((g.P + 127) / 128, g.P, 6), 128, 0, h->stream>>>(g, q, qo);
    CKL(h);
    fs.ring_pending = 0;
  }
  return 0;
}

void k_fused_release(pycs_handle h) {
  auto it = g_fused.find(h);
  if (it == g_fused.end()) return;
  if (it->second.rgc) cudaFree(it->second.rgc);
  if (it->second.gs) cudaFree(it->second.gs);
  if (it->second.part) cudaFree(it->second.part);
  if (it->second.counter) cudaFree(it->second.counter);
  if (it->second.bu) cudaFree(it->second.bu);
  if (it->second.bv) cudaFree(it->second.bv);
  if (it->second.map_i) cudaFree(it->second.map_i);
  if (it->second.map_b) cudaFree(it->second.map_b);
  if (it->second.e_fork) cudaEventDestroy(it->second.e_fork);
  if (it->second.e_join) cudaEventDestroy(it->second.e_join);
  if (it->second.s2) cudaStreamDestroy(it->second.s2);
  g_fused.erase(it);
}

// The plain Lagrange ghost fill (src/interpolation.py:154-314) of any centre field in one launch:
// same arithmetic as dg_phase1_kernel + dg_phase2_kernel of halo.cu, without the dependency
// between the two phases (corners recompute the neighbour's edge ghosts they read).
int k_dg_fill_single(pycs_handle h, double* q) {
  if (!h->kminE) {
    pycs_set_error("ET-DG ghost fill needs pycs_upload_lagrange first");
    return PYCS_ERR_STATE;
  }
  const int nbx = (h->g.N + 127) / 128;
  dg_fill_fused_kernel<<<nbx * 4 * 24 + 3, 128, 0, h->stream>>>(h->g, h->maps, q, h->kminE, h->wE, h->order, nullptr,
                                                                nullptr, 0, 0.0, h->red_out + 10, nullptr, 0, 0, nbx, nullptr, nullptr, 0ull);
  CKL(h);
  return 0;
}

// the rows this handle updates changed (pycs_mgpu_init): recompute the launch geometry
void k_fused_reset_grid(pycs_handle h) {
  FusedState& fs = g_fused[h];
  fs.rows = 0;
  if (fs.map_i) cudaFree(fs.map_i);      // CTA sets of the split step belong to the old grid
  if (fs.map_b) cudaFree(fs.map_b);
  fs.map_i = fs.map_b = nullptr;
  if (fs.split == 1) fs.split = 0;       // rebuilt for the new grid (row slabs of several GPUs) by fused_setup
}

// geometry was re-uploaded: 1/sqrtg and the t = 0 winds must be rebuilt
void k_fused_invalidate(pycs_handle h) {
  auto it = g_fused.find(h);
  if (it == g_fused.end()) return;
  if (it->second.rgc) cudaFree(it->second.rgc);
  it->second.rgc = nullptr;
  if (it->second.gs) cudaFree(it->second.gs);
  it->second.gs = nullptr;
  it->second.base_valid = 0;
}

void k_fused_invalidate_ghost_metric(pycs_handle h) {
  auto it = g_fused.find(h);
  if (it == g_fused.end()) return;
  if (it->second.gs) cudaFree(it->second.gs);
  it->second.gs = nullptr;
}

// Separable wind (vf = 3, RK1): the step kernel scales the contravariant wind of t = 0
// (interior + ghost edges, exactly what init_vars_adv leaves in ucontra_averaged,
// src/advection_vars.py:37-87) by cos(pi t / T).  The copy is private to the fused path:
// ucontra_averaged itself is overwritten by every non-separable step.  Building it
// overwrites U_pu / U_pv / U_pc; the resync before the last step of a run restores them.
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
  return 0;
}

// Bring the exposed wind state (U_pu / U_pv / U_pc arrays) to what the reference holds
// after update_adv(t_kprev), following a stretch of separable-wind steps that did not
// touch those arrays: wind(t_{kprev-1}) on the interior, the ghost fill of step kprev,
// then update_adv(t_kprev) (src/advection_timestep.py:31-37, :48-75).
int k_wind_resync(pycs_handle h, long long kprev) {
  if (kprev < 1) return 0;
  TRY(k_wind_interior(h, (double)(kprev - 1) * h->g.dt, 1, 1));
  TRY(k_wind_ghost_fill(h));
  return k_update_adv(h, (double)kprev * h->g.dt);
}

// Everything the reference's step k leaves in U_pu / U_pv / U_pc, rebuilt from the analytic wind
// after one or more separable-wind steps that did not touch those arrays: wind(t_{k-1}) on the
// interior, the ghost fill and the departure velocity of step k (src/advection_timestep.py:31-37),
// then update_adv(t_k) (:48-75).
int k_wind_catch_up(pycs_handle h, long long k) {
  if (k < 1 || h->prm.vf < 2) return 0;
  TRY(k_wind_interior(h, (double)(k - 1) * h->g.dt, 1, 1));
  TRY(k_wind_ghost_fill(h));
  TRY(k_time_averaged_velocity(h));
  return k_update_adv(h, (double)k * h->g.dt);
}

// ghost(sqrtg): the Lagrange fill applied to the metric field itself, once (see dg_fill_fused_kernel)
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
  const int nbx = (g.N + 127) / 128;
  dg_fill_fused_kernel<<<nbx * 4 * 24 + 3, 128, 0, h->stream>>>(g, h->maps, fs.gs, h->kminE, h->wE, h->order, nullptr,
                                                                nullptr, 0, 0.0, h->red_out + 10, nullptr, 0, 0, nbx,
                                                                nullptr, nullptr, 0ull);
  CKL(h);
  return 0;
}

// map / nmap / st: one launch of a split step (CTA subset on stream st); default: the whole grid on the handle's stream
static int launch_step_kernel(pycs_handle h, FusedState& fs, const double* qcur, double* qnext, int pend,
                              int mask, double ws, const int* map = nullptr, int nmap = 0, cudaStream_t st = nullptr,
                              bool wait_flags = false) {
  if (!st) st = h->stream;
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
  a.g = g;
  a.q = qcur; a.qn = qnext;
  if (mask == 2) { ua = fs.bu; va = fs.bv; }
  a.ua = ua; a.va = va; a.um = um; a.vm = vm;
  a.sgc = sgc; a.rgc = fs.rgc; a.sgu = sgu; a.sgv = sgv;
  a.part = fs.part;
  a.sum_out = h->red_out + 9;
  a.counter = fs.counter;
  a.corr = h->red_out + 8;
  a.rows_per_chunk = fs.rows; a.nstrips = fs.nstrips; a.wcols = fs.wcols;
  a.row_lo = h->row_lo; a.row_hi = h->row_hi;
  a.mg.world = 0;
  a.pdl = fs.pdl;
  a.gf.enable = 0;
  a.blk_map = map;
  a.nblk_total = fs.npart;
  a.wait_flags = nullptr;
  a.wait_world = 0;
  a.wait_epoch = 0;
  a.mg_err = nullptr;
  a.mg_timeout_ns = 0;
  if (map) {                      // the kernel forms the projection coefficient itself
    a.gf.sums = h->mg ? k_mg_sums(h) : h->red_out + 9;
    a.gf.nsums = h->mg ? h->mg->world : 1;
    a.gf.inv_a2 = pend ? 1.0 / h->a2 : 0.0;
    if (h->mg && wait_flags) {    // interior launch on several GPUs: the peers' sums arrive with their flags
      a.wait_flags = h->mg->sync->flag;
      a.wait_world = h->mg->world;
      a.wait_epoch = h->mg->epoch;
      a.mg_err = &h->mg->sync->err;
      a.mg_timeout_ns = h->mg->timeout_ns;
    }
  }
  if (fs.impl == 4 && fs.ghost_fused && !h->mg) {
    a.gf.enable = 1;
    a.gf.order = h->order;
    a.gf.kminE = h->kminE;
    a.gf.wE = h->wE;
    a.gf.gs = fs.gs;
    a.gf.sums = h->red_out + 9;
    a.gf.nsums = 1;
    a.gf.inv_a2 = pend ? 1.0 / h->a2 : 0.0;
    a.gf.corr_out = h->red_out + 8;
    a.gf.maps = h->maps;
  }
  if (h->mg && fs.impl == 4 && fs.mg_fused) TRY(k_mg_fill_args(h, qnext, &a.mg));   // exchange inside the kernel
  a.apply_corr = pend;
  a.cdx = g.dt / g.dx; a.cdy = g.dt / g.dy;
  a.ws = ws;
  if (fs.impl == 3)
    CK(pycs_launch_fused3(a, h->prm.recon, h->prm.opsplit, mask, fs.nw, fs.pf, fs.minb, fs.npart / fs.nw, h->stream));
  else if (fs.impl == 4)
    CK(pycs_launch_fused2b(a, h->prm.recon, h->prm.opsplit, mask, fs.tb, fs.pf, fs.minb, map ? nmap : fs.npart, st));
  else
    CK(launch_fused(a, h->prm.recon, h->prm.opsplit, mask, fs.npart, fs.tb, fs.depth, h->stream));
  CKL(h);
  return 0;
}

// Device time of `reps` back-to-back launches of the step kernel alone (ping-pong
// buffers, no ghost fill): the roofline measurement of bench.py.  Leaves Q undefined.
int k_fused_time_kernel(pycs_handle h, int reps, int separable, float* ms) {
  FusedState& fs = g_fused[h];
  TRY(fused_setup(h, fs));
  double *qa, *qb;
  TRY(pycs_field_ptr(h, PYCS_F_Q, &qa));
  TRY(pycs_field_ptr(h, PYCS_F_Q_NEXT, &qb));
  int mask = separable ? 2 : ((h->prm.dp == 2) ? 1 : 0);
  if (separable) TRY(ensure_base_winds(h, fs));
  TRY(ensure_gs(h, fs));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaEventRecord(h->ev0, h->stream));
  for (int r = 0; r < reps; ++r)
    TRY(launch_step_kernel(h, fs, (r & 1) ? qb : qa, (r & 1) ? qa : qb, h->prm.mf == 3 ? 1 : 0, mask, 0.999));
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaEventSynchronize(h->ev1));
  CK(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  fs.pending = 0;
  fs.ring_pending = 0;
  return 0;
}

int k_fused_kernel_name(pycs_handle h, char* out, int len) {
  FusedState& fs = g_fused[h];
  TRY(fused_setup(h, fs));
  if (fs.impl == 3)
    snprintf(out, len, "fused3_kernel<recon=%d,split=%d,NW=%d,PF=%d,MINB=%d> (csrc/fused3.cu)", h->prm.recon,
             h->prm.opsplit, fs.nw, fs.pf, fs.minb);
  else if (fs.impl == 4)
    snprintf(out, len, "fused2b_kernel<TB=%d,recon=%d,split=%d,PF=%d,MINB=%d> (csrc/fused2b.cu)", fs.tb, h->prm.recon,
             h->prm.opsplit, fs.pf, fs.minb);
  else
    snprintf(out, len, "fused_step_kernel<TB=%d,recon=%d,split=%d,D=%d> (csrc/fused.cu)", fs.tb, h->prm.recon,
             h->prm.opsplit, fs.depth);
  return 0;
}

int k_fused_grid_info(pycs_handle h, int* tb, int* rows, int* nblocks) {
  FusedState& fs = g_fused[h];
  TRY(fused_setup(h, fs));
  *tb = fs.impl == 3 ? (fs.nw + 1) * 32 : fs.tb;
  *rows = fs.rows;
  *nblocks = fs.impl == 3 ? fs.npart / fs.nw : fs.npart;
  return 0;
}

// separable != 0: wind field 3 with RK1 -- U(t) = U(0) cos(pi t / T) exactly
// (src/advection_ic.py:301-305), so the step reads the t = 0 winds (still in
// ucontra_averaged) and scales them; no wind kernels run.
int k_fused_step(pycs_handle h, long long k, double t, int separable) {
  const Geo& g = h->g;
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
  double *qa, *qb, *sgc;
  TRY(pycs_field_ptr(h, PYCS_F_Q, &qa));
  TRY(pycs_field_ptr(h, PYCS_F_Q_NEXT, &qb));
  TRY(pycs_field_ptr(h, PYCS_F_SQRTG_PC, &sgc));
  double* qcur = h->qcur ? qb : qa;
  double* qnext = h->qcur ? qa : qb;
  if (separable) TRY(ensure_base_winds(h, fs));
  TRY(ensure_gs(h, fs));

  const double* sums = h->red_out + 9;     // total of the last step kernel's partials
  int nsums = 1;
  const long long* mgflags = nullptr;
  int mgworld = 0;
  long long mgepoch = 0;
  int* mgerr = nullptr;
  unsigned long long mgtimeout = 0;
  if (h->mg) {
    mgerr = &h->mg->sync->err;
    mgtimeout = h->mg->timeout_ns;                       // the wait is folded into the ghost-fill kernel
    sums = k_mg_sums(h);
    nsums = h->mg->world;
    mgflags = h->mg->sync->flag;
    mgworld = h->mg->world;
    mgepoch = h->mg->epoch;
  }
  if (fs.prof < 0) fs.prof = getenv("PYCS_STEP_PROFILE") ? 1 : 0;
  auto mark = [&]() {
    if (!fs.prof || fs.ev.size() >= 4 * 4096) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, h->stream);
    fs.ev.push_back(e);
  };
  mark();
  // 1. ghost cells of Q (src/advection_timestep.py:28), folding in the pending MF-PR term
  int pend = fs.pending;
  const bool ghost_in_kernel = fs.impl == 4 && fs.ghost_fused && !h->mg;
  // split step: no wind kernel may sit between the ghost fill and the step kernel
  const bool split = fs.split == 1 && !ghost_in_kernel && !fs.pdl && !(h->mg && fs.mg_fused) &&
                     (h->prm.vf < 2 || separable);
  if (split) {
    const int mask = separable ? 2 : ((h->prm.dp == 2) ? 1 : 0);
    const double ws = separable ? cos(PYCS_PI * ((double)(k - 1) * g.dt) / PYCS_WIND_PERIOD) : 1.0;
    const int nbx = (g.N + 127) / 128;
    CK(cudaEventRecord(fs.e_fork, h->stream));            // everything before this step
    CK(cudaStreamWaitEvent(fs.s2, fs.e_fork, 0));
    mark();                                               // (four marks per step: the ghost fill has no interval of its own here)
    dg_fill_fused_kernel<<<nbx * 4 * 24 + 3, 128, 0, fs.s2>>>(g, h->maps, qcur, h->kminE, h->wE, h->order, fs.gs, sums,
                                                             pend ? nsums : 0, pend ? 1.0 / h->a2 : 0.0, h->red_out + 8,
                                                             mgflags, mgworld, mgepoch, nbx, nullptr, mgerr, mgtimeout);
    CKL(h);
    TRY(launch_step_kernel(h, fs, qcur, qnext, pend, mask, ws, fs.map_b, fs.n_b, fs.s2));     // needs the ghost cells
    CK(cudaEventRecord(fs.e_join, fs.s2));
    TRY(launch_step_kernel(h, fs, qcur, qnext, pend, mask, ws, fs.map_i, fs.n_i, h->stream, true)); // reads no ghost cell
    CK(cudaStreamWaitEvent(h->stream, fs.e_join, 0));
    mark();
    if (h->mg) TRY(k_mg_exchange(h, qnext, h->red_out + 9, 1));   // after both launches: boundary cells + the sum
    mark();
    h->last_step_kernel_launches++;
    h->qcur ^= 1;
    fs.last_pend = pend;
    fs.pending = (h->prm.mf == 3) ? 1 : 0;
    fs.ring_pending = 1;
    return 0;
  }
  if (!ghost_in_kernel) {
    const int nbx = (g.N + 127) / 128;
    if (fs.pdl && fs.impl == 4) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(nbx * 4 * 24 + 3);
      cfg.blockDim = dim3(128);
      cfg.stream = h->stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      CK(cudaLaunchKernelEx(&cfg, dg_fill_fused_kernel, g, h->maps, qcur, (const int*)h->kminE, (const double*)h->wE,
                            h->order, (const double*)fs.gs, sums, pend ? nsums : 0, pend ? 1.0 / h->a2 : 0.0,
                            h->red_out + 8, mgflags, mgworld, mgepoch, nbx, (const double*)nullptr, mgerr, mgtimeout));
    } else {
      dg_fill_fused_kernel<<<nbx * 4 * 24 + 3, 128, 0, h->stream>>>(
          g, h->maps, qcur, h->kminE, h->wE, h->order, fs.gs, sums, pend ? nsums : 0, pend ? 1.0 / h->a2 : 0.0,
          h->red_out + 8, mgflags, mgworld, mgepoch, nbx, nullptr, mgerr, mgtimeout);
    }
    CKL(h);
  }
  mark();
  // 2. winds (src/advection_timestep.py:31-37)
  if (h->prm.vf >= 2 && !separable) {
    TRY(k_wind_ghost_fill(h));
    TRY(k_time_averaged_velocity(h));
  }
  // 3. divergence + Q update
  int mask = separable ? 2 : ((h->prm.dp == 2) ? 1 : 0);    // RK1: averaged wind == instantaneous wind
  double ws = separable ? cos(PYCS_PI * ((double)(k - 1) * g.dt) / PYCS_WIND_PERIOD) : 1.0;
  TRY(launch_step_kernel(h, fs, qcur, qnext, pend, mask, ws));
  mark();
  if (h->mg && !(fs.impl == 4 && fs.mg_fused)) TRY(k_mg_exchange(h, qnext, h->red_out + 9, 1));
  mark();
  h->last_step_kernel_launches++;
  h->qcur ^= 1;
  fs.last_pend = pend;
  fs.pending = (h->prm.mf == 3) ? 1 : 0;
  fs.ring_pending = 1;
  // 4. wind refresh for the next step (src/advection_timestep.py:48-75)
  if (h->prm.vf >= 2 && !separable) TRY(k_update_adv(h, t));
  return 0;
}
