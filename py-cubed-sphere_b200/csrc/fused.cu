// Fused advection step, first generation (v2): one column per thread, 160 threads, two block barriers per
// marched row, uniform TMA ring.  The production kernel is v2b (fused2b.cu: same march, leaner arithmetic,
// const-slot ring); v2 remains because it carries the limited reconstructions PPM-CW84 / PPM-L04
// (src/reconstruction_1d.py:64-192), which v2b's weight-form core does not.  Single GPU, serial path only
// (stepper.cu): the ghost-fill kernel folds the pending MF-PR term into the ghost cells and hands the
// coefficient over in *corr_ptr.
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
#include <cstdio>
#include "pycs_common.cuh"
#include "fused_args.cuh"

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
  const long long step = *((const volatile long long*)&a.ctl->steps);
  const double corr = (a.apply_corr && a.corr_ptr) ? *a.corr_ptr : 0.0;
  const double ws = (MASK & 2) ? a.ws_tab[step & a.ws_mask] : 1.0;
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
        if (MASK & 2) u *= ws;
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
      if (MASK & 2) v *= ws;
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
      if (MASK & 2) v3 *= ws;
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
      *a.counter = 0u;
      if (!a.timing) {            // close the step (see StepCtl, fused_args.cuh)
        a.ctl->sum = tot;
        a.ctl->corr_applied = corr;
        a.ctl->pend = a.apply_corr;
        __threadfence();
        a.ctl->steps = step + 1;
      }
    }
  }
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

// 160 threads, ring depth 6 (3 CTAs / SM)
cudaError_t launch_v2(const FusedArgs& a, int recon, int split, int mask, int nblocks, cudaStream_t st) {
#define CASE(R, S) \
  if (recon == R && split == S) return launch_variant<160, R, S, 6>(a, mask, nblocks, st)
  CASE(2, 1); CASE(2, 2); CASE(2, 3);
  CASE(4, 1); CASE(4, 2); CASE(4, 3);
#undef CASE
  return cudaErrorInvalidValue;
}

}  // namespace

cudaError_t pycs_launch_fused_v2(const FusedArgs& a, int recon, int split, int mask, int nblocks, cudaStream_t st) {
  return launch_v2(a, recon, split, mask, nblocks, st);
}
int pycs_fused_v2_threads() { return 160; }
int pycs_fused_v2_resident() { return 3; }
