// Fused advection step for the duo-grid (ET-DG) schemes: the production path.
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
#include "pycs_common.cuh"

namespace {

struct FusedArgs {
  Geo g;
  const double* q;
  double* qn;
  const double *ua, *va;      // U_pu.ucontra_averaged, U_pv.vcontra_averaged
  const double *um, *vm;      // mask sources (U_pu.ucontra, U_pv.vcontra)
  const double *sgc, *rgc, *sgu, *sgv;
  double* part;
  const double* corr;         // device scalar: pending projection coefficient -sum(s)/a2
  int rows_per_chunk, nstrips, wcols, apply_corr;
  double cdx, cdy;            // dt/dx, dt/dy
};

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

template <int TB, int RECON, int SPLIT, int MASK>
__global__ void __launch_bounds__(TB) fused_step_kernel(FusedArgs a) {
  constexpr int MT = (SPLIT == 3) ? 2 : 1;
  __shared__ double sQ[TB], sX[TB], sF[TB], sG[TB], sC[TB];
  const Geo& g = a.g;
  int b = blockIdx.x;
  const int p = b % 6;
  b /= 6;
  const int strip = b % a.nstrips, chunk = b / a.nstrips;
  const int tid = threadIdx.x;
  const int jbase = g.lo + strip * a.wcols;
  const int jend = min(jbase + a.wcols, g.hi);
  const int j = jbase - 3 + tid;
  const int jc = min(j, g.P - 1);
  const int r0 = g.lo + chunk * a.rows_per_chunk;
  const int r1 = min(r0 + a.rows_per_chunk, g.hi);
  const bool out_lane = (tid >= 3) && (j < jend);
  const bool jint = (jc >= g.lo) && (jc < g.hi);
  const long long L = g.ld;
  const long long col = PYCS_JOFF + jc;
  const double* __restrict__ Q = a.q + (long long)p * g.ps + col;
  double* __restrict__ QN = a.qn + (long long)p * g.ps + col;
  const double* __restrict__ UA = a.ua + (long long)p * g.ps + col;
  const double* __restrict__ VA = a.va + (long long)p * g.ps + col;
  const double* __restrict__ UM = a.um + (long long)p * g.ps + col;
  const double* __restrict__ VM = a.vm + (long long)p * g.ps + col;
  const double* __restrict__ SGC = a.sgc + col;
  const double* __restrict__ RGC = a.rgc + col;
  const double* __restrict__ SGU = a.sgu + col;
  const double* __restrict__ SGV = a.sgv + col;
  const double corr = a.apply_corr ? *a.corr : 0.0;
  const double cdx = a.cdx, cdy = a.cdy;

  // rolling state, all for column j
  double q0 = 0, q1 = 0, q2 = 0, q3 = 0, q4 = 0;       // Q rows r-4 .. r
  double y0 = 0, y1 = 0, y2 = 0, y3 = 0, y4 = 0;       // Qy rows r-4 .. r
  double pqR = 0, pq6 = 0, pdq = 0;                    // parabola of Q, cell r-3 (u >= 0 side)
  double pyR = 0, py6 = 0, pyd = 0;                    // parabola of Qy, cell r-3
  double fin_prev = 0, fout_prev = 0;                  // x-fluxes at edge r-3
  double cm_prev = 0;                                  // sqrtg_pu * cx at edge r-3 (SPLIT != 1)
  double psum = 0;

  for (int r = r0 - 3; r <= r1 + 2; ++r) {
    const long long ro = (long long)r * L;
    q0 = q1; q1 = q2; q2 = q3; q3 = q4;
    y0 = y1; y1 = y2; y2 = y3; y3 = y4;
    {
      double v = Q[ro];
      if (a.apply_corr && jint && r >= g.lo && r < g.hi) v = fma(SGC[ro], corr, v);
      q4 = v;
    }
    const bool have_cell = (r >= r0 + 1);   // cell r-2 has its 5 rows (r-4 >= r0-3)
    const bool have_edge = (r >= r0 + 2);   // edge r-2: cells r-3 and r-2 both have parabolas
    const bool outp = (r >= r0 + 3);        // output row r-3

    // ---------------- phase 1: inner x-flux at edge r-2 (registers), Qx row r-3
    double nqL = 0, nqR = 0, nq6 = 0, ndq = 0;          // parabola of Q, cell r-2
    double u = 0, cxe = 0, sup = 1.0, gue = 0;
    bool up = true;
    double fin = 0;
    if (have_cell) {
      const long long co = ro - 2 * L;
      double gl = SGU[co], gr = SGU[co + L], gcc = SGC[co];
      parabola<RECON, MT>(q0, q1, q2, q3, q4, gl, gr, gcc, nqL, nqR, nq6, ndq);
      if (have_edge) {
        u = UA[co];
        double um = MASK ? UM[co] : u;
        up = um >= 0;
        sup = up ? 1.0 : -1.0;
        cxe = u * cdx;
        gue = gl;                                   // sqrtg_pu at edge r-2
        double e = up ? pqR : nqL, s6 = up ? pq6 : nq6, sd = up ? pdq : ndq;
        fin = ppm_flux(e, s6, sd, sup, cxe) * u;
        if (MT == 2) fin *= gue;
      }
    }
    double qx = 0;
    if (outp) {
      double dFx = -(fin - fin_prev) * cdx;
      double rg = RGC[ro - 3 * L];
      if (SPLIT == 1) qx = fma(0.5 * dFx, rg, q1);
      else {
        double cd = gue * cxe - cm_prev;            // c1x - c2x
        if (SPLIT == 2) qx = fma(0.5 * fma(cd, q1, dFx), rg, q1);
        else qx = 0.5 * (q1 + (q1 + dFx) / (1.0 - cd));
      }
    }
    sQ[tid] = q4;
    sX[tid] = qx;
    __syncthreads();

    // ---------------- phase 2: y-fluxes at edge j: inner on row r, outer on row r-3
    double v_r = VA[ro];
    double gvo = 0;
    {
      double vm = MASK ? VM[ro] : v_r;
      bool vp = vm >= 0;
      int cu = vp ? tid - 1 : tid;
      cu = max(2, min(cu, TB - 3));
      long long mo = ro + (cu - tid);
      double gl = SGV[mo], gr = SGV[mo + 1], gcc = SGC[mo];
      double eL, eR, s6, sd;
      parabola<RECON, MT>(sQ[cu - 2], sQ[cu - 1], sQ[cu], sQ[cu + 1], sQ[cu + 2], gl, gr, gcc, eL, eR, s6, sd);
      double cy = v_r * cdy;
      double f = ppm_flux(vp ? eR : eL, s6, sd, vp ? 1.0 : -1.0, cy) * v_r;
      if (MT == 2 || SPLIT != 1) gvo = SGV[ro];
      if (MT == 2) f *= gvo;
      sF[tid] = f;
      if (SPLIT != 1) sC[tid] = gvo * cy;
    }
    if (outp) {
      const long long r3 = ro - 3 * L;
      double v3 = VA[r3];
      double vm = MASK ? VM[r3] : v3;
      bool vp = vm >= 0;
      int cu = vp ? tid - 1 : tid;
      cu = max(2, min(cu, TB - 3));
      long long mo = r3 + (cu - tid);
      double gl = SGV[mo], gr = SGV[mo + 1], gcc = SGC[mo];
      double eL, eR, s6, sd;
      parabola<RECON, MT>(sX[cu - 2], sX[cu - 1], sX[cu], sX[cu + 1], sX[cu + 2], gl, gr, gcc, eL, eR, s6, sd);
      double cy = v3 * cdy;
      double f = ppm_flux(vp ? eR : eL, s6, sd, vp ? 1.0 : -1.0, cy) * v3;
      if (MT == 2) f *= SGV[r3];
      sG[tid] = f;
    }
    __syncthreads();

    // ---------------- phase 3: Qy row r, outer x-flux at edge r-2 on Qy, output row r-3
    {
      int tn = min(tid + 1, TB - 1);
      double dFy = -(sF[tn] - sF[tid]) * cdy;
      double rg = RGC[ro];
      if (SPLIT == 1) y4 = fma(0.5 * dFy, rg, q4);
      else {
        double cd = sC[tn] - sC[tid];               // c1y - c2y
        if (SPLIT == 2) y4 = fma(0.5 * fma(cd, q4, dFy), rg, q4);
        else y4 = 0.5 * (q4 + (q4 + dFy) / (1.0 - cd));
      }
    }
    double fout = 0;
    if (have_cell) {
      const long long co = ro - 2 * L;
      double gl = SGU[co], gr = SGU[co + L], gcc = SGC[co];
      double nyL, nyR, ny6, nyd;
      parabola<RECON, MT>(y0, y1, y2, y3, y4, gl, gr, gcc, nyL, nyR, ny6, nyd);
      if (have_edge) {
        double e = up ? pyR : nyL, s6 = up ? py6 : ny6, sd = up ? pyd : nyd;
        fout = ppm_flux(e, s6, sd, sup, cxe) * u;
        if (MT == 2) fout *= gue;
      }
      pyR = nyR; py6 = ny6; pyd = nyd;
    }
    if (outp) {
      int tn = min(tid + 1, TB - 1);
      double s = -(fout - fout_prev) * cdx - (sG[tn] - sG[tid]) * cdy;   // pxdF + pydF
      const long long r3 = ro - 3 * L;
      if (out_lane) {
        // Q - dt*div with div = -(pxdF+pydF)/(dt*sqrtg)  (src/discrete_operators.py:95,
        // src/advection_timestep.py:43)
        QN[r3] = fma(s, RGC[r3], q1);
        psum += s;
      }
    }
    pqR = nqR; pq6 = nq6; pdq = ndq;
    fin_prev = fin;
    fout_prev = fout;
    if (SPLIT != 1) cm_prev = gue * cxe;
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
  }
}

// ---- Lagrange ghost fill that honours the pending projection term ------------
// Same arithmetic as dg_phase1_kernel in halo.cu (src/interpolation.py:200-248);
// source cells are interior cells of the neighbour, which still miss sqrtg*corr.
__device__ __forceinline__ double halo_src(const double* __restrict__ q, const double* __restrict__ sgc,
                                           const Geo& g, const SideMap& m, int a, int b, double corr) {
  int i = m.ci + m.ai * a + m.bi * b, j = m.cj + m.aj * a + m.bj * b;
  return fma(sgc[gidx(g, 0, i, j)], corr, q[gidx(g, m.nb, i, j)]);
}

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

__global__ void dg_phase1_corr_kernel(Geo g, HaloMaps maps, double* __restrict__ q,
                                      const int* __restrict__ kminE, const double* __restrict__ wE, int order,
                                      const double* __restrict__ sgc, const double* __restrict__ part,
                                      int npart, double inv_a2, double* __restrict__ corr_out) {
  __shared__ double sh[32];
  double corr = 0.0;
  if (npart > 0) {
    corr = -reduce_partials(part, npart, sh) * inv_a2;
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) *corr_out = corr;
  }
  int k = g.lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= g.hi) return;
  int gl = blockIdx.y;
  int p = blockIdx.z >> 2, s = blockIdx.z & 3;
  const SideMap& m = maps.m[p][s];
  int ge = (s == SIDE_E || s == SIDE_N) ? gl : PYCS_NG - 1 - gl;
  int km = kminE[ge * g.P + k];
  const double* w = wE + ((long long)ge * g.P + k) * order;
  double acc = 0.0;
  for (int l = 0; l < order; ++l) {
    double v = (s < 2) ? halo_src(q, sgc, g, m, gl, km + l, corr) : halo_src(q, sgc, g, m, km + l, gl, corr);
    acc = __dadd_rn(acc, __dmul_rn(v, w[l]));
  }
  int i, j;
  if (s == SIDE_E) { i = g.hi + gl; j = k; }
  else if (s == SIDE_W) { i = gl; j = k; }
  else if (s == SIDE_N) { i = k; j = g.hi + gl; }
  else { i = k; j = gl; }
  q[gidx(g, p, i, j)] = acc;
}

// Corner stencils straddle the interior / ghost boundary of the neighbour strip:
// interior sources still miss the pending term, ghost sources already have it.
__global__ void dg_phase2_corr_kernel(Geo g, HaloMaps maps, double* __restrict__ q,
                                      const int* __restrict__ kminE, const double* __restrict__ wE, int order,
                                      const double* __restrict__ sgc, const double* __restrict__ corr_p,
                                      int apply_corr) {
  const double corr = apply_corr ? *corr_p : 0.0;
  int t = threadIdx.x;
  int gl = t >> 3, c = t & 7;
  int k = (c < 4) ? c : g.hi + (c - 4);
  int p = blockIdx.x >> 1, s = blockIdx.x & 1;
  const SideMap& m = maps.m[p][s];
  int ge = (s == SIDE_E) ? gl : PYCS_NG - 1 - gl;
  int km = kminE[ge * g.P + k];
  const double* w = wE + ((long long)ge * g.P + k) * order;
  double acc = 0.0;
  for (int l = 0; l < order; ++l) {
    int a_ = gl, b_ = km + l;
    int si = m.ci + m.ai * a_ + m.bi * b_, sj = m.cj + m.aj * a_ + m.bj * b_;
    double v = q[gidx(g, m.nb, si, sj)];
    if (si >= g.lo && si < g.hi && sj >= g.lo && sj < g.hi) v = fma(sgc[gidx(g, 0, si, sj)], corr, v);
    acc = __dadd_rn(acc, __dmul_rn(v, w[l]));
  }
  int i = (s == SIDE_E) ? g.hi + gl : gl;
  q[gidx(g, p, i, k)] = acc;
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

template <int TB, int RECON, int SPLIT>
void launch_variant(const FusedArgs& a, int mask, int nblocks, cudaStream_t st) {
  if (mask) fused_step_kernel<TB, RECON, SPLIT, 1><<<nblocks, TB, 0, st>>>(a);
  else fused_step_kernel<TB, RECON, SPLIT, 0><<<nblocks, TB, 0, st>>>(a);
}

template <int TB>
void launch_fused(const FusedArgs& a, int recon, int split, int mask, int nblocks, cudaStream_t st) {
#define CASE(R, S) \
  if (recon == R && split == S) return launch_variant<TB, R, S>(a, mask, nblocks, st)
  CASE(3, 1); CASE(3, 2); CASE(3, 3);
  CASE(1, 1); CASE(1, 2); CASE(1, 3);
  CASE(2, 1); CASE(2, 2); CASE(2, 3);
  CASE(4, 1); CASE(4, 2); CASE(4, 3);
#undef CASE
}

}  // namespace

// fused-path state kept next to the handle (one per handle, keyed by pointer)
struct FusedState {
  double* rgc = nullptr;       // 1/sqrtg_pc
  double* part = nullptr;
  int npart_cap = 0;
  int pending = 0;             // partials of the last step wait to be applied
  int ring_pending = 0;        // ghost ring of the current buffer is stale
  int npart = 0;
  int tb = 160, rows = 0, nstrips = 0, wcols = 0, nchunks = 0;
};

#include <map>
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
    // strips of TB-6 columns; chunks sized so that the grid fills whole waves
    const char* e = getenv("PYCS_FUSED_TB");
    int tb = e ? atoi(e) : 160;
    if (tb != 128 && tb != 160 && tb != 192 && tb != 256) tb = 160;
    fs.tb = tb;
    int wmax = tb - 6;
    fs.nstrips = (g.N + wmax - 1) / wmax;
    fs.wcols = (g.N + fs.nstrips - 1) / fs.nstrips;
    const char* er = getenv("PYCS_FUSED_ROWS");
    int rows = er ? atoi(er) : 0;
    if (rows <= 0) {
      // aim at an integer number of waves of resident CTAs (4 CTAs/SM at <=96 regs)
      int resident = h->sm_count * 4;
      int cols = 6 * fs.nstrips;
      int best = 0;
      double best_cost = 1e30;
      for (int nch = 1; nch <= g.N; ++nch) {
        int rr = (g.N + nch - 1) / nch;
        if (rr < 8 && nch > 1) break;
        int nb = cols * ((g.N + rr - 1) / rr);
        int waves = (nb + resident - 1) / resident;
        double cost = (double)waves * (rr + 4.0);      // time ~ waves * (rows + ramp)
        if (cost < best_cost) { best_cost = cost; best = rr; }
      }
      rows = best;
    }
    fs.rows = rows;
    fs.nchunks = (g.N + rows - 1) / rows;
  }
  int nb = 6 * fs.nstrips * fs.nchunks;
  if (fs.npart_cap < nb) {
    if (fs.part) cudaFree(fs.part);
    CK(cudaMalloc(&fs.part, sizeof(double) * nb));
    fs.npart_cap = nb;
  }
  fs.npart = nb;
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
    flush_corr_kernel<<<dim3((g.N + 127) / 128, g.N, 6), 128, 0, h->stream>>>(g, q, sgc, fs.part, fs.npart,
                                                                             1.0 / h->a2);
    CKL(h);
    fs.pending = 0;
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
  if (it->second.part) cudaFree(it->second.part);
  g_fused.erase(it);
}

int k_fused_step(pycs_handle h, long long k, double t) {
  (void)k;
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
  double *qa, *qb, *sgc, *sgu, *sgv, *ua, *va, *um, *vm;
  TRY(pycs_field_ptr(h, PYCS_F_Q, &qa));
  TRY(pycs_field_ptr(h, PYCS_F_Q_NEXT, &qb));
  TRY(pycs_field_ptr(h, PYCS_F_SQRTG_PC, &sgc));
  TRY(pycs_field_ptr(h, PYCS_F_SQRTG_PU, &sgu));
  TRY(pycs_field_ptr(h, PYCS_F_SQRTG_PV, &sgv));
  TRY(pycs_field_ptr(h, PYCS_F_PU_UAVG, &ua));
  TRY(pycs_field_ptr(h, PYCS_F_PV_VAVG, &va));
  TRY(pycs_field_ptr(h, PYCS_F_PU_UCONTRA, &um));
  TRY(pycs_field_ptr(h, PYCS_F_PV_VCONTRA, &vm));
  double* qcur = h->qcur ? qb : qa;
  double* qnext = h->qcur ? qa : qb;

  // 1. ghost cells of Q (src/advection_timestep.py:28), folding in the pending MF-PR term
  int pend = fs.pending;
  dg_phase1_corr_kernel<<<dim3((g.N + 127) / 128, 4, 24), 128, 0, h->stream>>>(
      g, h->maps, qcur, h->kminE, h->wE, h->order, sgc, fs.part, pend ? fs.npart : 0,
      pend ? 1.0 / h->a2 : 0.0, h->red_out + 8);
  CKL(h);
  dg_phase2_corr_kernel<<<12, 32, 0, h->stream>>>(g, h->maps, qcur, h->kminE, h->wE, h->order, sgc,
                                                  h->red_out + 8, pend);
  CKL(h);
  // 2. winds (src/advection_timestep.py:31-37)
  if (h->prm.vf >= 2) {
    TRY(k_wind_ghost_fill(h));
    TRY(k_time_averaged_velocity(h));
  }
  // 3. divergence + Q update
  FusedArgs a;
  a.g = g;
  a.q = qcur; a.qn = qnext;
  a.ua = ua; a.va = va; a.um = um; a.vm = vm;
  a.sgc = sgc; a.rgc = fs.rgc; a.sgu = sgu; a.sgv = sgv;
  a.part = fs.part;
  a.corr = h->red_out + 8;
  a.rows_per_chunk = fs.rows; a.nstrips = fs.nstrips; a.wcols = fs.wcols;
  a.apply_corr = pend;
  a.cdx = g.dt / g.dx; a.cdy = g.dt / g.dy;
  int mask = (h->prm.dp == 2) ? 1 : 0;    // RK1: averaged wind == instantaneous wind
  int nb = fs.npart;
  switch (fs.tb) {
    case 128: launch_fused<128>(a, h->prm.recon, h->prm.opsplit, mask, nb, h->stream); break;
    case 192: launch_fused<192>(a, h->prm.recon, h->prm.opsplit, mask, nb, h->stream); break;
    case 256: launch_fused<256>(a, h->prm.recon, h->prm.opsplit, mask, nb, h->stream); break;
    default: launch_fused<160>(a, h->prm.recon, h->prm.opsplit, mask, nb, h->stream); break;
  }
  CKL(h);
  h->last_step_kernel_launches++;
  h->qcur ^= 1;
  fs.pending = (h->prm.mf == 3) ? 1 : 0;
  fs.ring_pending = 1;
  // 4. wind refresh for the next step (src/advection_timestep.py:48-75)
  if (h->prm.vf >= 2) TRY(k_update_adv(h, t));
  return 0;
}
