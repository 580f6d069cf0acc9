// Per-lane arithmetic of the v3 fused advection step -- shared verbatim by the CUDA
// kernel (fused3.cu) and by the host emulator used in the CPU tests
// (tests/emul/fused3_emul.cpp), so the numerics and the indexing of the kernel can be
// checked against the oracle without a GPU.
//
// Reference arithmetic (file:line under /root/reference):
//   src/reconstruction_1d.py:36-62   PPM-0 / PPM-PL07 edge values q_L, q_R
//   src/flux.py:27-70                metric weighting (MT-0 / MT-PL07), CW84 flux, f_upw = f * u
//   src/discrete_operators.py:49-73  inner update  SP-AVLT / SP-L04 / SP-PL07
//   src/discrete_operators.py:95     div = -(pxdF + pydF) / (dt sqrtg)
//   src/advection_timestep.py:43     Q -= dt * div
//
// Flux in weight form.  With the upwind cell's edge values (E on the edge itself, O on
// the cell's other edge, both metric-weighted under MT-0), G = q * sqrtg_c and b = |c|
// taken with the upwind sign (b = c if the mask says u >= 0, else -c), the reference's
//     f = E + c/2 (+-q6 - dq) - q6 c^2 / 3,   q6 = 6 G - 3 (R + L),  dq = R - L
// is identically
//     f = E (1-b)^2 - O b (1-b) + G b (3 - 2b).
// The three weights depend on the wind and the metric only, so they are computed once per
// edge and serve the inner and the outer operator; dt/dx and the final "* u" are folded
// in (flux' = f * u * dt/dx), so dF = flux'[i] - flux'[i+1].
#pragma once
#if defined(__CUDACC__)
#define PYCS_HD __host__ __device__ __forceinline__
typedef double2 dbl2;
#else
#include <cmath>
#define PYCS_HD inline
struct alignas(16) dbl2 { double x, y; };
#endif

namespace f3 {

constexpr int NC = 2;              // columns per lane
constexpr int WARP_COLS = 32 * NC; // columns marched by one consumer warp
constexpr int SXW = WARP_COLS + 8; // private Qx row: 4 pad columns on each side
// arrays staged per row in a ring slot
enum { A_Q = 0, A_V = 1, A_SGC = 2, A_SGV = 3, A_RGC = 4, A_SGU = 5, A_U = 6, A_VM = 7, A_UM = 8 };

template <int NW> struct RowWidth { static constexpr int value = 58 * NW + 12; };   // doubles per staged row

PYCS_HD dbl2 ld2(const double* p) { return *reinterpret_cast<const dbl2*>(p); }
PYCS_HD void st2(double* p, double a, double b) {
  dbl2 v; v.x = a; v.y = b;
  *reinterpret_cast<dbl2*>(p) = v;
}
PYCS_HD double pick(const dbl2& v, int c) { return c == 0 ? v.x : v.y; }

// src/reconstruction_1d.py:36-62 (q3 is the cell itself)
template <int RECON>
PYCS_HD void edge_values(double q1, double q2, double q3, double q4, double q5, double& l, double& r) {
  if (RECON == 3) {
    const double a1 = 2.0 / 60.0, a2 = -13.0 / 60.0, a3 = 47.0 / 60.0, a4 = 27.0 / 60.0, a5 = -3.0 / 60.0;
    r = fma(a5, q5, fma(a4, q4, fma(a3, q3, fma(a2, q2, a1 * q1))));
    l = fma(a1, q5, fma(a2, q4, fma(a3, q3, fma(a4, q2, a5 * q1))));
  } else {
    const double c7 = 7.0 / 12.0, c1 = 1.0 / 12.0;
    l = fma(c7, q3 + q2, -(c1 * (q4 + q1)));
    r = fma(c7, q4 + q3, -(c1 * (q5 + q2)));
  }
}

// Weights of one edge.  ubar: wind used in the flux (time averaged); up: upwind mask;
// cd = dt/dx; gE, gO, gC: sqrtg at the edge, at the other edge of the upwind cell and at
// its centre.  c_out = ubar*dt/dx (the CFL number cx of src/cfl.py:10).
// SAMEMASK: the mask is the sign of ubar itself (RK1), so b = |c| (for u = -0.0 every
// weight is zero and the side does not matter).
template <int MT, bool SAMEMASK>
PYCS_HD void edge_weights(double ubar, bool up, double cd, double gE, double gO, double gC, double& WE,
                          double& WO, double& WG, double& c_out) {
  const double c = ubar * cd;
  const double b = SAMEMASK ? fabs(c) : (up ? c : -c);
  const double m = 1.0 - b;
  const double t = fma(2.0, m, 1.0);          // 3 - 2b
  const double cs = (MT == 2) ? c * gE : c;   // MT-PL07: the whole flux times sqrtg at the edge
  const double cm = cs * m;
  const double cb = cs * b;
  const double w1 = cm * m, w2 = -(cb * m), w3 = cb * t;
  if (MT == 1) { WE = w1 * gE; WO = w2 * gO; WG = w3 * gC; }
  else { WE = w1; WO = w2; WG = w3; }
  c_out = c;
}

// inner update of one cell (src/discrete_operators.py:49-73): q + half a flux difference
template <int SPLIT>
PYCS_HD double inner_update(double q, double d, double rg, double cdv) {
  if (SPLIT == 1) return fma(0.5 * d, rg, q);
  if (SPLIT == 2) return fma(0.5 * fma(cdv, q, d), rg, q);
  return 0.5 * (q + (q + d) / (1.0 - cdv));
}

// Rolling state of one lane: everything is per column c of the lane's NC columns.
struct Lane {
  double qw[NC][5];       // Q   rows r-4 .. r
  double yw[NC][5];       // Qy  rows r-4 .. r
  double pl[NC], pr[NC];  // edge values of Q,  cell r-3
  double yl[NC], yr[NC];  // edge values of Qy, cell r-3
  double fin_prev[NC], fout_prev[NC];   // x-fluxes (inner on Q, outer on Qy) at edge r-3
  double su2[NC], su3[NC];              // sqrtg_pu rows r-2, r-3
  double cmx_prev[NC];                  // sqrtg_pu*cx at edge r-3 (SPLIT != 1)
  double psum;                          // sum of pxdF + pydF over this lane's outputs (MF-PR)
};

struct XEdge {            // x-edge r-2 quantities computed in phase 1, reused in phase 3
  double WE[NC], WO[NC], WG[NC], rg3[NC];
  bool up[NC];
};

PYCS_HD void lane_init(Lane& L) {
  for (int c = 0; c < NC; ++c) {
    for (int k = 0; k < 5; ++k) { L.qw[c][k] = 0.0; L.yw[c][k] = 0.0; }
    L.pl[c] = L.pr[c] = L.yl[c] = L.yr[c] = 0.0;
    L.fin_prev[c] = L.fout_prev[c] = 0.0;
    L.su2[c] = L.su3[c] = 0.0;
    L.cmx_prev[c] = 0.0;
  }
  L.psum = 0.0;
}

// ---- phase 1: row r enters; inner x-flux at edge r-2; Qx row r-3 -> qx[NC] -------------
// R0..R3: ring slots of rows r, r-1, r-2, r-3; ca: index of the lane's first column in a
// staged row (even).
template <int RECON, int SPLIT, int MASK, int RW>
PYCS_HD void phase_x_inner(Lane& L, XEdge& X, const double* R0, const double* R1, const double* R2,
                           const double* R3, int ca, double cdx, double ws, double qx[NC]) {
  constexpr int MT = (SPLIT == 3) ? 2 : 1;
  const dbl2 qn = ld2(R0 + A_Q * RW + ca);
  const dbl2 u2 = ld2(R2 + A_U * RW + ca);
  const dbl2 su1 = ld2(R1 + A_SGU * RW + ca);
  const dbl2 sc3 = ld2(R3 + A_SGC * RW + ca);
  const dbl2 sc2 = ld2(R2 + A_SGC * RW + ca);
  const dbl2 rg = ld2(R3 + A_RGC * RW + ca);
  dbl2 um2 = u2;
  if (MASK & 1) um2 = ld2(R2 + A_UM * RW + ca);
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    double* q = L.qw[c];
    q[0] = q[1]; q[1] = q[2]; q[2] = q[3]; q[3] = q[4];
    q[4] = pick(qn, c);
    double l2, r2;                                   // cell r-2
    edge_values<RECON>(q[0], q[1], q[2], q[3], q[4], l2, r2);
    double ub = pick(u2, c);
    if (MASK & 2) ub *= ws;
    const bool up = ((MASK & 1) ? pick(um2, c) : ub) >= 0.0;
    const double su1c = pick(su1, c);
    const double gE = L.su2[c];
    const double gO = up ? L.su3[c] : su1c;
    const double gC = up ? pick(sc3, c) : pick(sc2, c);
    double WE, WO, WG, cc;
    edge_weights<MT, !(MASK & 1)>(ub, up, cdx, gE, gO, gC, WE, WO, WG, cc);
    const double E = up ? L.pr[c] : l2, O = up ? L.pl[c] : r2, qc = up ? q[1] : q[2];
    const double fin = fma(WE, E, fma(WO, O, WG * qc));
    double cdv = 0.0;
    if (SPLIT != 1) {
      const double cmx = gE * cc;
      cdv = cmx - L.cmx_prev[c];
      L.cmx_prev[c] = cmx;
    }
    qx[c] = inner_update<SPLIT>(q[1], L.fin_prev[c] - fin, pick(rg, c), cdv);
    L.pl[c] = l2; L.pr[c] = r2;
    L.fin_prev[c] = fin;
    L.su3[c] = gE; L.su2[c] = su1c;
    X.WE[c] = WE; X.WO[c] = WO; X.WG[c] = WG; X.rg3[c] = pick(rg, c);
    X.up[c] = up;
  }
}

// ---- phase 2: y-fluxes at the lane's NC edges of one row ------------------------------------
// Rk: ring slot holding V / sqrtg of that row; src: the advected row (Q or Qx) positioned
// so that src[k] is the value k columns right of the lane's first column.
template <int RECON, int SPLIT, int MASK, int RW>
PYCS_HD void yflux_pair(const double* Rk, int ca, const double* src, double cdy, double ws, double f[NC],
                        double cmy[NC]) {
  constexpr int MT = (SPLIT == 3) ? 2 : 1;
  const dbl2 v2 = ld2(Rk + A_V * RW + ca);
  const dbl2 sgv = ld2(Rk + A_SGV * RW + ca);
  dbl2 vm2 = v2;
  if (MASK & 1) vm2 = ld2(Rk + A_VM * RW + ca);
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    double vb = pick(v2, c);
    if (MASK & 2) vb *= ws;
    const bool vp = ((MASK & 1) ? pick(vm2, c) : vb) >= 0.0;
    const int e = ca + c;
    const double gE = pick(sgv, c);
    const double gO = Rk[A_SGV * RW + (vp ? e - 1 : e + 1)];
    const double gC = Rk[A_SGC * RW + (vp ? e - 1 : e)];
    double WE, WO, WG, cc;
    edge_weights<MT, !(MASK & 1)>(vb, vp, cdy, gE, gO, gC, WE, WO, WG, cc);
    const double* s = src + c + (vp ? -1 : 0);      // upwind cell
    double l, r;
    edge_values<RECON>(s[-2], s[-1], s[0], s[1], s[2], l, r);
    const double E = vp ? r : l, O = vp ? l : r;
    f[c] = fma(WE, E, fma(WO, O, WG * s[0]));
    if (SPLIT != 1) cmy[c] = gE * cc;
  }
}

// ---- phase 3: Qy row r; outer x-flux at edge r-2 on Qy; output row r-3 ----------------------
// F, G: inner (row r) and outer (row r-3) y-fluxes at the lane's columns and at the next
// column (index NC); CM likewise sqrtg_pv*cy of row r.  out[c] = new Q of row r-3,
// sdiv[c] = pxdF + pydF of that cell.
template <int RECON, int SPLIT, int RW>
PYCS_HD void phase_x_outer(Lane& L, const XEdge& X, const double* R0, int ca, const double F[NC + 1],
                           const double G[NC + 1], const double CM[NC + 1], double out[NC], double sdiv[NC]) {
  const dbl2 rg0 = ld2(R0 + A_RGC * RW + ca);
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    double* y = L.yw[c];
    const double cdv = (SPLIT != 1) ? CM[c + 1] - CM[c] : 0.0;
    const double qy = inner_update<SPLIT>(L.qw[c][4], F[c] - F[c + 1], pick(rg0, c), cdv);
    y[0] = y[1]; y[1] = y[2]; y[2] = y[3]; y[3] = y[4];
    y[4] = qy;
    double l2, r2;                                   // Qy cell r-2
    edge_values<RECON>(y[0], y[1], y[2], y[3], y[4], l2, r2);
    const bool up = X.up[c];
    const double E = up ? L.yr[c] : l2, O = up ? L.yl[c] : r2, qc = up ? y[1] : y[2];
    const double fo = fma(X.WE[c], E, fma(X.WO[c], O, X.WG[c] * qc));
    const double s = (L.fout_prev[c] - fo) + (G[c] - G[c + 1]);
    out[c] = fma(s, X.rg3[c], L.qw[c][1]);
    sdiv[c] = s;
    L.yl[c] = l2; L.yr[c] = r2;
    L.fout_prev[c] = fo;
  }
}

// Column range of consumer warp w of a strip whose useful columns are [js0, js1): the warp
// marches the 64 columns starting at cw0 (even, so that column pairs are 16-byte aligned)
// and owns the outputs [us, ue).
PYCS_HD void warp_columns(int js0, int js1, int w, int& cw0, int& us, int& ue) {
  us = js0;
  for (int k = 0;; ++k) {
    cw0 = (us - 3) & ~1;
    ue = cw0 + WARP_COLS - 3;
    if (ue > js1) ue = js1;
    if (ue < us) ue = us;
    if (k == w) return;
    us = ue;
  }
}
// useful columns NW consumer warps can cover (first warp 57, the others 58)
PYCS_HD int strip_capacity(int nw) { return 57 + 58 * (nw - 1); }

}  // namespace f3
