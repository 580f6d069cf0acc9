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
#ifndef PYCS_FUSED3_CORE_PRELUDE
#define PYCS_FUSED3_CORE_PRELUDE
#if defined(__CUDACC__)
#define PYCS_HD __host__ __device__ __forceinline__
typedef double2 dbl2;
#else
#include <cmath>
#define PYCS_HD inline
struct alignas(16) dbl2 { double x, y; };
#endif
#endif

// The body below can be included more than once with different
//   F3_NAMESPACE (default f3), F3_NC (columns per lane, default 2), F3_CSTEP (distance between
//   a lane's columns, default 32)
// fused3.cu uses f3 / 2 / 32 (warp-autonomous kernel), fused2b.cu uses f1 / 1 / 0 (one column
// per thread, block-synchronous kernel).
#ifndef F3_NAMESPACE
#define F3_NAMESPACE f3
#define F3_NC 2
#define F3_CSTEP 32
#endif

namespace F3_NAMESPACE {

constexpr int NC = F3_NC;           // columns per lane
constexpr int CSTEP = F3_CSTEP;     // lane l owns columns cw0 + l and cw0 + CSTEP + l (conflict-free LDS.64)
constexpr int WARP_COLS = 32 * NC;  // columns marched by one consumer warp
constexpr int WARP_USE = WARP_COLS - 6;   // of which outputs (3 halo columns on each side)
constexpr int SXW = WARP_COLS + 8;  // private Qx row: 4 pad columns on each side
// Arrays staged per row.  Short ring (each row is needed by one march step only):
//   Q[r], U[r-2], sqrtg_pu[r-1] (+ mask source UM[r-2]);
// long ring (row r is needed by march steps r .. r+3):
//   V[r], sqrtg_pv[r], sqrtg_pc[r], 1/sqrtg_pc[r] (+ mask source VM[r]).
enum { S_Q = 0, S_U = 1, S_SGU = 2, S_UM = 3 };
enum { L_V = 0, L_SGV = 1, L_SGC = 2, L_RGC = 3, L_VM = 4 };

template <int NW> struct RowWidth { static constexpr int value = WARP_USE * NW + 14; };   // doubles per staged row

// Pointers to the staged rows one march step reads, already offset to the lane's first column.
struct RowPtrs {
  const double *q, *u, *um, *su1;        // short slot of row r
  const double *v0, *vm0, *sgv0, *sgc0, *rg0;   // long slot of row r
  const double *sgc2;                    // long slot of row r-2
  const double *v3, *vm3, *sgv3, *sgc3, *rg3;   // long slot of row r-3
};

PYCS_HD void st2(double* p, double a, double b) {
  dbl2 v; v.x = a; v.y = b;
  *reinterpret_cast<dbl2*>(p) = v;
}

// src/reconstruction_1d.py:36-62 (q3 is the cell itself)
// CBANK (device only, needs F3_COEF_BANK = a __constant__ double[5] holding the same five numbers):
// the coefficients are read as constant-bank operands of the DFMAs instead of being rebuilt with
// two UMOVs each wherever the uniform registers run out (9 UMOVs per row in the const-slot march).
template <int RECON, bool CBANK = false>
PYCS_HD void edge_values(double q1, double q2, double q3, double q4, double q5, double& l, double& r) {
  if (RECON == 3) {
#if defined(__CUDA_ARCH__) && defined(F3_COEF_BANK)
    const double a1 = CBANK ? F3_COEF_BANK[0] : 2.0 / 60.0, a2 = CBANK ? F3_COEF_BANK[1] : -13.0 / 60.0,
                 a3 = CBANK ? F3_COEF_BANK[2] : 47.0 / 60.0, a4 = CBANK ? F3_COEF_BANK[3] : 27.0 / 60.0,
                 a5 = CBANK ? F3_COEF_BANK[4] : -3.0 / 60.0;
#else
    const double a1 = 2.0 / 60.0, a2 = -13.0 / 60.0, a3 = 47.0 / 60.0, a4 = 27.0 / 60.0, a5 = -3.0 / 60.0;
#endif
    r = fma(a5, q5, fma(a4, q4, fma(a3, q3, fma(a2, q2, a1 * q1))));
    l = fma(a1, q5, fma(a2, q4, fma(a3, q3, fma(a4, q2, a5 * q1))));
  } else {
    const double c7 = 7.0 / 12.0, c1 = 1.0 / 12.0;
    l = fma(c7, q3 + q2, -(c1 * (q4 + q1)));
    r = fma(c7, q4 + q3, -(c1 * (q5 + q2)));
  }
}

// Weights of one edge.  c = ubar*dt/dx: the CFL number of the wind used in the flux (time
// averaged; src/cfl.py:10) -- the callers fold a separable wind's time factor into dt/dx, so the
// scaled wind itself is never formed; up: upwind mask; gE, gO, gC: sqrtg at the edge, at the
// other edge of the upwind cell and at its centre.
// SAMEMASK: the mask is the sign of the wind itself (RK1), so b = |c| (for u = -0.0 every
// weight is zero and the side does not matter).
template <int MT, bool SAMEMASK>
PYCS_HD void edge_weights(double c, bool up, double gE, double gO, double gC, double& WE, double& WO,
                          double& WG) {
  const double b = SAMEMASK ? fabs(c) : (up ? c : -c);
  const double m = 1.0 - b;
  const double t = fma(2.0, m, 1.0);          // 3 - 2b
  const double cs = (MT == 2) ? c * gE : c;   // MT-PL07: the whole flux times sqrtg at the edge
  const double cm = cs * m;
  const double cb = cs * b;
  const double w1 = cm * m, w2 = -(cb * m), w3 = cb * t;
  if (MT == 1) { WE = w1 * gE; WO = w2 * gO; WG = w3 * gC; }
  else { WE = w1; WO = w2; WG = w3; }
}

// inner update of one cell (src/discrete_operators.py:49-73): q + half a flux difference
template <int SPLIT>
PYCS_HD double inner_update(double q, double d, double rg, double cdv) {
  if (SPLIT == 1) return fma(0.5 * d, rg, q);
  if (SPLIT == 2) return fma(0.5 * fma(cdv, q, d), rg, q);
  return 0.5 * (q + (q + d) / (1.0 - cdv));
}

// The five-row windows come in two forms.  K < 0 (default): elements 0..4 hold rows r-4..r and
// shift every row (free when the march is unrolled by 5).  K >= 0: circular buffer of W
// registers, row r lives in element K = (r - first row) % W, so a march unrolled by W -- the
// period of the staged-row rings -- touches statically named registers and never moves one.
// W = 6 for the one-row march (rings of 3 and 6 slots), W = 8 for the two-row march (4 and 8).
constexpr int WLEN = 6;
constexpr int WMAX = 8;
template <int K, int I, int W = WLEN> PYCS_HD constexpr int widx() { return K < 0 ? I : (K + W - 4 + I) % W; }

// Rolling state of one lane: everything is per column c of the lane's NC columns.
struct Lane {
  double qw[NC][WMAX];    // Q   rows r-4 .. r (see widx)
  double yw[NC][WMAX];    // Qy  rows r-4 .. r
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
    for (int k = 0; k < WMAX; ++k) { L.qw[c][k] = 0.0; L.yw[c][k] = 0.0; }
    L.pl[c] = L.pr[c] = L.yl[c] = L.yr[c] = 0.0;
    L.fin_prev[c] = L.fout_prev[c] = 0.0;
    L.su2[c] = L.su3[c] = 0.0;
    L.cmx_prev[c] = 0.0;
  }
  L.psum = 0.0;
}

// ---- phase 1: row r enters; inner x-flux at edge r-2; Qx row r-3 -> qx[NC] -------------
// qnew[c]: Q of row r in the lane's columns (the caller loads it, and patches it when a
// projection term is pending).
// cdxw = dt/dx times the time factor of a separable wind (1 otherwise).
// DYNGC: sqrtg at the upwind centre through one dynamically addressed load (behind the wind-sign compare) instead
// of loading both candidates
template <int RECON, int SPLIT, int MASK, int K = -1, int W = WLEN, bool DYNGC = false>
PYCS_HD void phase_x_inner(Lane& L, XEdge& X, const RowPtrs& R, const double qnew[NC], double cdxw, double qx[NC]) {
  constexpr int MT = (SPLIT == 3) ? 2 : 1;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int o = c * CSTEP;
    double* q = L.qw[c];
    if (K < 0) { q[0] = q[1]; q[1] = q[2]; q[2] = q[3]; q[3] = q[4]; }
    q[widx<K, 4, W>()] = qnew[c];
    double l2, r2;                                   // cell r-2
    edge_values<RECON, (K >= 0)>(q[widx<K, 0, W>()], q[widx<K, 1, W>()], q[widx<K, 2, W>()], q[widx<K, 3, W>()], q[widx<K, 4, W>()], l2, r2);
    const double cc = R.u[o] * cdxw;                // CFL number at edge r-2
    const bool up = ((MASK & 1) ? R.um[o] : cc) >= 0.0;
    const double su1c = R.su1[o];
    const double gE = L.su2[c];
    const double gO = up ? L.su3[c] : su1c;
    const double gC = DYNGC ? (up ? R.sgc3 : R.sgc2)[o]
                            : (up ? R.sgc3[o] : R.sgc2[o]);   // both loads are issued early; a pointer select
                                                              // puts the LDS behind the wind-sign compare
    const double rg = R.rg3[o];
    double WE, WO, WG;
    edge_weights<MT, !(MASK & 1)>(cc, up, gE, gO, gC, WE, WO, WG);
    const double E = up ? L.pr[c] : l2, O = up ? L.pl[c] : r2, qc = up ? q[widx<K, 1, W>()] : q[widx<K, 2, W>()];
    const double fin = fma(WE, E, fma(WO, O, WG * qc));
    double cdv = 0.0;
    if (SPLIT != 1) {
      const double cmx = gE * cc;
      cdv = cmx - L.cmx_prev[c];
      L.cmx_prev[c] = cmx;
    }
    qx[c] = inner_update<SPLIT>(q[widx<K, 1, W>()], L.fin_prev[c] - fin, rg, cdv);
    L.pl[c] = l2; L.pr[c] = r2;
    L.fin_prev[c] = fin;
    L.su3[c] = gE; L.su2[c] = su1c;
    X.WE[c] = WE; X.WO[c] = WO; X.WG[c] = WG; X.rg3[c] = rg;
    X.up[c] = up;
  }
}

// ---- phase 2: y-fluxes at the lane's NC edges of one row ------------------------------------
// v, vm, sgv, sgc: staged rows of that row (lane-offset); src: the advected row (Q or Qx),
// lane-offset as well, so that src[k] is the value k columns right of the lane's first column.
// OWN (const-slot march only): own[c] = src[c * CSTEP], the lane's own cell, which it holds in a register.  The
// stencil is then read mirrored about the edge -- x1..x5 run from the far upwind cell towards the downwind one,
// so that E = r(x) and O = l(x) whichever way the wind blows -- and costs four loads instead of five.
// HCC: cch[c] = v[c * CSTEP] * cdyw was formed by the caller already (before the barrier in front of this phase: the
// wind row is TMA-staged, not thread-written, so the head of the dependency chain need not wait for the barrier).
template <int RECON, int SPLIT, int MASK, int K = -1, bool OWN = false, bool HCC = false>
PYCS_HD void yflux_pair(const double* v, const double* vm, const double* sgv, const double* sgc,
                        const double* src, double cdyw, double f[NC], double cmy[NC], const double* own = nullptr,
                        const double* cch = nullptr) {
  constexpr int MT = (SPLIT == 3) ? 2 : 1;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int o = c * CSTEP;
    const double cc = HCC ? cch[c] : v[o] * cdyw;
    const bool vp = ((MASK & 1) ? vm[o] : cc) >= 0.0;
    const int up1 = vp ? -1 : 0;                     // upwind cell relative to the edge
    const double gE = sgv[o];
    const double gO = sgv[o + 1 + 2 * up1];          // other edge of the upwind cell
    const double gC = sgc[o + up1];
    if (K >= 0) {
      // a y-edge's weights serve one flux only, so the factored form
      //   f = c [ m (m E' - b O') + b (3 - 2b) G' ],  E' = E gE, O' = O gO, G' = q gC  (MT-0)
      // is two FP64 operations shorter than forming the three weights (const-slot march only:
      // the shifting-window kernels keep the arithmetic they were measured with)
      double E, O, qc;
      if (OWN) {
        const double* s0 = src + o;
        const double xa = s0[-2], xb = s0[-1], xc = s0[1], x1 = s0[vp ? -3 : 2], xo = own[c];
        const double x2 = vp ? xa : xc, x5 = vp ? xc : xa, x3 = vp ? xb : xo, x4 = vp ? xo : xb;
        edge_values<RECON, true>(x1, x2, x3, x4, x5, O, E);
        qc = x3;
      } else {
        const double* s = src + o + up1;
        double l, r;
        edge_values<RECON, true>(s[-2], s[-1], s[0], s[1], s[2], l, r);
        E = vp ? r : l;
        O = vp ? l : r;
        qc = s[0];
      }
      const double b = !(MASK & 1) ? fabs(cc) : (vp ? cc : -cc);
      const double m = 1.0 - b, bt = b * fma(2.0, m, 1.0);
      const double Eg = (MT == 1) ? E * gE : E, Og = (MT == 1) ? O * gO : O, Gg = (MT == 1) ? qc * gC : qc;
      const double cs = (MT == 2) ? cc * gE : cc;
      f[c] = cs * fma(m, fma(m, Eg, -(b * Og)), bt * Gg);
    } else {
      double WE, WO, WG;
      edge_weights<MT, !(MASK & 1)>(cc, vp, gE, gO, gC, WE, WO, WG);
      const double* s = src + o + up1;
      double l, r;
      edge_values<RECON>(s[-2], s[-1], s[0], s[1], s[2], l, r);
      const double E = vp ? r : l, O = vp ? l : r;
      f[c] = fma(WE, E, fma(WO, O, WG * s[0]));
    }
    if (SPLIT != 1) cmy[c] = gE * cc;
  }
}

// ---- phase 3: Qy row r; outer x-flux at edge r-2 on Qy; output row r-3 ----------------------
// F, Fn: inner y-fluxes (row r) at the lane's columns and at the columns right of them;
// G, Gn likewise the outer y-fluxes of row r-3; CM, CMn sqrtg_pv*cy of row r (SPLIT != 1).
// out[c] = new Q of row r-3, sdiv[c] = pxdF + pydF of that cell.
template <int RECON, int SPLIT, int K = -1, int W = WLEN>
PYCS_HD void phase_x_outer(Lane& L, const XEdge& X, const RowPtrs& R, const double F[NC], const double Fn[NC],
                           const double G[NC], const double Gn[NC], const double CM[NC], const double CMn[NC],
                           double out[NC], double sdiv[NC]) {
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    double* y = L.yw[c];
    const double cdv = (SPLIT != 1) ? CMn[c] - CM[c] : 0.0;
    const double qy = inner_update<SPLIT>(L.qw[c][widx<K, 4, W>()], F[c] - Fn[c], R.rg0[c * CSTEP], cdv);
    if (K < 0) { y[0] = y[1]; y[1] = y[2]; y[2] = y[3]; y[3] = y[4]; }
    y[widx<K, 4, W>()] = qy;
    double l2, r2;                                   // Qy cell r-2
    edge_values<RECON, (K >= 0)>(y[widx<K, 0, W>()], y[widx<K, 1, W>()], y[widx<K, 2, W>()], y[widx<K, 3, W>()], y[widx<K, 4, W>()], l2, r2);
    const bool up = X.up[c];
    const double E = up ? L.yr[c] : l2, O = up ? L.yl[c] : r2, qc = up ? y[widx<K, 1, W>()] : y[widx<K, 2, W>()];
    const double fo = fma(X.WE[c], E, fma(X.WO[c], O, X.WG[c] * qc));
    const double s = (L.fout_prev[c] - fo) + (G[c] - Gn[c]);
    out[c] = fma(s, X.rg3[c], L.qw[c][widx<K, 1, W>()]);
    sdiv[c] = s;
    L.yl[c] = l2; L.yr[c] = r2;
    L.fout_prev[c] = fo;
  }
}

// Consumer warp w of a strip whose useful columns start at js0 marches the 64 columns from
// cw0 = js0 - 3 + 58 w and owns the outputs [us, ue), clipped to the strip end js1.
PYCS_HD void warp_columns(int js0, int js1, int w, int& cw0, int& us, int& ue) {
  us = js0 + WARP_USE * w;
  cw0 = us - 3;
  ue = us + WARP_USE;
  if (ue > js1) ue = js1;
  if (us > js1) us = js1;
}
// useful columns NW consumer warps can cover
PYCS_HD int strip_capacity(int nw) { return WARP_USE * nw; }
// first staged column of a strip (even, so that the TMA source is 16-byte aligned) and the
// index of column cw0 of warp w inside a staged row
PYCS_HD int strip_c0(int js0) { return ((js0 - 3) & ~1) - 4; }

}  // namespace F3_NAMESPACE
#undef F3_NAMESPACE
#undef F3_NC
#undef F3_CSTEP
