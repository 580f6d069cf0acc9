// Host emulator of the v3 fused step kernel -- TEST INFRASTRUCTURE.
//
// Compiles py-cubed-sphere_b200/csrc/fused3_core.cuh (the per-lane arithmetic the CUDA
// kernel runs, shared verbatim) with g++ and replays the kernel's decomposition on the
// CPU: CTAs = (panel, strip, chunk), consumer warps, 32 lanes x 2 columns, the two rings
// of staged rows filled the way the TMA producer fills it (row segment copy + MF-PR
// patch), the warp-private Qx row and the lane-to-lane flux exchange.  tests/ compares its
// output with the numpy oracle, which validates the numerics and every index of the kernel
// without a GPU; the mbarrier protocol itself is only exercised on the device.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../py-cubed-sphere_b200/csrc/fused3_core.cuh"
#define F3_NAMESPACE f1
#define F3_NC 1
#define F3_CSTEP 0
#include "../../py-cubed-sphere_b200/csrc/fused3_core.cuh"

namespace {
constexpr int JOFF = 12;   // PYCS_JOFF

struct Args {
  int N, P, ld, lo, hi;
  long long ps;
  const double *q, *ua, *va, *um, *vm, *sgc, *rgc, *sgu, *sgv;
  double *qn, *part;
  double corr, cdx, cdy, ws;
  int apply_corr, rows_per_chunk, nstrips, wcols, depth;
  int circ = 0;      // v2b: circular six-register windows (the const-slot march of fused2b.cu, MINB >= 30)
};

template <int RECON, int SPLIT, int MASK, int NW>
void run(const Args& a) {
  using namespace f3;
  const double cdxw = a.cdx * ((MASK & 2) ? a.ws : 1.0), cdyw = a.cdy * ((MASK & 2) ? a.ws : 1.0);
  constexpr int RW = RowWidth<NW>::value;
  constexpr int NS = (MASK & 1) ? 4 : 3, NL = (MASK & 1) ? 5 : 4;
  constexpr int SSLOT = NS * RW, LSLOT = NL * RW;
  const int PF = a.depth, DS = PF + 1, DL = PF + 4;        // "depth" = rows in flight
  const int nchunks = (a.N + a.rows_per_chunk - 1) / a.rows_per_chunk;
  std::vector<double> ringbuf((size_t)DS * SSLOT + (size_t)DL * LSLOT + NW * SXW);
  double* ringS = ringbuf.data();
  double* ringL = ringS + DS * SSLOT;
  double* sxall = ringL + DL * LSLOT;
  for (int blk = 0; blk < 6 * a.nstrips * nchunks; ++blk) {
    int b = blk;
    const int p = b % 6;
    b /= 6;
    const int strip = b % a.nstrips, chunk = b / a.nstrips;
    const int js0 = a.lo + strip * a.wcols;
    const int js1 = std::min(js0 + a.wcols, a.hi);
    const int r0 = a.lo + chunk * a.rows_per_chunk;   // single GPU: row_lo = lo, row_hi = hi
    const int r1 = std::min(r0 + a.rows_per_chunk, a.hi);
    const int rfirst = r0 - 3, rlast = r1 + 2;
    const int c0 = strip_c0(js0);
    const int len = std::min(RW, a.ld - JOFF - c0) & ~1;
    const long long colb = (long long)p * a.ps + JOFF + c0, colm = JOFF + c0;
    for (int warp = 0; warp < NW; ++warp) {
      std::fill(ringbuf.begin(), ringbuf.end(), 0.0);
      double* sxrow = sxall + warp * SXW;
      int cw0, us, ue;
      warp_columns(js0, js1, warp, cw0, us, ue);
      Lane L[32];
      for (int l = 0; l < 32; ++l) lane_init(L[l]);
      int oS = 0, oL0 = 0, oL1 = (DL - 1) * LSLOT, oL2 = (DL - 2) * LSLOT, oL3 = (DL - 3) * LSLOT;
      for (int r = rfirst; r <= rlast; ++r) {
        // producer: stage row r into its slots, patch the pending MF-PR term
        {
          const int k = r - rfirst;
          double* dS = ringS + (k % DS) * SSLOT;
          double* dL = ringL + (k % DL) * LSLOT;
          const long long rr = (long long)r * a.ld, rm1 = (long long)std::max(r - 1, 0) * a.ld,
                          rm2 = (long long)std::max(r - 2, 0) * a.ld;
          auto cp = [&](double* dst, const double* src) { std::memcpy(dst, src, sizeof(double) * len); };
          cp(dS + S_Q * RW, a.q + colb + rr); cp(dL + L_V * RW, a.va + colb + rr);
          cp(dL + L_SGC * RW, a.sgc + colm + rr); cp(dL + L_SGV * RW, a.sgv + colm + rr);
          cp(dL + L_RGC * RW, a.rgc + colm + rr); cp(dS + S_SGU * RW, a.sgu + colm + rm1);
          cp(dS + S_U * RW, a.ua + colb + rm2);
          if (MASK & 1) { cp(dL + L_VM * RW, a.vm + colb + rr); cp(dS + S_UM * RW, a.um + colb + rm2); }
          if (a.apply_corr && r >= a.lo && r < a.hi)
            for (int i = 0; i < len; ++i) {
              const int j = c0 + i;
              if (j >= a.lo && j < a.hi) dS[S_Q * RW + i] = fma(dL[L_SGC * RW + i], a.corr, dS[S_Q * RW + i]);
            }
          if ((k % DS) * SSLOT != oS || (k % DL) * LSLOT != oL0) std::abort();
        }
        RowPtrs R[32];
        XEdge X[32];
        double F[32][NC], G[32][NC], CF[32][NC], CG[NC], Fn[32][NC], Gn[32][NC], CFn[32][NC];
        for (int l = 0; l < 32; ++l) {                 // phase 1
          const int ca = cw0 - c0 + l;
          RowPtrs& P = R[l];
          P.q = ringS + oS + S_Q * RW + ca; P.u = ringS + oS + S_U * RW + ca;
          P.um = ringS + oS + S_UM * RW + ca; P.su1 = ringS + oS + S_SGU * RW + ca;
          P.v0 = ringL + oL0 + L_V * RW + ca; P.vm0 = ringL + oL0 + L_VM * RW + ca;
          P.sgv0 = ringL + oL0 + L_SGV * RW + ca; P.sgc0 = ringL + oL0 + L_SGC * RW + ca;
          P.rg0 = ringL + oL0 + L_RGC * RW + ca; P.sgc2 = ringL + oL2 + L_SGC * RW + ca;
          P.v3 = ringL + oL3 + L_V * RW + ca; P.vm3 = ringL + oL3 + L_VM * RW + ca;
          P.sgv3 = ringL + oL3 + L_SGV * RW + ca; P.sgc3 = ringL + oL3 + L_SGC * RW + ca;
          P.rg3 = ringL + oL3 + L_RGC * RW + ca;
          double qx[NC];
          const double qnew[NC] = {P.q[0], P.q[CSTEP]};
          phase_x_inner<RECON, SPLIT, MASK>(L[l], X[l], P, qnew, cdxw, qx);
          sxrow[4 + l] = qx[0];
          sxrow[4 + l + CSTEP] = qx[1];
        }
        for (int l = 0; l < 32; ++l) {                 // phase 2 (after __syncwarp)
          CF[l][0] = CF[l][1] = 0.0;
          yflux_pair<RECON, SPLIT, MASK>(R[l].v0, R[l].vm0, R[l].sgv0, R[l].sgc0, R[l].q, cdyw, F[l], CF[l]);
          yflux_pair<RECON, SPLIT, MASK>(R[l].v3, R[l].vm3, R[l].sgv3, R[l].sgc3, sxrow + 4 + l, cdyw, G[l], CG);
        }
        for (int l = 0; l < 32; ++l) {                 // shuffles: right neighbour of each own column
          const int n = l < 31 ? l + 1 : l;            // __shfl_down keeps the own value in lane 31
          Fn[l][0] = l == 31 ? F[0][1] : F[n][0]; Fn[l][1] = F[n][1];
          Gn[l][0] = l == 31 ? G[0][1] : G[n][0]; Gn[l][1] = G[n][1];
          CFn[l][0] = l == 31 ? CF[0][1] : CF[n][0]; CFn[l][1] = CF[n][1];
        }
        for (int l = 0; l < 32; ++l) {                 // phase 3
          const int col = cw0 + l;
          double out[NC], sdiv[NC];
          phase_x_outer<RECON, SPLIT>(L[l], X[l], R[l], F[l], Fn[l], G[l], Gn[l], CF[l], CFn[l], out, sdiv);
          if (r >= r0 + 3) {
            double* QN = a.qn + (long long)p * a.ps + JOFF + (long long)(r - 3) * a.ld;
            for (int c = 0; c < NC; ++c) {
              const int cc = col + c * CSTEP;
              if (cc >= us && cc < ue) { QN[cc] = out[c]; L[l].psum += sdiv[c]; }
            }
          }
        }
        oS = (oS + SSLOT == DS * SSLOT) ? 0 : oS + SSLOT;
        oL3 = oL2; oL2 = oL1; oL1 = oL0;
        oL0 = (oL0 + LSLOT == DL * LSLOT) ? 0 : oL0 + LSLOT;
      }
      // warp reduction in the kernel's order (shfl_down tree)
      double v[32];
      for (int l = 0; l < 32; ++l) v[l] = L[l].psum;
      for (int o = 16; o > 0; o >>= 1)
        for (int l = 0; l + o < 32; ++l) v[l] += v[l + o];
      a.part[(long long)blk * NW + warp] = v[0];
    }
  }
}

// ---- v2b (csrc/fused2b.cu): one column per thread, CTA = strip of TB-6 columns -------------
// optional CTA subset of the next v2b runs (the split step of FusedArgs::blk_map): nullptr = whole grid
static const int* g_blk_list = nullptr;
static int g_blk_count = 0;
// compile-time row phase k = (r - first row) % 6 of the circular-window march
// march variant of the const-slot emulation (bits as F2B_VAR in csrc/fused2b.cu: 1 = copies of row r+3 issued after
// barrier B, 8 = own cell from the register in the y-stencils, 16 = single dynamically addressed sqrtg load)
static int g_var = 0;
extern "C" void f3_emul_set_variant(int v) { g_var = v; }
template <int RECON, int SPLIT, int MASK, int W = f1::WLEN>
void x_inner_k(int k, f1::Lane& L, f1::XEdge& X, const f1::RowPtrs& R, const double* qnew, double cdxw, double* qx) {
  using namespace f1;
  if (k >= 0 && k < 6 && (g_var & 16)) {
    switch (k) {
      case 0: phase_x_inner<RECON, SPLIT, MASK, 0, W, true>(L, X, R, qnew, cdxw, qx); break;
      case 1: phase_x_inner<RECON, SPLIT, MASK, 1, W, true>(L, X, R, qnew, cdxw, qx); break;
      case 2: phase_x_inner<RECON, SPLIT, MASK, 2, W, true>(L, X, R, qnew, cdxw, qx); break;
      case 3: phase_x_inner<RECON, SPLIT, MASK, 3, W, true>(L, X, R, qnew, cdxw, qx); break;
      case 4: phase_x_inner<RECON, SPLIT, MASK, 4, W, true>(L, X, R, qnew, cdxw, qx); break;
      default: phase_x_inner<RECON, SPLIT, MASK, 5, W, true>(L, X, R, qnew, cdxw, qx); break;
    }
    return;
  }
  switch (k) {
    case 0: phase_x_inner<RECON, SPLIT, MASK, 0, W>(L, X, R, qnew, cdxw, qx); break;
    case 1: phase_x_inner<RECON, SPLIT, MASK, 1, W>(L, X, R, qnew, cdxw, qx); break;
    case 2: phase_x_inner<RECON, SPLIT, MASK, 2, W>(L, X, R, qnew, cdxw, qx); break;
    case 3: phase_x_inner<RECON, SPLIT, MASK, 3, W>(L, X, R, qnew, cdxw, qx); break;
    case 4: phase_x_inner<RECON, SPLIT, MASK, 4, W>(L, X, R, qnew, cdxw, qx); break;
    case 5: phase_x_inner<RECON, SPLIT, MASK, 5, W>(L, X, R, qnew, cdxw, qx); break;
    case 6: phase_x_inner<RECON, SPLIT, MASK, 6, f1::WMAX>(L, X, R, qnew, cdxw, qx); break;
    case 7: phase_x_inner<RECON, SPLIT, MASK, 7, f1::WMAX>(L, X, R, qnew, cdxw, qx); break;
    default: phase_x_inner<RECON, SPLIT, MASK>(L, X, R, qnew, cdxw, qx);
  }
}
template <int RECON, int SPLIT, int W = f1::WLEN>
void x_outer_k(int k, f1::Lane& L, const f1::XEdge& X, const f1::RowPtrs& R, const double* f, const double* fn,
               const double* g, const double* gn, const double* cf, const double* cfn, double* out, double* sdiv) {
  using namespace f1;
  switch (k) {
    case 0: phase_x_outer<RECON, SPLIT, 0, W>(L, X, R, f, fn, g, gn, cf, cfn, out, sdiv); break;
    case 1: phase_x_outer<RECON, SPLIT, 1, W>(L, X, R, f, fn, g, gn, cf, cfn, out, sdiv); break;
    case 2: phase_x_outer<RECON, SPLIT, 2, W>(L, X, R, f, fn, g, gn, cf, cfn, out, sdiv); break;
    case 3: phase_x_outer<RECON, SPLIT, 3, W>(L, X, R, f, fn, g, gn, cf, cfn, out, sdiv); break;
    case 4: phase_x_outer<RECON, SPLIT, 4, W>(L, X, R, f, fn, g, gn, cf, cfn, out, sdiv); break;
    case 5: phase_x_outer<RECON, SPLIT, 5, W>(L, X, R, f, fn, g, gn, cf, cfn, out, sdiv); break;
    case 6: phase_x_outer<RECON, SPLIT, 6, f1::WMAX>(L, X, R, f, fn, g, gn, cf, cfn, out, sdiv); break;
    case 7: phase_x_outer<RECON, SPLIT, 7, f1::WMAX>(L, X, R, f, fn, g, gn, cf, cfn, out, sdiv); break;
    default: phase_x_outer<RECON, SPLIT>(L, X, R, f, fn, g, gn, cf, cfn, out, sdiv);
  }
}

template <int RECON, int SPLIT, int MASK>
void run_block(const Args& a, int TB) {
  using namespace f1;
  const double cdxw = a.cdx * ((MASK & 2) ? a.ws : 1.0), cdyw = a.cdy * ((MASK & 2) ? a.ws : 1.0);
  const int RW = TB + 6;
  const int NS = (MASK & 1) ? 4 : 3, NL = (MASK & 1) ? 5 : 4;
  const int SSLOT = NS * RW, LSLOT = NL * RW;
  const int PF = a.depth, DS = PF + 1, DL = PF + 4;
  const int wmax = TB - 6;
  const int nstrips = (a.N + wmax - 1) / wmax;
  int wcols = (a.N + nstrips - 1) / nstrips;
  wcols += wcols & 1;
  const int nchunks = (a.N + a.rows_per_chunk - 1) / a.rows_per_chunk;
  std::vector<double> buf((size_t)DS * SSLOT + (size_t)DL * LSLOT + 4 * RW);
  double* ringS = buf.data();
  double* ringL = ringS + DS * SSLOT;
  double* sX = ringL + DL * LSLOT;
  double *sF = sX + RW, *sG = sF + RW, *sC = sG + RW;
  std::vector<Lane> L(TB);
  std::vector<XEdge> X(TB);
  std::vector<RowPtrs> R(TB);
  std::vector<double> F(TB), G(TB), CF(TB);
  for (int idx = 0; idx < (g_blk_list ? g_blk_count : 6 * nstrips * nchunks); ++idx) {
    const int blk = g_blk_list ? g_blk_list[idx] : idx;
    int b = blk;
    const int p = b % 6;
    b /= 6;
    const int strip = b % nstrips, chunk = b / nstrips;
    const int jbase = a.lo + strip * wcols;
    const int jend = std::min(jbase + wcols, a.hi);
    const int r0 = a.lo + chunk * a.rows_per_chunk;
    const int r1 = std::min(r0 + a.rows_per_chunk, a.hi);
    const int rfirst = r0 - 3, rlast = r1 + 2;
    const int c0 = jbase - 6;
    const int len = std::min(RW, a.ld - JOFF - c0) & ~1;
    const long long colb = (long long)p * a.ps + JOFF + c0, colm = JOFF + c0;
    std::fill(buf.begin(), buf.end(), 0.0);
    for (int t = 0; t < TB; ++t) lane_init(L[t]);
    int oS = 0, oL0 = 0, oL1 = (DL - 1) * LSLOT, oL2 = (DL - 2) * LSLOT, oL3 = (DL - 3) * LSLOT;
    double psum = 0.0;
    // TMA row copies are made when the kernel issues them: rows rfirst .. rfirst+PF-1 before the march, row
    // r+PF right after barrier A of row r -- a slot that is overwritten too early shows up as a parity failure
    auto issue = [&](int r) {
      const int k = r - rfirst;
      double* dS = ringS + (k % DS) * SSLOT;
      double* dL = ringL + (k % DL) * LSLOT;
      const long long rr = (long long)r * a.ld, rm1 = (long long)std::max(r - 1, 0) * a.ld,
                      rm2 = (long long)std::max(r - 2, 0) * a.ld;
      auto cp = [&](double* dst, const double* src) { std::memcpy(dst, src, sizeof(double) * len); };
      cp(dS + S_Q * RW, a.q + colb + rr); cp(dL + L_V * RW, a.va + colb + rr);
      cp(dL + L_SGC * RW, a.sgc + colm + rr); cp(dL + L_SGV * RW, a.sgv + colm + rr);
      cp(dL + L_RGC * RW, a.rgc + colm + rr); cp(dS + S_SGU * RW, a.sgu + colm + rm1);
      cp(dS + S_U * RW, a.ua + colb + rm2);
      if (MASK & 1) { cp(dL + L_VM * RW, a.vm + colb + rr); cp(dS + S_UM * RW, a.um + colb + rm2); }
    };
    const bool lateB = a.circ && (g_var & 1);      // copies of row r+3 after barrier B instead of row r+PF after barrier A
    const int ahead = lateB ? 3 : PF;
    std::vector<double> QXo(TB), QNo(TB);          // the threads' registers qx, qnew (variant 8)
    for (int r = rfirst; r < rfirst + ahead && r <= rlast; ++r) issue(r);
    for (int r = rfirst; r <= rlast; ++r) {
      const int kc = a.circ ? (r - rfirst) % WLEN : -1;
      for (int t = 0; t < TB; ++t) {               // patch + phase 1
        const int e = t + 3, j = jbase - 3 + t;
        RowPtrs& P = R[t];
        P.q = ringS + oS + S_Q * RW + e; P.u = ringS + oS + S_U * RW + e;
        P.um = ringS + oS + S_UM * RW + e; P.su1 = ringS + oS + S_SGU * RW + e;
        P.v0 = ringL + oL0 + L_V * RW + e; P.vm0 = ringL + oL0 + L_VM * RW + e;
        P.sgv0 = ringL + oL0 + L_SGV * RW + e; P.sgc0 = ringL + oL0 + L_SGC * RW + e;
        P.rg0 = ringL + oL0 + L_RGC * RW + e; P.sgc2 = ringL + oL2 + L_SGC * RW + e;
        P.v3 = ringL + oL3 + L_V * RW + e; P.vm3 = ringL + oL3 + L_VM * RW + e;
        P.sgv3 = ringL + oL3 + L_SGV * RW + e; P.sgc3 = ringL + oL3 + L_SGC * RW + e;
        P.rg3 = ringL + oL3 + L_RGC * RW + e;
        double qnew[1] = {P.q[0]};
        if (a.apply_corr && j >= a.lo && j < a.hi && r >= a.lo && r < a.hi) {
          qnew[0] = fma(P.sgc0[0], a.corr, qnew[0]);
          ringS[oS + S_Q * RW + e] = qnew[0];
        }
        double qx[1];
        x_inner_k<RECON, SPLIT, MASK>(kc, L[t], X[t], P, qnew, cdxw, qx);
        sX[e] = qx[0];
        QXo[t] = qx[0]; QNo[t] = qnew[0];
      }
      if (!lateB && r + PF <= rlast) issue(r + PF);   // barrier A passed: warp 0 issues the copies of row r+PF
      for (int t = 0; t < TB; ++t) {               // barrier A; phase 2
        const int e = t + 3;
        double f[1], g[1], cf[1] = {0.0}, cg[1];
        if (a.circ && (g_var & 8)) {
          yflux_pair<RECON, SPLIT, MASK, 0, true>(R[t].v0, R[t].vm0, R[t].sgv0, R[t].sgc0, R[t].q, cdyw, f, cf, &QNo[t]);
          yflux_pair<RECON, SPLIT, MASK, 0, true>(R[t].v3, R[t].vm3, R[t].sgv3, R[t].sgc3, sX + e, cdyw, g, cg, &QXo[t]);
        } else if (a.circ) {   // K >= 0 selects the factored y-flux of the const-slot march (any phase: K is only a flag here)
          yflux_pair<RECON, SPLIT, MASK, 0>(R[t].v0, R[t].vm0, R[t].sgv0, R[t].sgc0, R[t].q, cdyw, f, cf);
          yflux_pair<RECON, SPLIT, MASK, 0>(R[t].v3, R[t].vm3, R[t].sgv3, R[t].sgc3, sX + e, cdyw, g, cg);
        } else {
          yflux_pair<RECON, SPLIT, MASK>(R[t].v0, R[t].vm0, R[t].sgv0, R[t].sgc0, R[t].q, cdyw, f, cf);
          yflux_pair<RECON, SPLIT, MASK>(R[t].v3, R[t].vm3, R[t].sgv3, R[t].sgc3, sX + e, cdyw, g, cg);
        }
        F[t] = f[0]; G[t] = g[0]; CF[t] = cf[0];
      }
      for (int t = 0; t < TB; ++t) { sF[t + 3] = F[t]; sG[t + 3] = G[t]; sC[t + 3] = CF[t]; }
      if (lateB && r + 3 <= rlast) issue(r + 3);   // barrier B passed: the copies of row r+3 land while phase 3 runs
      for (int t = 0; t < TB; ++t) {               // barrier B; phase 3
        const int e = t + 3, j = jbase - 3 + t;
        double f[1] = {F[t]}, g[1] = {G[t]}, cf[1] = {CF[t]};
        double fn[1] = {sF[e + 1]}, gn[1] = {sG[e + 1]}, cfn[1] = {sC[e + 1]}, out[1], sdiv[1];
        x_outer_k<RECON, SPLIT>(kc, L[t], X[t], R[t], f, fn, g, gn, cf, cfn, out, sdiv);
        if (r >= r0 + 3 && t >= 3 && j < jend) {
          a.qn[(long long)p * a.ps + JOFF + j + (long long)(r - 3) * a.ld] = out[0];
          L[t].psum += sdiv[0];
        }
      }
      oS = (oS + SSLOT == DS * SSLOT) ? 0 : oS + SSLOT;
      oL3 = oL2; oL2 = oL1; oL1 = oL0;
      oL0 = (oL0 + LSLOT == DL * LSLOT) ? 0 : oL0 + LSLOT;
    }
    for (int t = 0; t < TB; ++t) psum += L[t].psum;
    a.part[blk] = psum;
  }
}

// ---- v2b, two rows per pair of barriers (csrc/fused2b.cu, MINB >= 50): rings of 4 and 8 slots, the copies of
// step s+2 issued after barrier B of step s, circular windows of 8 registers, rows of a pair interleaved
// phase by phase exactly as in the kernel.
template <int RECON, int SPLIT, int MASK>
void run_block_pair(const Args& a, int TB) {
  using namespace f1;
  const double cdxw = a.cdx * ((MASK & 2) ? a.ws : 1.0), cdyw = a.cdy * ((MASK & 2) ? a.ws : 1.0);
  const int RW = TB + 6;
  const int NS = (MASK & 1) ? 4 : 3, NL = (MASK & 1) ? 5 : 4;
  const int SSLOT = NS * RW, LSLOT = NL * RW;
  const int DS = 4, DL = 8;
  const int wmax = TB - 6;
  const int nstrips = (a.N + wmax - 1) / wmax;
  int wcols = (a.N + nstrips - 1) / nstrips;
  wcols += wcols & 1;
  const int nchunks = (a.N + a.rows_per_chunk - 1) / a.rows_per_chunk;
  std::vector<double> buf((size_t)DS * SSLOT + (size_t)DL * LSLOT + 8 * RW);
  double* ringS = buf.data();
  double* ringL = ringS + DS * SSLOT;
  double* sX[2] = {ringL + DL * LSLOT, ringL + DL * LSLOT + 4 * RW};
  double* sF[2] = {sX[0] + RW, sX[1] + RW};
  double* sG[2] = {sF[0] + RW, sF[1] + RW};
  double* sC[2] = {sG[0] + RW, sG[1] + RW};
  std::vector<Lane> L(TB);
  std::vector<XEdge> X[2] = {std::vector<XEdge>(TB), std::vector<XEdge>(TB)};
  std::vector<RowPtrs> R[2] = {std::vector<RowPtrs>(TB), std::vector<RowPtrs>(TB)};
  std::vector<double> F[2] = {std::vector<double>(TB), std::vector<double>(TB)}, G[2] = {F[0], F[0]}, CF[2] = {F[0], F[0]};
  for (int idx = 0; idx < (g_blk_list ? g_blk_count : 6 * nstrips * nchunks); ++idx) {
    const int blk = g_blk_list ? g_blk_list[idx] : idx;
    int b = blk;
    const int p = b % 6;
    b /= 6;
    const int strip = b % nstrips, chunk = b / nstrips;
    const int jbase = a.lo + strip * wcols;
    const int jend = std::min(jbase + wcols, a.hi);
    const int r0 = a.lo + chunk * a.rows_per_chunk;
    const int r1 = std::min(r0 + a.rows_per_chunk, a.hi);
    const int rfirst = r0 - 3, rlast = r1 + 2;
    const int rlastp = rlast + ((rlast - rfirst + 1) & 1);
    const int c0 = jbase - 6;
    const int len = std::min(RW, a.ld - JOFF - c0) & ~1;
    const long long colb = (long long)p * a.ps + JOFF + c0, colm = JOFF + c0;
    std::fill(buf.begin(), buf.end(), 0.0);
    for (int t = 0; t < TB; ++t) lane_init(L[t]);
    auto issue = [&](int r) {     // TMA row copies of row r
      const int k = (r - rfirst) % DL;
      double* dS = ringS + (k % DS) * SSLOT;
      double* dL = ringL + k * LSLOT;
      const long long rr = (long long)r * a.ld, rm1 = (long long)std::max(r - 1, 0) * a.ld,
                      rm2 = (long long)std::max(r - 2, 0) * a.ld;
      auto cp = [&](double* dst, const double* src) { std::memcpy(dst, src, sizeof(double) * len); };
      cp(dS + S_Q * RW, a.q + colb + rr); cp(dL + L_V * RW, a.va + colb + rr);
      cp(dL + L_SGC * RW, a.sgc + colm + rr); cp(dL + L_SGV * RW, a.sgv + colm + rr);
      cp(dL + L_RGC * RW, a.rgc + colm + rr); cp(dS + S_SGU * RW, a.sgu + colm + rm1);
      cp(dS + S_U * RW, a.ua + colb + rm2);
      if (MASK & 1) { cp(dL + L_VM * RW, a.vm + colb + rr); cp(dS + S_UM * RW, a.um + colb + rm2); }
    };
    for (int r = rfirst; r < rfirst + 4 && r <= rlastp; ++r) issue(r);
    double psum = 0.0;
    for (int ra = rfirst; ra <= rlastp; ra += 2) {
      const int rows[2] = {ra, ra + 1};
      for (int t = 0; t < TB; ++t) {               // patch + phase 1 of both rows
        const int e = t + 3, j = jbase - 3 + t;
        for (int h = 0; h < 2; ++h) {
          const int r = rows[h], k = (r - rfirst) % DL;
          const int oS = (k % DS) * SSLOT, oL0 = k * LSLOT, oL2 = ((k + DL - 2) % DL) * LSLOT, oL3 = ((k + DL - 3) % DL) * LSLOT;
          RowPtrs& P = R[h][t];
          P.q = ringS + oS + S_Q * RW + e; P.u = ringS + oS + S_U * RW + e;
          P.um = ringS + oS + S_UM * RW + e; P.su1 = ringS + oS + S_SGU * RW + e;
          P.v0 = ringL + oL0 + L_V * RW + e; P.vm0 = ringL + oL0 + L_VM * RW + e;
          P.sgv0 = ringL + oL0 + L_SGV * RW + e; P.sgc0 = ringL + oL0 + L_SGC * RW + e;
          P.rg0 = ringL + oL0 + L_RGC * RW + e; P.sgc2 = ringL + oL2 + L_SGC * RW + e;
          P.v3 = ringL + oL3 + L_V * RW + e; P.vm3 = ringL + oL3 + L_VM * RW + e;
          P.sgv3 = ringL + oL3 + L_SGV * RW + e; P.sgc3 = ringL + oL3 + L_SGC * RW + e;
          P.rg3 = ringL + oL3 + L_RGC * RW + e;
        }
        double qn[2][1];
        for (int h = 0; h < 2; ++h) {
          const int r = rows[h], k = (r - rfirst) % DL;
          RowPtrs& P = R[h][t];
          qn[h][0] = P.q[0];
          if (a.apply_corr && j >= a.lo && j < a.hi && r >= a.lo && r < a.hi) {
            qn[h][0] = fma(P.sgc0[0], a.corr, qn[h][0]);
            ringS[(k % DS) * SSLOT + S_Q * RW + e] = qn[h][0];
          }
        }
        for (int h = 0; h < 2; ++h) {
          const int k = (rows[h] - rfirst) % DL;
          double qx[1];
          x_inner_k<RECON, SPLIT, MASK, WMAX>(k, L[t], X[h][t], R[h][t], qn[h], cdxw, qx);
          sX[h][e] = qx[0];
        }
      }
      for (int t = 0; t < TB; ++t) {               // barrier A; phase 2
        const int e = t + 3;
        for (int h = 0; h < 2; ++h) {
          double f[1], g[1], cf[1] = {0.0}, cg[1];
          yflux_pair<RECON, SPLIT, MASK, 0>(R[h][t].v0, R[h][t].vm0, R[h][t].sgv0, R[h][t].sgc0, R[h][t].q, cdyw, f, cf);
          yflux_pair<RECON, SPLIT, MASK, 0>(R[h][t].v3, R[h][t].vm3, R[h][t].sgv3, R[h][t].sgc3, sX[h] + e, cdyw, g, cg);
          F[h][t] = f[0]; G[h][t] = g[0]; CF[h][t] = cf[0];
        }
      }
      for (int h = 0; h < 2; ++h)
        for (int t = 0; t < TB; ++t) { sF[h][t + 3] = F[h][t]; sG[h][t + 3] = G[h][t]; sC[h][t + 3] = CF[h][t]; }
      // barrier B; the copies of step s+2 overwrite the slots of this step and of rows ra-4, ra-3
      if (ra + 4 <= rlastp) { issue(ra + 4); issue(ra + 5); }
      for (int t = 0; t < TB; ++t) {               // phase 3 of both rows
        const int e = t + 3, j = jbase - 3 + t;
        for (int h = 0; h < 2; ++h) {
          const int r = rows[h], k = (r - rfirst) % DL;
          double f[1] = {F[h][t]}, g[1] = {G[h][t]}, cf[1] = {CF[h][t]};
          double fn[1] = {sF[h][e + 1]}, gn[1] = {sG[h][e + 1]}, cfn[1] = {sC[h][e + 1]}, out[1], sdiv[1];
          x_outer_k<RECON, SPLIT, WMAX>(k, L[t], X[h][t], R[h][t], f, fn, g, gn, cf, cfn, out, sdiv);
          if (r >= r0 + 3 && r <= rlast && t >= 3 && j < jend) {
            a.qn[(long long)p * a.ps + JOFF + j + (long long)(r - 3) * a.ld] = out[0];
            L[t].psum += sdiv[0];
          }
        }
      }
    }
    for (int t = 0; t < TB; ++t) psum += L[t].psum;
    a.part[blk] = psum;
  }
}

template <int RECON, int SPLIT>
int run_block_mask(const Args& a, int mask, int TB) {
  if (a.circ == 2) {
    if (mask == 0) run_block_pair<RECON, SPLIT, 0>(a, TB);
    else if (mask == 1) run_block_pair<RECON, SPLIT, 1>(a, TB);
    else if (mask == 2) run_block_pair<RECON, SPLIT, 2>(a, TB);
    else return -1;
    return 0;
  }
  if (mask == 0) run_block<RECON, SPLIT, 0>(a, TB);
  else if (mask == 1) run_block<RECON, SPLIT, 1>(a, TB);
  else if (mask == 2) run_block<RECON, SPLIT, 2>(a, TB);
  else return -1;
  return 0;
}

template <int RECON, int SPLIT, int NW>
int run_mask(const Args& a, int mask) {
  if (mask == 0) run<RECON, SPLIT, 0, NW>(a);
  else if (mask == 1) run<RECON, SPLIT, 1, NW>(a);
  else if (mask == 2) run<RECON, SPLIT, 2, NW>(a);
  else return -1;
  return 0;
}
template <int NW>
int run_scheme(const Args& a, int recon, int split, int mask) {
#define CASE(R, S) if (recon == R && split == S) return run_mask<R, S, NW>(a, mask)
  CASE(3, 1); CASE(3, 2); CASE(3, 3); CASE(1, 1); CASE(1, 2); CASE(1, 3);
#undef CASE
  return -1;
}
}  // namespace

extern "C" {
// geometry helpers (mirror fused_setup in fused.cu)
int f3_emul_ld(int N) { return ((JOFF + N + 8 + 1 + 15) / 16) * 16; }
int f3_emul_grid(int N, int nw, int rows_per_chunk, int* nstrips, int* wcols, int* nchunks) {
  const int cap = f3::strip_capacity(nw);
  *nstrips = (N + cap - 1) / cap;
  *wcols = (N + *nstrips - 1) / *nstrips;
  *wcols += *wcols & 1;
  *nchunks = (N + rows_per_chunk - 1) / rows_per_chunk;
  return 6 * *nstrips * *nchunks * nw;
}
// v2b decomposition (csrc/fused2b.cu): TB threads per CTA, one column each; depth = rows in flight.
// Returns the number of partial sums written to part (one per CTA).
int f3_emul_step_block_impl(int circ, int N, int recon, int split, int mask, int TB, int depth, int rows_per_chunk, const double* q,
                       double* qn, const double* ua, const double* va, const double* um, const double* vm,
                       const double* sgc, const double* rgc, const double* sgu, const double* sgv, double* part,
                       double corr, int apply_corr, double cdx, double cdy, double ws) {
  Args a;
  a.circ = circ;
  a.N = N; a.P = N + 8; a.ld = f3_emul_ld(N); a.lo = 4; a.hi = N + 4;
  a.ps = (long long)(a.P + 1) * a.ld;
  a.q = q; a.qn = qn; a.ua = ua; a.va = va; a.um = um; a.vm = vm;
  a.sgc = sgc; a.rgc = rgc; a.sgu = sgu; a.sgv = sgv; a.part = part;
  a.corr = corr; a.apply_corr = apply_corr; a.cdx = cdx; a.cdy = cdy; a.ws = ws;
  a.rows_per_chunk = rows_per_chunk; a.depth = depth;
  a.nstrips = a.wcols = 0;
  if (depth < 1 || TB < 16) return -2;
#define CASE(R, S) if (recon == R && split == S) return run_block_mask<R, S>(a, mask, TB)
  CASE(3, 1); CASE(3, 2); CASE(3, 3); CASE(1, 1); CASE(1, 2); CASE(1, 3);
#undef CASE
  return -1;
}
#define STEP_BLOCK_ARGS \
  int N, int recon, int split, int mask, int TB, int depth, int rows_per_chunk, const double *q, double *qn, \
      const double *ua, const double *va, const double *um, const double *vm, const double *sgc, const double *rgc, \
      const double *sgu, const double *sgv, double *part, double corr, int apply_corr, double cdx, double cdy, double ws
#define STEP_BLOCK_PASS \
  N, recon, split, mask, TB, depth, rows_per_chunk, q, qn, ua, va, um, vm, sgc, rgc, sgu, sgv, part, corr, apply_corr, \
      cdx, cdy, ws
int f3_emul_step_block(STEP_BLOCK_ARGS) { return f3_emul_step_block_impl(0, STEP_BLOCK_PASS); }
// the const-slot march (MINB >= 30): rings of 3 and 6 rows, circular six-register windows
int f3_emul_step_block_circ(STEP_BLOCK_ARGS) {
  if (depth != 2) return -2;
  return f3_emul_step_block_impl(1, STEP_BLOCK_PASS);
}
// the two-row march (MINB >= 50); depth is ignored (rings of 4 and 8 slots)
int f3_emul_step_block_pair(STEP_BLOCK_ARGS) { return f3_emul_step_block_impl(2, STEP_BLOCK_PASS); }
// run only these CTAs (indices in the full grid) in the following f3_emul_step_block* calls; n = 0 resets
void f3_emul_set_block_list(const int* list, int n) { g_blk_list = n > 0 ? list : nullptr; g_blk_count = n; }
int f3_emul_block_grid(int N, int TB, int rows_per_chunk) {
  const int wmax = TB - 6;
  return 6 * ((N + wmax - 1) / wmax) * ((N + rows_per_chunk - 1) / rows_per_chunk);
}
// One launch of the step kernel.  Arrays are in DEVICE layout ([panel][i][ld], column j at
// j + 12); metric arrays hold one panel.  part has f3_emul_grid(...) entries.
int f3_emul_step(int N, int recon, int split, int mask, int nw, int depth, int rows_per_chunk, const double* q,
                 double* qn, const double* ua, const double* va, const double* um, const double* vm,
                 const double* sgc, const double* rgc, const double* sgu, const double* sgv, double* part,
                 double corr, int apply_corr, double cdx, double cdy, double ws) {
  Args a;
  a.N = N; a.P = N + 8; a.ld = f3_emul_ld(N); a.lo = 4; a.hi = N + 4;
  a.ps = (long long)(a.P + 1) * a.ld;
  a.q = q; a.qn = qn; a.ua = ua; a.va = va; a.um = um; a.vm = vm;
  a.sgc = sgc; a.rgc = rgc; a.sgu = sgu; a.sgv = sgv; a.part = part;
  a.corr = corr; a.apply_corr = apply_corr; a.cdx = cdx; a.cdy = cdy; a.ws = ws;
  a.rows_per_chunk = rows_per_chunk; a.depth = depth;
  int nchunks;
  f3_emul_grid(N, nw, rows_per_chunk, &a.nstrips, &a.wcols, &nchunks);
  if (depth < 1) return -2;
  if (nw == 3) return run_scheme<3>(a, recon, split, mask);
  if (nw == 4) return run_scheme<4>(a, recon, split, mask);
  if (nw == 2) return run_scheme<2>(a, recon, split, mask);
  return -3;
}
}
