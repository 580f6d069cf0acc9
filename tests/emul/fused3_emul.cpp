// Host emulator of the v3 fused step kernel -- TEST INFRASTRUCTURE.
//
// Compiles py-cubed-sphere_b200/csrc/fused3_core.cuh (the per-lane arithmetic the CUDA
// kernel runs, shared verbatim) with g++ and replays the kernel's decomposition on the
// CPU: CTAs = (panel, strip, chunk), consumer warps, 32 lanes x 2 columns, a D-slot ring
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

using namespace f3;

namespace {
constexpr int JOFF = 12;   // PYCS_JOFF

struct Args {
  int N, P, ld, lo, hi;
  long long ps;
  const double *q, *ua, *va, *um, *vm, *sgc, *rgc, *sgu, *sgv;
  double *qn, *part;
  double corr, cdx, cdy, ws;
  int apply_corr, rows_per_chunk, nstrips, wcols, depth;
};

template <int RECON, int SPLIT, int MASK, int NW>
void run(const Args& a) {
  constexpr int RW = RowWidth<NW>::value;
  constexpr int NARR = (MASK & 1) ? 9 : 7;
  constexpr int SLOT = NARR * RW;
  const int D = a.depth;
  const int nchunks = (a.N + a.rows_per_chunk - 1) / a.rows_per_chunk;
  std::vector<dbl2> ringbuf((size_t)(D * SLOT + NW * SXW) / 2 + 8);
  double* ring = reinterpret_cast<double*>(ringbuf.data());
  for (int blk = 0; blk < 6 * a.nstrips * nchunks; ++blk) {
    int b = blk;
    const int p = b % 6;
    b /= 6;
    const int strip = b % a.nstrips, chunk = b / a.nstrips;
    const int js0 = a.lo + strip * a.wcols;
    const int js1 = std::min(js0 + a.wcols, a.hi);
    const int r0 = a.lo + chunk * a.rows_per_chunk;
    const int r1 = std::min(r0 + a.rows_per_chunk, a.hi);
    const int rfirst = r0 - 3, rlast = r1 + 2;
    const int c0 = ((js0 - 3) & ~1) - 4;
    const int len = std::min(RW, a.ld - JOFF - c0) & ~1;
    const long long colb = (long long)p * a.ps + JOFF + c0, colm = JOFF + c0;
    for (int warp = 0; warp < NW; ++warp) {
      std::fill(ring, ring + D * SLOT + NW * SXW, 0.0);
      double* sxrow = ring + D * SLOT + warp * SXW;
      int cw0, us, ue;
      warp_columns(js0, js1, warp, cw0, us, ue);
      Lane L[32];
      for (int l = 0; l < 32; ++l) lane_init(L[l]);
      int o0 = 0, o1 = (D - 1) * SLOT, o2 = (D - 2) * SLOT, o3 = (D - 3) * SLOT;
      for (int r = rfirst; r <= rlast; ++r) {
        // producer: stage row r into its slot, patch the pending MF-PR term
        {
          const int s = (r - rfirst) % D;
          double* dst = ring + s * SLOT;
          const long long rr = (long long)r * a.ld;
          auto cp = [&](int arr, const double* src) { std::memcpy(dst + arr * RW, src, sizeof(double) * len); };
          cp(A_Q, a.q + colb + rr); cp(A_V, a.va + colb + rr);
          cp(A_SGC, a.sgc + colm + rr); cp(A_SGV, a.sgv + colm + rr); cp(A_RGC, a.rgc + colm + rr);
          cp(A_SGU, a.sgu + colm + rr); cp(A_U, a.ua + colb + rr);
          if (MASK & 1) { cp(A_VM, a.vm + colb + rr); cp(A_UM, a.um + colb + rr); }
          if (a.apply_corr && r >= a.lo && r < a.hi)
            for (int k = 0; k < len; ++k) {
              const int j = c0 + k;
              if (j >= a.lo && j < a.hi) dst[A_Q * RW + k] = fma(dst[A_SGC * RW + k], a.corr, dst[A_Q * RW + k]);
            }
          if (s * SLOT != o0) std::abort();
        }
        const double *R0 = ring + o0, *R1 = ring + o1, *R2 = ring + o2, *R3 = ring + o3;
        XEdge X[32];
        double F[32][NC + 1], G[32][NC + 1], CF[32][NC + 1], CG[NC];
        for (int l = 0; l < 32; ++l) {                 // phase 1
          double qx[NC];
          phase_x_inner<RECON, SPLIT, MASK, RW>(L[l], X[l], R0, R1, R2, R3, cw0 - c0 + NC * l, a.cdx, a.ws, qx);
          st2(sxrow + 4 + NC * l, qx[0], qx[1]);
        }
        for (int l = 0; l < 32; ++l) {                 // phase 2 (after __syncwarp)
          const int ca = cw0 - c0 + NC * l;
          CF[l][0] = CF[l][1] = 0.0;
          yflux_pair<RECON, SPLIT, MASK, RW>(R0, ca, R0 + A_Q * RW + ca, a.cdy, a.ws, F[l], CF[l]);
          yflux_pair<RECON, SPLIT, MASK, RW>(R3, ca, sxrow + 4 + NC * l, a.cdy, a.ws, G[l], CG);
        }
        for (int l = 0; l < 32; ++l) {                 // __shfl_down(.., 1): lane 31 keeps its own value
          const int n = l < 31 ? l + 1 : l;
          F[l][NC] = F[n][0]; G[l][NC] = G[n][0]; CF[l][NC] = CF[n][0];
        }
        for (int l = 0; l < 32; ++l) {                 // phase 3
          const int ca = cw0 - c0 + NC * l, col = cw0 + NC * l;
          double out[NC], sdiv[NC];
          phase_x_outer<RECON, SPLIT, RW>(L[l], X[l], R0, ca, F[l], G[l], CF[l], out, sdiv);
          if (r >= r0 + 3) {
            double* QN = a.qn + (long long)p * a.ps + JOFF + (long long)(r - 3) * a.ld;
            for (int c = 0; c < NC; ++c)
              if (col + c >= us && col + c < ue) { QN[col + c] = out[c]; L[l].psum += sdiv[c]; }
          }
        }
        o3 = o2; o2 = o1; o1 = o0;
        o0 = (o0 + SLOT == D * SLOT) ? 0 : o0 + SLOT;
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
  const int cap = strip_capacity(nw);
  *nstrips = (N + cap - 1) / cap;
  *wcols = (N + *nstrips - 1) / *nstrips;
  *wcols += *wcols & 1;
  *nchunks = (N + rows_per_chunk - 1) / rows_per_chunk;
  return 6 * *nstrips * *nchunks * nw;
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
  if (depth < 5) return -2;
  if (nw == 3) return run_scheme<3>(a, recon, split, mask);
  if (nw == 4) return run_scheme<4>(a, recon, split, mask);
  if (nw == 2) return run_scheme<2>(a, recon, split, mask);
  return -3;
}
}
