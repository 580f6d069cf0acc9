// Panel-edge halo kernels: strip gather, copy fill, Lagrange (duo-grid) fill.
//
// Reference behaviour (file:line under /root/reference):
//   src/halo_data.py:15-400        the 24 strip orientations      -> HaloMaps
//   src/interpolation.py:154-314   two-phase Lagrange ghost fill  -> dg_phase1/2
//   src/interpolation.py:320-340   adjacent-panel copy fill       -> gather+scatter
// Compiled with -fmad=false: sums are evaluated left to right exactly like the
// numpy expressions, so the ghost cells agree with the reference bit for bit.
#include "pycs_common.cuh"

namespace {

struct Spec { int nb; int kind; int nops; int ops[2]; };
enum { ILO, IHI, JLO, JHI };
enum { OP_T, OP_F0, OP_F1 };

// src/halo_data.py:33-182 as data: [panel][side E,W,N,S]
const Spec kSpec[6][4] = {
    {{1, ILO, 0, {0, 0}}, {3, IHI, 0, {0, 0}}, {4, JLO, 0, {0, 0}}, {5, JHI, 0, {0, 0}}},
    {{2, ILO, 0, {0, 0}}, {0, IHI, 0, {0, 0}}, {4, IHI, 2, {OP_T, OP_F1}}, {5, IHI, 2, {OP_T, OP_F0}}},
    {{3, ILO, 0, {0, 0}}, {1, IHI, 0, {0, 0}}, {4, JHI, 2, {OP_F0, OP_F1}}, {5, JLO, 2, {OP_F1, OP_F0}}},
    {{0, ILO, 0, {0, 0}}, {2, IHI, 0, {0, 0}}, {4, ILO, 2, {OP_T, OP_F0}}, {5, ILO, 2, {OP_T, OP_F1}}},
    {{1, JHI, 2, {OP_F1, OP_T}}, {3, JHI, 2, {OP_T, OP_F1}}, {2, JHI, 2, {OP_F0, OP_F1}}, {0, JHI, 0, {0, 0}}},
    {{1, JLO, 2, {OP_T, OP_F1}}, {3, JLO, 2, {OP_T, OP_F0}}, {0, JLO, 0, {0, 0}}, {2, JLO, 2, {OP_F0, OP_F1}}},
};

}  // namespace

void pycs_build_halo_maps(const Geo& g, HaloMaps* maps) {
  for (int p = 0; p < 6; ++p)
    for (int s = 0; s < 4; ++s) {
      const Spec& sp = kSpec[p][s];
      SideMap m{};
      m.nb = sp.nb;
      int n0, n1;
      if (sp.kind == ILO || sp.kind == IHI) {
        m.ci = (sp.kind == ILO) ? g.lo : g.hi - PYCS_NG; m.ai = 1; m.bi = 0;
        m.cj = 0; m.aj = 0; m.bj = 1; n0 = PYCS_NG; n1 = g.P;
      } else {
        m.ci = 0; m.ai = 1; m.bi = 0;
        m.cj = (sp.kind == JLO) ? g.lo : g.hi - PYCS_NG; m.aj = 0; m.bj = 1; n0 = g.P; n1 = PYCS_NG;
      }
      for (int o = 0; o < sp.nops; ++o) {
        int op = sp.ops[o];
        if (op == OP_T) {
          int t = m.ai; m.ai = m.bi; m.bi = t;
          t = m.aj; m.aj = m.bj; m.bj = t;
          t = n0; n0 = n1; n1 = t;
          m.rot ^= 1;
        } else if (op == OP_F0) {
          m.ci += m.ai * (n0 - 1); m.ai = -m.ai;
          m.cj += m.aj * (n0 - 1); m.aj = -m.aj;
        } else {
          m.ci += m.bi * (n1 - 1); m.bi = -m.bi;
          m.cj += m.bj * (n1 - 1); m.bj = -m.bj;
        }
      }
      maps->m[p][s] = m;
    }
}

namespace {

__device__ __forceinline__ double halo_at(const double* __restrict__ q, const Geo& g,
                                          const SideMap& m, int a, int b) {
  return q[gidx(g, m.nb, m.ci + m.ai * a + m.bi * b, m.cj + m.aj * a + m.bj * b)];
}

// buf layout: [side][panel][4*P] ; E/W entries [a][b] = a*P+b (a<4), N/S [a][b] = a*4+b (a<P)
__global__ void gather_kernel(Geo g, HaloMaps maps, const double* __restrict__ fx,
                              const double* __restrict__ fy, double* __restrict__ buf) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int n = 4 * g.P;
  if (t >= n) return;
  int p = blockIdx.y, s = blockIdx.z;
  const SideMap& m = maps.m[p][s];
  int a, b;
  const double* src;
  if (s < 2) { a = t / g.P; b = t - a * g.P; src = m.rot ? fx : fy; }
  else       { a = t >> 2;  b = t & 3;       src = m.rot ? fy : fx; }
  buf[((long long)s * 6 + p) * n + t] = halo_at(src, g, m, a, b);
}

__global__ void scatter_copy_kernel(Geo g, double* __restrict__ fx, double* __restrict__ fy,
                                    const double* __restrict__ buf, int same) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int n = 4 * g.P;
  if (t >= n) return;
  int p = blockIdx.y, s = blockIdx.z;
  double v = buf[((long long)s * 6 + p) * n + t];
  if (s < 2) {                       // W/E strips go to Qy (src/interpolation.py:337-338)
    int a = t / g.P, b = t - a * g.P;
    // when Qx is Qy the S/N assignment (:339-340) overwrites the corners afterwards
    if (same && (b < g.lo || b >= g.hi)) return;
    int i = (s == SIDE_W) ? a : g.hi + a;
    fy[gidx(g, p, i, b)] = v;
  } else {                           // S/N strips go to Qx
    int a = t >> 2, b = t & 3;
    int j = (s == SIDE_S) ? b : g.hi + b;
    fx[gidx(g, p, a, j)] = v;
  }
}

// Phase 1 (src/interpolation.py:200-248): edge ghosts, interior extent only.
__global__ void dg_phase1_kernel(Geo g, HaloMaps maps, double* __restrict__ q,
                                 const int* __restrict__ kminE, const double* __restrict__ wE,
                                 int order) {
  int k = g.lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= g.hi) return;
  int gl = blockIdx.y;                 // ghost layer 0..3
  int p = blockIdx.z >> 2, s = blockIdx.z & 3;
  const SideMap& m = maps.m[p][s];
  // W and S tables are the E table flipped in the layer index (src/lagrange.py:144-152)
  int ge = (s == SIDE_E || s == SIDE_N) ? gl : PYCS_NG - 1 - gl;
  int km = kminE[ge * g.P + k];
  const double* w = wE + ((long long)ge * g.P + k) * order;
  double acc = 0.0;
  for (int l = 0; l < order; ++l) {
    double v = (s < 2) ? halo_at(q, g, m, gl, km + l) : halo_at(q, g, m, km + l, gl);
    acc = acc + v * w[l];
  }
  int i, j;
  if (s == SIDE_E) { i = g.hi + gl; j = k; }
  else if (s == SIDE_W) { i = gl; j = k; }
  else if (s == SIDE_N) { i = k; j = g.hi + gl; }
  else { i = k; j = gl; }
  q[gidx(g, p, i, j)] = acc;
}

// Phase 2 (src/interpolation.py:250-314): the four 4x4 corners from the E/W strips,
// which now contain the neighbours' phase-1 ghosts.
__global__ void dg_phase2_kernel(Geo g, HaloMaps maps, double* __restrict__ q,
                                 const int* __restrict__ kminE, const double* __restrict__ wE,
                                 int order) {
  int t = threadIdx.x;                 // 0..31: layer (4) x corner column (8)
  int gl = t >> 3;
  int c = t & 7;
  int k = (c < 4) ? c : g.hi + (c - 4);
  int p = blockIdx.x >> 1, s = blockIdx.x & 1;   // E or W
  const SideMap& m = maps.m[p][s];
  int ge = (s == SIDE_E) ? gl : PYCS_NG - 1 - gl;
  int km = kminE[ge * g.P + k];
  const double* w = wE + ((long long)ge * g.P + k) * order;
  double acc = 0.0;
  for (int l = 0; l < order; ++l) acc = acc + halo_at(q, g, m, gl, km + l) * w[l];
  int i = (s == SIDE_E) ? g.hi + gl : gl;
  q[gidx(g, p, i, k)] = acc;
}

}  // namespace

int k_halo_gather(pycs_handle h, const double* fx, const double* fy, double* buf) {
  int n = 4 * h->g.P;
  dim3 grid((n + 127) / 128, 6, 4);
  gather_kernel<<<grid, 128, 0, h->stream>>>(h->g, h->maps, fx, fy, buf);
  CKL(h);
  return 0;
}

int k_halo_scatter_copy(pycs_handle h, double* fx, double* fy, const double* buf) {
  int n = 4 * h->g.P;
  dim3 grid((n + 127) / 128, 6, 4);
  scatter_copy_kernel<<<grid, 128, 0, h->stream>>>(h->g, fx, fy, buf, fx == fy ? 1 : 0);
  CKL(h);
  return 0;
}

int k_dg_fill(pycs_handle h, double* q) {
  if (!h->kminE) {
    pycs_set_error("ET-DG ghost fill needs pycs_upload_lagrange first");
    return PYCS_ERR_STATE;
  }
  dim3 grid((h->g.N + 127) / 128, 4, 24);
  dg_phase1_kernel<<<grid, 128, 0, h->stream>>>(h->g, h->maps, q, h->kminE, h->wE, h->order);
  CKL(h);
  dg_phase2_kernel<<<12, 32, 0, h->stream>>>(h->g, h->maps, q, h->kminE, h->wE, h->order);
  CKL(h);
  return 0;
}
