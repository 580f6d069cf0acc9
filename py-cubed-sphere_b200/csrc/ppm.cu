// Operator-by-operator kernels of the split PPM divergence.  One kernel per
// reference operator so that every intermediate the reference exposes
// (px.q_L, px.f_upw, px.dF, simulation.div ...) exists on the device and can be
// compared with the oracle.  Reference (file:line under /root/reference):
//   src/reconstruction_1d.py:26-372   ppm_reconstruction_x / _y    -> recon_kernel
//   src/edges_treatment.py:82-190     edges_extrapolation           -> extrap/avg/ghost kernels
//   src/flux.py:20-128                numerical_flux_ppm_x / _y     -> coeff_kernel, upwind_kernel
//   src/discrete_operators.py:109-138 F_operator / G_operator       -> flux_diff_kernel
//   src/discrete_operators.py:45-73   inner (splitting) update      -> inner_update_kernel
//   src/edges_treatment.py:231-278    average_flux_cube_edges       -> avg_flux_kernel
//   src/discrete_operators.py:95-101  divergence + MF-PR projection -> div_kernel, project_kernel
// DIR = 0: sweep along i (x direction, rows, stride ld); DIR = 1: along j.
// Compiled with -fmad=false so each expression rounds like the numpy original.
// The production path is the fused kernel in fused.cu; this file is the
// operator surface + its cross-check.
#include "pycs_common.cuh"
#include "cube_edges.cuh"

namespace {

constexpr int BX = 128;

__device__ __forceinline__ double sgn(double x) { return (x > 0.0) ? 1.0 : ((x < 0.0) ? -1.0 : 0.0); }

// ---------------------------------------------------------------- reconstruction
template <int DIR>
__global__ void recon_kernel(Geo g, int recon, const double* __restrict__ q,
                             double* __restrict__ qL, double* __restrict__ qR) {
  int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y, p = blockIdx.z;
  if (j >= g.P || i >= g.P) return;
  int s = DIR == 0 ? i : j;
  if (s < g.lo - 1 || s > g.hi) return;
  const long long st = DIR == 0 ? g.ld : 1;
  const long long id = gidx(g, p, i, j);
  const double* c = q + id;
#define Q_(k) c[(k) * st]
  double l, r;
  if (recon == 3) {                       // PPM-PL07, :46-62
    const double a1 = 2.0 / 60.0, a2 = -13.0 / 60.0, a3 = 47.0 / 60.0, a4 = 27.0 / 60.0, a5 = -3.0 / 60.0;
    double q1 = Q_(-2), q2 = Q_(-1), q3 = Q_(0), q4 = Q_(1), q5 = Q_(2);
    r = a1 * q1 + a2 * q2 + a3 * q3 + a4 * q4 + a5 * q5;
    l = a5 * q1 + a4 * q2 + a3 * q3 + a2 * q4 + a1 * q5;
  } else if (recon == 1) {                // PPM-0, :36-44
    l = (7.0 / 12.0) * (Q_(0) + Q_(-1)) - (Q_(1) + Q_(-2)) / 12.0;
    r = (7.0 / 12.0) * (Q_(1) + Q_(0)) - (Q_(2) + Q_(-1)) / 12.0;
  } else if (recon == 2) {                // PPM-CW84, :64-153
    double dQ[3];
#pragma unroll
    for (int k = -1; k <= 1; ++k) {
      double qm = Q_(k - 1), q0 = Q_(k), qp = Q_(k + 1);
      double d0 = 0.5 * (qp - qm), d1 = 2.0 * (qp - q0), d2 = 2.0 * (q0 - qm);
      double d = fmin(fmin(fabs(d0), fabs(d1)), fabs(d2)) * sgn(d0);
      if (!((qp - q0) * (q0 - qm) > 0.0)) d = 0.0;
      dQ[k + 1] = d;
    }
    l = 0.5 * (Q_(0) + Q_(-1)) - (dQ[1] - dQ[0]) / 6.0;
    r = 0.5 * (Q_(1) + Q_(0)) - (dQ[2] - dQ[1]) / 6.0;
    double qc = Q_(0);
    double dq = r - l, q6 = 6 * qc - 3 * (r + l);       // pre-limiter values (:118-119)
    if ((r - qc) * (qc - l) <= 0) { r = qc; l = qc; }   // :124-128
    bool over = fabs(dq) < fabs(q6);                    // :133-134
    double rl = r - l, mid = qc - 0.5 * (r + l);
    bool left = rl * mid > (rl * rl) / 6.0;             // :140
    bool right = -(rl * rl) / 6.0 > rl * mid;           // :144
    if (over && left) l = 3.0 * qc - 2.0 * r;           // :149-150
    if (over && right) r = 3.0 * qc - 2.0 * l;          // :151-153 (sees the new l)
  } else {                                // PPM-L04, :155-192
    double mono[3];
#pragma unroll
    for (int k = -1; k <= 1; ++k) {
      double qm = Q_(k - 1), q0 = Q_(k), qp = Q_(k + 1);
      double d = 0.25 * (qp - qm);
      double dmin = fmax(fmax(qm, q0), qp) - q0;
      double dmax = q0 - fmin(fmin(qm, q0), qp);
      mono[k + 1] = fmin(fmin(fabs(d), dmin), dmax) * sgn(d);
    }
    l = 0.5 * (Q_(0) + Q_(-1)) - (mono[1] - mono[0]) / 3.0;
    r = 0.5 * (Q_(1) + Q_(0)) - (mono[2] - mono[1]) / 3.0;
    double q0 = Q_(0), m = mono[1];
    double qmin = fmin(2.0 * fabs(m), fabs(l - q0)) * sgn(2.0 * m);
    l = q0 - qmin;
    qmin = fmin(2.0 * fabs(m), fabs(r - q0)) * sgn(2.0 * m);
    r = q0 + qmin;
  }
#undef Q_
  qL[id] = l;
  qR[id] = r;
}

// ---------------------------------------------------------------- ET-PL07 edge treatment
struct Par { double* qL; double* qR; };

// element t (0..N-1 along the edge) of the line at sweep index `idx` of array a
__device__ __forceinline__ long long line_at(const Geo& g, int dir, int panel, int idx, int t) {
  return dir == 0 ? gidx(g, panel, idx, g.lo + t) : gidx(g, panel, g.lo + t, idx);
}

// src/edges_treatment.py:91-115: one-sided values next to every panel edge
__global__ void extrap_kernel(Geo g, int recon, const double* __restrict__ qx,
                              const double* __restrict__ qy, Par px, Par py) {
  int t = blockIdx.x * BX + threadIdx.x;
  if (t >= g.N) return;
  int p = blockIdx.y, dir = blockIdx.z >> 1, side = blockIdx.z & 1;
  const double* q = dir == 0 ? qx : qy;
  Par par = dir == 0 ? px : py;
  int lo = g.lo, hi = g.hi;
#define AT(idx) line_at(g, dir, p, (idx), t)
  if (side == 0) {
    double q0 = q[AT(lo)], q1 = q[AT(lo + 1)], q2 = q[AT(lo + 2)];
    par.qL[AT(lo)] = 1.5 * q0 - 0.5 * q1;                                   // eq. 47
    double e = (3.0 * q0 + 11.0 * q1 - 2.0 * (q2 - q0)) / 14.0;             // eq. 49
    par.qR[AT(lo)] = e;
    par.qL[AT(lo + 1)] = e;
    if (recon == 3) par.qR[AT(lo + 1)] = par.qL[AT(lo + 2)];
  } else {
    double q0 = q[AT(hi - 1)], q1 = q[AT(hi - 2)], q2 = q[AT(hi - 3)];
    par.qR[AT(hi - 1)] = 1.5 * q0 - 0.5 * q1;
    double e = (3.0 * q0 + 11.0 * q1 - 2.0 * (q2 - q0)) / 14.0;
    par.qL[AT(hi - 1)] = e;
    par.qR[AT(hi - 2)] = e;
    if (recon == 3) par.qL[AT(hi - 2)] = par.qR[AT(hi - 3)];
  }
#undef AT
}

// boundary-cell edge value facing the cube edge: q_L at lo, q_R at hi-1
__device__ __forceinline__ double* facing(const Geo& g, const EdgeEnd& e, Par px, Par py, int t) {
  Par par = e.dir == 0 ? px : py;
  return e.side == 0 ? par.qL + line_at(g, e.dir, e.panel, g.lo, t)
                     : par.qR + line_at(g, e.dir, e.panel, g.hi - 1, t);
}

// src/edges_treatment.py:31-76
__global__ void avg_parabola_kernel(Geo g, Par px, Par py) {
  int t = blockIdx.x * BX + threadIdx.x;
  if (t >= g.N) return;
  CubeEdge ce = cube_edge(blockIdx.y);
  int tb = ce.flip ? g.N - 1 - t : t;
  double* a = facing(g, ce.a, px, py, t);
  double* b = facing(g, ce.b, px, py, tb);
  double v = (*a + *b) * 0.5;
  *a = v;
  *b = v;
}

// src/edges_treatment.py:120-190: parabolas of the first ghost cell beyond each edge
__global__ void ghost_parabola_kernel(Geo g, Par px, Par py) {
  int t = blockIdx.x * BX + threadIdx.x;
  if (t >= g.N) return;
  CubeEdge ce = cube_edge(blockIdx.y);
  int swap = ce.a.side == ce.b.side;
  for (int d = 0; d < 2; ++d) {
    EdgeEnd dst = d == 0 ? ce.a : ce.b, src = d == 0 ? ce.b : ce.a;
    int td = t, ts = ce.flip ? g.N - 1 - t : t;
    Par pd = dst.dir == 0 ? px : py, psrc = src.dir == 0 ? px : py;
    long long di = line_at(g, dst.dir, dst.panel, dst.side == 0 ? g.lo - 1 : g.hi, td);
    long long si = line_at(g, src.dir, src.panel, src.side == 0 ? g.lo : g.hi - 1, ts);
    double sl = psrc.qL[si], sr = psrc.qR[si];
    pd.qL[di] = swap ? sr : sl;
    pd.qR[di] = swap ? sl : sr;
  }
}

// ---------------------------------------------------------------- flux
// src/flux.py:27-44: metric weighting (in place) and parabola coefficients
template <int DIR>
__global__ void coeff_kernel(Geo g, int mt, const double* __restrict__ qa, double* __restrict__ qL,
                             double* __restrict__ qR, double* __restrict__ dq, double* __restrict__ q6,
                             const double* __restrict__ sgc, const double* __restrict__ sge) {
  int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y, p = blockIdx.z;
  if (j >= g.P || i >= g.P) return;
  int s = DIR == 0 ? i : j;
  if (s < g.lo - 1 || s > g.hi) return;
  const long long st = DIR == 0 ? g.ld : 1;
  long long id = gidx(g, p, i, j), id0 = gidx(g, 0, i, j);
  double l = qL[id], r = qR[id], q = qa[id];
  if (mt == 1) {
    l = l * sge[id0];
    r = r * sge[id0 + st];
    q = q * sgc[id0];
    qL[id] = l;
    qR[id] = r;
  }
  dq[id] = r - l;
  q6[id] = 6 * q - 3 * (r + l);
}

// src/flux.py:46-70: upwind flux at edges lo..hi
template <int DIR>
__global__ void upwind_kernel(Geo g, int mt, const double* __restrict__ qL, const double* __restrict__ qR,
                              const double* __restrict__ dq, const double* __restrict__ q6,
                              const double* __restrict__ cfl, const double* __restrict__ umask,
                              const double* __restrict__ uavg, const double* __restrict__ sge,
                              double* __restrict__ fL, double* __restrict__ fR, double* __restrict__ fup) {
  int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y, p = blockIdx.z;
  int ni = DIR == 0 ? g.P + 1 : g.P, nj = DIR == 0 ? g.P : g.P + 1;
  if (j >= nj || i >= ni) return;
  int s = DIR == 0 ? i : j;
  if (s < g.lo || s > g.hi) return;
  const long long st = DIR == 0 ? g.ld : 1;
  long long id = gidx(g, p, i, j);
  double c = cfl[id];
  double f;
  if (umask[id] >= 0) {                  // U_pu.upos, src/averaged_velocity.py:21-23
    long long ic = id - st;
    f = qR[ic] + c * 0.5 * (q6[ic] - dq[ic]) - q6[ic] * c * c / 3.0;
    fL[id] = f;
  } else {
    f = qL[id] - c * 0.5 * (q6[id] + dq[id]) - q6[id] * c * c / 3.0;
    fR[id] = f;
  }
  f = f * uavg[id];
  if (mt == 2) f = f * sge[gidx(g, 0, i, j)];
  fup[id] = f;
}

// src/discrete_operators.py:109-138
template <int DIR>
__global__ void flux_diff_kernel(Geo g, const double* __restrict__ fup, double* __restrict__ dF) {
  int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y, p = blockIdx.z;
  if (j >= g.P || i >= g.P) return;
  int s = DIR == 0 ? i : j;
  if (s < g.lo || s >= g.hi) return;
  const long long st = DIR == 0 ? g.ld : 1;
  long long id = gidx(g, p, i, j);
  double d = -(fup[id + st] - fup[id]);
  dF[id] = d * g.dt / (DIR == 0 ? g.dx : g.dy);
}

__global__ void scale_kernel(Geo g, double* __restrict__ dst, const double* __restrict__ src,
                             int ni, int nj, double num, double den) {
  int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y, p = blockIdx.z;
  if (j >= nj || i >= ni) return;
  long long id = gidx(g, p, i, j);
  dst[id] = src[id] * num / den;        // cfl_x: u*dt/dx (src/cfl.py:10)
}

__global__ void mul_metric_kernel(Geo g, double* __restrict__ gq, const double* __restrict__ q,
                                  const double* __restrict__ sgc) {
  int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y, p = blockIdx.z;
  if (j >= g.P) return;
  long long id = gidx(g, p, i, j);
  gq[id] = q[id] * sgc[gidx(g, 0, i, j)];
}

// src/discrete_operators.py:45-73
__global__ void inner_update_kernel(Geo g, int split, const double* __restrict__ Q,
                                    const double* __restrict__ gQ, const double* __restrict__ dFx,
                                    const double* __restrict__ dFy, const double* __restrict__ cx,
                                    const double* __restrict__ cy, const double* __restrict__ sgc,
                                    const double* __restrict__ sgu, const double* __restrict__ sgv,
                                    double* __restrict__ Qx, double* __restrict__ Qy) {
  int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y, p = blockIdx.z;
  if (j >= g.P) return;
  long long id = gidx(g, p, i, j), id0 = gidx(g, 0, i, j);
  double q = Q[id], gq = gQ[id], m = sgc[id0];
  double rx, ry;
  if (split == 1) {
    rx = (gq + 0.5 * dFx[id]) / m;
    ry = (gq + 0.5 * dFy[id]) / m;
  } else {
    double c1x = sgu[id0 + g.ld] * cx[id + g.ld], c2x = sgu[id0] * cx[id];
    double c1y = sgv[id0 + 1] * cy[id + 1], c2y = sgv[id0] * cy[id];
    if (split == 2) {
      rx = (gq + 0.5 * dFx[id] + 0.5 * (c1x - c2x) * q) / m;
      ry = (gq + 0.5 * dFy[id] + 0.5 * (c1y - c2y) * q) / m;
    } else {
      rx = 0.5 * (q + (q + dFx[id]) / (1.0 - (c1x - c2x)));
      ry = 0.5 * (q + (q + dFy[id]) / (1.0 - (c1y - c2y)));
    }
  }
  Qx[id] = rx;
  Qy[id] = ry;
}

// src/edges_treatment.py:231-278
__global__ void avg_flux_kernel(Geo g, double* __restrict__ fx, double* __restrict__ fy) {
  int t = blockIdx.x * BX + threadIdx.x;
  if (t >= g.N) return;
  CubeEdge ce = cube_edge(blockIdx.y);
  int tb = ce.flip ? g.N - 1 - t : t;
  double sg = ce.a.side == ce.b.side ? -1.0 : 1.0;
  double* a = (ce.a.dir == 0 ? fx : fy) + line_at(g, ce.a.dir, ce.a.panel, ce.a.side == 0 ? g.lo : g.hi, t);
  double* b = (ce.b.dir == 0 ? fx : fy) + line_at(g, ce.b.dir, ce.b.panel, ce.b.side == 0 ? g.lo : g.hi, tb);
  double v = 0.5 * (*a) + sg * (0.5 * (*b));
  *a = v;
  *b = sg * v;
}

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ double block_sum(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = (l < (blockDim.x >> 5)) ? sh[l] : 0.0;
    for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
  }
  __syncthreads();
  return r;
}
__device__ __forceinline__ double block_max(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = (l < (blockDim.x >> 5)) ? sh[l] : 0.0;
    for (int o = 16; o > 0; o >>= 1) r = fmax(r, __shfl_down_sync(0xffffffffu, r, o));
  }
  __syncthreads();
  return r;
}

// div = -(dFx+dFy)/(dt*sqrtg) on the whole panel (:95) + per-block sum(div*sqrtg) (:99)
__global__ void div_kernel(Geo g, const double* __restrict__ dFx, const double* __restrict__ dFy,
                           const double* __restrict__ sgc, double* __restrict__ div,
                           double* __restrict__ part) {
  __shared__ double sh[32];
  int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y, p = blockIdx.z;
  double contrib = 0.0;
  if (j < g.P) {
    long long id = gidx(g, p, i, j);
    double m = sgc[gidx(g, 0, i, j)];
    double d = -(dFx[id] + dFy[id]) / (g.dt * m);
    div[id] = d;
    if (i >= g.lo && i < g.hi && j >= g.lo && j < g.hi) contrib = d * m;
  }
  double s = block_sum(contrib, sh);
  if (threadIdx.x == 0) part[(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = s;
}

// fixed-order final reduction of n partials into out[slot]
__global__ void final_sum_kernel(const double* __restrict__ part, int n, double* __restrict__ out, int slot) {
  __shared__ double sh[32];
  double v = 0.0;
  for (int k = threadIdx.x; k < n; k += blockDim.x) v += part[k];
  double s = block_sum(v, sh);
  if (threadIdx.x == 0) out[slot] = s;
}
__global__ void final_max_kernel(const double* __restrict__ part, int n, double* __restrict__ out, int slot) {
  __shared__ double sh[32];
  double v = 0.0;
  for (int k = threadIdx.x; k < n; k += blockDim.x) v = fmax(v, part[k]);
  double s = block_max(v, sh);
  if (threadIdx.x == 0) out[slot] = s;
}

// div[int] -= sqrtg * m0 / a2 (:101)
__global__ void project_kernel(Geo g, double* __restrict__ div, const double* __restrict__ sgc,
                               const double* __restrict__ m0, double a2) {
  int j = g.lo + blockIdx.x * BX + threadIdx.x, i = g.lo + blockIdx.y, p = blockIdx.z;
  if (j >= g.hi) return;
  long long id = gidx(g, p, i, j);
  div[id] = div[id] - sgc[gidx(g, 0, i, j)] * (*m0) / a2;
}

// Q[int] -= dt*div (src/advection_timestep.py:43)
__global__ void q_update_kernel(Geo g, double* __restrict__ Q, const double* __restrict__ div) {
  int j = g.lo + blockIdx.x * BX + threadIdx.x, i = g.lo + blockIdx.y, p = blockIdx.z;
  if (j >= g.hi) return;
  long long id = gidx(g, p, i, j);
  Q[id] = Q[id] - g.dt * div[id];
}

// mode 0: sum sqrtg^2; 1: sum Q*sqrtg*dx*dy; 2: |qe-Q| max; 3: sum |qe-Q|; 4: sum (qe-Q)^2
__global__ void interior_reduce_kernel(Geo g, int mode, const double* __restrict__ Q,
                                       const double* __restrict__ sgc, const double* __restrict__ qe,
                                       double* __restrict__ part) {
  __shared__ double sh[32];
  int j = g.lo + blockIdx.x * BX + threadIdx.x, i = g.lo + blockIdx.y, p = blockIdx.z;
  double v = 0.0;
  if (j < g.hi) {
    long long id = gidx(g, p, i, j);
    if (mode == 0) { double m = sgc[gidx(g, 0, i, j)]; v = m * m; }
    else if (mode == 1) v = Q[id] * sgc[gidx(g, 0, i, j)] * g.dx * g.dy;
    else {
      double e = fabs(qe[id] - Q[id]);
      v = (mode == 4) ? e * e : e;
    }
  }
  double s = (mode == 2) ? block_max(v, sh) : block_sum(v, sh);
  if (threadIdx.x == 0) part[(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = s;
}

inline dim3 grid_all(const Geo& g, int ni, int nj) { return dim3((nj + BX - 1) / BX, ni, 6); }
inline dim3 grid_int(const Geo& g) { return dim3((g.N + BX - 1) / BX, g.N, 6); }

}  // namespace

#define F(h, id, var)    \
  double* var = nullptr; \
  TRY(pycs_field_ptr(h, id, &var))

int k_cfl(pycs_handle h, double* dst, const double* src, int dir) {
  const Geo& g = h->g;
  int ni = dir == 0 ? g.P + 1 : g.P, nj = dir == 0 ? g.P : g.P + 1;
  scale_kernel<<<grid_all(g, ni, nj), BX, 0, h->stream>>>(g, dst, src, ni, nj, g.dt, dir == 0 ? g.dx : g.dy);
  CKL(h);
  return 0;
}

int k_mul_metric(pycs_handle h, double* gq, const double* q) {
  F(h, PYCS_F_SQRTG_PC, sgc);
  mul_metric_kernel<<<grid_all(h->g, h->g.P, h->g.P), BX, 0, h->stream>>>(h->g, gq, q, sgc);
  CKL(h);
  return 0;
}

int k_recon(pycs_handle h, const double* qx, const double* qy) {
  const Geo& g = h->g;
  F(h, PYCS_F_PX_QL, xl); F(h, PYCS_F_PX_QR, xr); F(h, PYCS_F_PY_QL, yl); F(h, PYCS_F_PY_QR, yr);
  recon_kernel<0><<<grid_all(g, g.P, g.P), BX, 0, h->stream>>>(g, h->prm.recon, qx, xl, xr);
  CKL(h);
  recon_kernel<1><<<grid_all(g, g.P, g.P), BX, 0, h->stream>>>(g, h->prm.recon, qy, yl, yr);
  CKL(h);
  return 0;
}

int k_edges_extrapolation(pycs_handle h, const double* qx, const double* qy) {
  const Geo& g = h->g;
  F(h, PYCS_F_PX_QL, xl); F(h, PYCS_F_PX_QR, xr); F(h, PYCS_F_PY_QL, yl); F(h, PYCS_F_PY_QR, yr);
  Par px{xl, xr}, py{yl, yr};
  int nb = (g.N + BX - 1) / BX;
  extrap_kernel<<<dim3(nb, 6, 4), BX, 0, h->stream>>>(g, h->prm.recon, qx, qy, px, py);
  CKL(h);
  avg_parabola_kernel<<<dim3(nb, 12), BX, 0, h->stream>>>(g, px, py);
  CKL(h);
  ghost_parabola_kernel<<<dim3(nb, 12), BX, 0, h->stream>>>(g, px, py);
  CKL(h);
  return 0;
}

int k_flux(pycs_handle h, const double* qx, const double* qy) {
  const Geo& g = h->g;
  F(h, PYCS_F_PX_QL, xl); F(h, PYCS_F_PX_QR, xr); F(h, PYCS_F_PX_DQ, xdq); F(h, PYCS_F_PX_Q6, xq6);
  F(h, PYCS_F_PY_QL, yl); F(h, PYCS_F_PY_QR, yr); F(h, PYCS_F_PY_DQ, ydq); F(h, PYCS_F_PY_Q6, yq6);
  F(h, PYCS_F_PX_FL, xfl); F(h, PYCS_F_PX_FR, xfr); F(h, PYCS_F_PX_FUPW, xfu);
  F(h, PYCS_F_PY_FL, yfl); F(h, PYCS_F_PY_FR, yfr); F(h, PYCS_F_PY_FUPW, yfu);
  F(h, PYCS_F_SQRTG_PC, sgc); F(h, PYCS_F_SQRTG_PU, sgu); F(h, PYCS_F_SQRTG_PV, sgv);
  F(h, PYCS_F_CX, cx); F(h, PYCS_F_CY, cy);
  F(h, PYCS_F_PU_UCONTRA, uc); F(h, PYCS_F_PV_VCONTRA, vc);
  F(h, PYCS_F_PU_UAVG, ua); F(h, PYCS_F_PV_VAVG, va);
  int mt = h->prm.mt;
  coeff_kernel<0><<<grid_all(g, g.P, g.P), BX, 0, h->stream>>>(g, mt, qx, xl, xr, xdq, xq6, sgc, sgu);
  CKL(h);
  coeff_kernel<1><<<grid_all(g, g.P, g.P), BX, 0, h->stream>>>(g, mt, qy, yl, yr, ydq, yq6, sgc, sgv);
  CKL(h);
  upwind_kernel<0><<<grid_all(g, g.P + 1, g.P), BX, 0, h->stream>>>(g, mt, xl, xr, xdq, xq6, cx, uc, ua, sgu, xfl, xfr, xfu);
  CKL(h);
  upwind_kernel<1><<<grid_all(g, g.P, g.P + 1), BX, 0, h->stream>>>(g, mt, yl, yr, ydq, yq6, cy, vc, va, sgv, yfl, yfr, yfu);
  CKL(h);
  return 0;
}

int k_flux_diff(pycs_handle h, int dir) {
  const Geo& g = h->g;
  if (dir == 0) {
    F(h, PYCS_F_PX_FUPW, fu); F(h, PYCS_F_PX_DF, df);
    flux_diff_kernel<0><<<grid_all(g, g.P, g.P), BX, 0, h->stream>>>(g, fu, df);
  } else {
    F(h, PYCS_F_PY_FUPW, fu); F(h, PYCS_F_PY_DF, df);
    flux_diff_kernel<1><<<grid_all(g, g.P, g.P), BX, 0, h->stream>>>(g, fu, df);
  }
  CKL(h);
  return 0;
}

int k_inner_update(pycs_handle h) {
  const Geo& g = h->g;
  F(h, PYCS_F_Q, Q); F(h, PYCS_F_GQ, gQ); F(h, PYCS_F_PX_DF, dfx); F(h, PYCS_F_PY_DF, dfy);
  F(h, PYCS_F_CX, cx); F(h, PYCS_F_CY, cy); F(h, PYCS_F_QX, Qx); F(h, PYCS_F_QY, Qy);
  F(h, PYCS_F_SQRTG_PC, sgc); F(h, PYCS_F_SQRTG_PU, sgu); F(h, PYCS_F_SQRTG_PV, sgv);
  inner_update_kernel<<<grid_all(g, g.P, g.P), BX, 0, h->stream>>>(g, h->prm.opsplit, Q, gQ, dfx, dfy, cx, cy,
                                                                   sgc, sgu, sgv, Qx, Qy);
  CKL(h);
  return 0;
}

int k_average_flux_edges(pycs_handle h) {
  const Geo& g = h->g;
  F(h, PYCS_F_PX_FUPW, fx); F(h, PYCS_F_PY_FUPW, fy);
  avg_flux_kernel<<<dim3((g.N + BX - 1) / BX, 12), BX, 0, h->stream>>>(g, fx, fy);
  CKL(h);
  return 0;
}

static int ensure_partials(pycs_handle h, int n) {
  if (h->red_blocks >= n) return 0;
  if (h->red_part) cudaFree(h->red_part);
  CK(cudaMalloc(&h->red_part, sizeof(double) * n));
  h->red_blocks = n;
  return 0;
}

int k_sum_sq_metric(pycs_handle h, double* out_host) {
  const Geo& g = h->g;
  F(h, PYCS_F_SQRTG_PC, sgc);
  dim3 gr = grid_int(g);
  int n = gr.x * gr.y * gr.z;
  TRY(ensure_partials(h, n));
  interior_reduce_kernel<<<gr, BX, 0, h->stream>>>(g, 0, nullptr, sgc, nullptr, h->red_part);
  CKL(h);
  final_sum_kernel<<<1, 1024, 0, h->stream>>>(h->red_part, n, h->red_out, 1);
  CKL(h);
  CK(cudaMemcpyAsync(out_host, h->red_out + 1, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int k_div_and_fix(pycs_handle h) {
  const Geo& g = h->g;
  F(h, PYCS_F_PX_DF, dfx); F(h, PYCS_F_PY_DF, dfy); F(h, PYCS_F_DIV, div); F(h, PYCS_F_SQRTG_PC, sgc);
  dim3 gr = grid_all(g, g.P, g.P);
  int n = gr.x * gr.y * gr.z;
  TRY(ensure_partials(h, n));
  if (h->prm.mf == 3 && !h->a2_valid) {
    TRY(k_sum_sq_metric(h, &h->a2));
    h->a2_valid = 1;
  }
  div_kernel<<<gr, BX, 0, h->stream>>>(g, dfx, dfy, sgc, div, h->red_part);
  CKL(h);
  if (h->prm.mf == 3) {
    final_sum_kernel<<<1, 1024, 0, h->stream>>>(h->red_part, n, h->red_out, 0);
    CKL(h);
    project_kernel<<<grid_int(g), BX, 0, h->stream>>>(g, div, sgc, h->red_out, h->a2);
    CKL(h);
  }
  return 0;
}

int k_q_update(pycs_handle h) {
  F(h, PYCS_F_Q, Q); F(h, PYCS_F_DIV, div);
  q_update_kernel<<<grid_int(h->g), BX, 0, h->stream>>>(h->g, Q, div);
  CKL(h);
  return 0;
}

int k_errors(pycs_handle h, const double* qexact_dev, double* out3_host) {
  const Geo& g = h->g;
  F(h, PYCS_F_Q, Q);
  dim3 gr = grid_int(g);
  int n = gr.x * gr.y * gr.z;
  TRY(ensure_partials(h, n));
  for (int mode = 2; mode <= 4; ++mode) {
    interior_reduce_kernel<<<gr, BX, 0, h->stream>>>(g, mode, Q, nullptr, qexact_dev, h->red_part);
    CKL(h);
    if (mode == 2) final_max_kernel<<<1, 1024, 0, h->stream>>>(h->red_part, n, h->red_out, 2);
    else final_sum_kernel<<<1, 1024, 0, h->stream>>>(h->red_part, n, h->red_out, mode);
    CKL(h);
  }
  double r[3];
  CK(cudaMemcpyAsync(r, h->red_out + 2, 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  double cnt = 6.0 * g.N * g.N;
  out3_host[0] = r[0];
  out3_host[1] = r[1] / cnt;
  out3_host[2] = sqrt(r[2] / cnt);
  return 0;
}

int k_mass(pycs_handle h, double* out_host) {
  const Geo& g = h->g;
  F(h, PYCS_F_Q, Q); F(h, PYCS_F_SQRTG_PC, sgc);
  dim3 gr = grid_int(g);
  int n = gr.x * gr.y * gr.z;
  TRY(ensure_partials(h, n));
  interior_reduce_kernel<<<gr, BX, 0, h->stream>>>(g, 1, Q, sgc, nullptr, h->red_part);
  CKL(h);
  final_sum_kernel<<<1, 1024, 0, h->stream>>>(h->red_part, n, h->red_out, 5);
  CKL(h);
  CK(cudaMemcpyAsync(out_host, h->red_out + 5, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}
