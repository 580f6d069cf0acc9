// Wind path kernels.  Reference (file:line under /root/reference):
//   src/advection_ic.py:287-313      velocity_adv                     -> velocity_kernel
//   src/sphgeo.py:120-133            latlon <-> contravariant         -> ll2contra / contra2ll
//   src/advection_timestep.py:48-75  update_adv                       -> k_update_adv
//   src/averaged_velocity.py:14-62   time_averaged_velocity           -> averaged_kernel
//   src/interpolation.py:347-430     wind_edges2center_cubic_...      -> ring_kernel (+ dg fill)
//   src/interpolation.py:436-532     wind_center2ghostedge_cubic_...  -> ghost_edge_kernel
//   src/edges_treatment.py:306-347   RK2 line copies for ET-S72/PL07  -> rk2_copy_kernel
// Compiled with -fmad=false (same rounding as the numpy expressions; only the
// device sin/cos may differ from libm by an ulp).
#include <cstdlib>
#include "pycs_common.cuh"
#include "cube_edges.cuh"

namespace {

constexpr int BX = 128;
#define PI_ PYCS_PI

struct Conv { const double *exlon, *exlat, *eylon, *eylat, *det; };

__device__ __forceinline__ void ll2contra(double ulon, double vlat, const Conv& c, long long id,
                                          double* u, double* v) {
  double a = c.eylat[id] * ulon - c.eylon[id] * vlat;      // src/sphgeo.py:121-124
  double b = -c.exlat[id] * ulon + c.exlon[id] * vlat;
  *u = a / c.det[id];
  *v = b / c.det[id];
}

// velocity_adv (src/advection_ic.py:287-313) at one point
__device__ __forceinline__ void velocity_point(int vf, double t, double lo_, double la, double* uo, double* vo) {
  double u, v;
  const double pi = PI_;
  if (vf == 1) {
    double alpha = -45.0 * (1.0 / (180.0 / pi));
    double u0 = 2.0 * pi / 5.0;
    u = u0 * (cos(la) * cos(alpha) + sin(la) * cos(lo_) * sin(alpha));
    v = -u0 * sin(lo_) * sin(alpha);
  } else if (vf == 2) {
    double T = PYCS_WIND_PERIOD, k = 2.0;
    double lonp = lo_ - 2 * pi * t / T;
    double s1 = sin((lonp + pi));
    u = k * (s1 * s1) * (sin(2. * la)) * (cos(pi * t / T)) + 2. * pi * cos(la) / T;
    v = k * (sin(2 * (lonp + pi))) * (cos(la)) * (cos(pi * t / T));
  } else if (vf == 3) {
    double T = PYCS_WIND_PERIOD, k = 1.0;
    double s1 = sin((lo_ + pi) / 2.0), c1 = cos(la);
    u = -k * (s1 * s1) * (sin(2.0 * la)) * (c1 * c1) * (cos(pi * t / T));
    v = (k / 2.0) * (sin((lo_ + pi))) * (pow(c1, 3.0)) * (cos(pi * t / T));
  } else {
    double c3 = pow(cos(la), 3.0);
    u = -1 * (sin(lo_) * sin(lo_) * c3);
    v = -4 * c3 * sin(la) * cos(lo_) * sin(lo_);
  }
  *uo = u;
  *vo = v;
}

// The time-dependent wind fields are finite sums  W(lon, lat, t) = sum_m f_m(t) B_m(lon, lat)  of static
// lat-lon fields (angle addition on src/advection_ic.py:294-305, with c = cos(pi t / T), phi = 4 pi t / T):
//   field 2:  B_0 = (2 pi cos(lat) / T, 0)                                  f_0 = 1
//             B_1 = (k/2 sin(2 lat), 0)                                     f_1 = c
//             B_2 = (-k/2 sin(2 lat) cos(2 lon),  k cos(lat) sin(2 lon))    f_2 = c cos(phi)
//             B_3 = (-k/2 sin(2 lat) sin(2 lon), -k cos(lat) cos(2 lon))    f_3 = c sin(phi)
//   field 3:  B_0 = W(lon, lat, 0)                                          f_0 = c
// Everything between the lat-lon wind and the contravariant wind on the edges incl. the ghost edges is
// linear (conversion, cubic interpolation, Lagrange fill), so the contravariant wind at any time is the
// same combination of the basis fields pushed through that pipeline once (stepper.cu: basis winds).
__host__ __device__ inline int wind_basis_count(int vf) { return vf == 2 ? 4 : (vf == 3 ? 1 : 0); }
__device__ __forceinline__ void velocity_basis_point(int vf, int m, double lo_, double la, double* uo, double* vo) {
  const double pi = PI_;
  if (vf == 3) {
    velocity_point(3, 0.0, lo_, la, uo, vo);
    return;
  }
  const double T = PYCS_WIND_PERIOD, k = 2.0;
  double u = 0.0, v = 0.0;
  if (m == 0) u = 2. * pi * cos(la) / T;
  else if (m == 1) u = 0.5 * k * sin(2. * la);
  else if (m == 2) { u = -0.5 * k * sin(2. * la) * cos(2. * lo_); v = k * cos(la) * sin(2. * lo_); }
  else { u = -0.5 * k * sin(2. * la) * sin(2. * lo_); v = -k * cos(la) * cos(2. * lo_); }
  *uo = u;
  *vo = v;
}
// coefficients f_m(t) of the basis fields
__host__ __device__ inline void wind_basis_coef(int vf, double t, double* f) {
  const double pi = PI_, T = PYCS_WIND_PERIOD;
  const double c = cos(pi * t / T);
  if (vf == 3) { f[0] = c; return; }
  const double phi = 4.0 * pi * t / T;
  f[0] = 1.0; f[1] = c; f[2] = c * cos(phi); f[3] = c * sin(phi);
}

// velocity_adv on the interior edge points of one position (pu: i in [lo,hi], j in [lo,hi);
// pv: i in [lo,hi), j in [lo,hi]); basis >= 0: that basis field instead of the wind at time t
__global__ void velocity_kernel(Geo g, int vf, double t, int is_pu, const double* __restrict__ lon,
                                const double* __restrict__ lat, double* __restrict__ ulon,
                                double* __restrict__ vlat, int basis) {
  int j = g.lo + blockIdx.x * BX + threadIdx.x, i = g.lo + blockIdx.y, p = blockIdx.z;
  int ih = is_pu ? g.hi : g.hi - 1, jh = is_pu ? g.hi - 1 : g.hi;
  if (i > ih || j > jh) return;
  long long id = gidx(g, p, i, j);
  double u, v;
  if (basis >= 0) velocity_basis_point(vf, basis, lon[id], lat[id], &u, &v);
  else velocity_point(vf, t, lon[id], lat[id], &u, &v);
  ulon[id] = u;
  vlat[id] = v;
}

// Contravariant wind of a step from the basis fields (one launch per direction):
//   u  = sum_m f_m B_m           the instantaneous wind the step's upwind masks look at (src/averaged_velocity.py:21-27)
//   u* = sum_m g_m B_m           g = 1.5 f(t_{k-1}) - 0.5 f(t_{k-2}): the time-extrapolated wind (:42)
//   ubar = departure-point average of u* (:44-49, :54-62), or u itself for RK1
// on the edges lo .. hi along the direction, all transversal indices (where time_averaged_velocity writes).
struct WindBasisArgs {
  Geo g;
  const double* bu[4];
  const double* bv[4];
  double *ua, *um, *va, *vm;
  const double* coef;         // [WS_CAP][8]: f_0..f_3, g_0..g_3 of a step
  const long long* steps;     // StepCtl::steps: which row of the table
  int cmask, nb, rk2;
  double dto2;
};
// dir = 1 (v on the y-edges): the neighbours along the direction are the neighbouring threads' columns, i.e. the
// same cache lines -- plain loads.  blockIdx.z = panel.
__global__ void wind_basis_v_kernel(const __grid_constant__ WindBasisArgs a) {
  const Geo& g = a.g;
  pdl_trigger();               // PDL (pycs_common.cuh): no-ops unless launched with the attribute
  pdl_wait();
  const int p = blockIdx.z;
  const int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y;
  if (j > g.P || i >= g.P) return;
  if (j < g.lo || j > g.hi) return;
  const double* c = a.coef + 8 * (*a.steps & a.cmask);
  const long long id = gidx(g, p, i, j);
  double u = 0.0, us0 = 0.0, usm = 0.0, usp = 0.0;
  for (int m = 0; m < a.nb; ++m) {
    const double b0 = a.bv[m][id];
    u = fma(c[m], b0, u);
    if (a.rk2) {
      us0 = fma(c[4 + m], b0, us0);
      usm = fma(c[4 + m], a.bv[m][id - 1], usm);
      usp = fma(c[4 + m], a.bv[m][id + 1], usp);
    }
  }
  double r = u;
  if (a.rk2) {
    const double aa = u * a.dto2 / g.dy;
    r = (u >= 0) ? fma(aa, usm - us0, us0) : fma(aa, us0 - usp, us0);     // (1-a) u*_j + a u*_{j-1} | -a u*_{j+1} + (1+a) u*_j
    a.vm[id] = u;
  }
  a.va[id] = r;
}
// dir = 0 (u on the x-edges): the neighbours along the direction are the rows above and below; a thread owns a
// column and marches a chunk of rows with u* of three consecutive rows in registers, so that every basis
// value is loaded once.  blockIdx.y = chunk of WB_ROWS edges, blockIdx.z = panel.
constexpr int WB_ROWS = 32;
__global__ void wind_basis_u_kernel(const __grid_constant__ WindBasisArgs a) {
  const Geo& g = a.g;
  pdl_trigger();
  pdl_wait();
  const int p = blockIdx.z;
  const int j = blockIdx.x * BX + threadIdx.x;
  if (j >= g.P) return;
  const int i0 = g.lo + blockIdx.y * WB_ROWS, i1 = min(i0 + WB_ROWS, g.hi + 1);      // edges lo .. hi
  const double* c = a.coef + 8 * (*a.steps & a.cmask);
  double cf[4], cg[4];
  for (int m = 0; m < 4; ++m) { cf[m] = m < a.nb ? c[m] : 0.0; cg[m] = m < a.nb ? c[4 + m] : 0.0; }
  const long long L = g.ld;
  long long id = gidx(g, p, i0, j);
  auto combo = [&](long long at, double& f, double& gs) {
    f = 0.0; gs = 0.0;
    for (int m = 0; m < a.nb; ++m) {
      const double b = a.bu[m][at];
      f = fma(cf[m], b, f);
      gs = fma(cg[m], b, gs);
    }
  };
  double fm, usm, f0, us0, fp, usp;
  if (a.rk2) combo(id - L, fm, usm);
  combo(id, f0, us0);
  for (int i = i0; i < i1; ++i, id += L) {
    double r = f0;
    if (a.rk2) {
      combo(id + L, fp, usp);
      const double aa = f0 * a.dto2 / g.dx;
      r = (f0 >= 0) ? fma(aa, usm - us0, us0) : fma(aa, us0 - usp, us0);
      a.um[id] = f0;
      usm = us0; us0 = usp; f0 = fp;
      a.ua[id] = r;
    } else {
      a.ua[id] = r;
      if (i + 1 < i1) combo(id + L, f0, us0);
    }
  }
}
__global__ void wind_coef_fill_kernel(double* tab, int cmask, long long s0, long long k0, int n, double dt, int vf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long k = k0 + i;                         // reference step k reads the wind of t_{k-1}, old = wind of t_{k-2}
  const double tp = (double)(k - 1) * dt, tpp = (k >= 2) ? (double)(k - 2) * dt : tp;
  double f[4] = {0, 0, 0, 0}, fo[4] = {0, 0, 0, 0};
  wind_basis_coef(vf, tp, f);
  wind_basis_coef(vf, tpp, fo);
  double* row = tab + 8 * ((s0 + i) & cmask);
  for (int m = 0; m < 4; ++m) {
    row[m] = f[m];
    row[4 + m] = 1.5 * f[m] - 0.5 * fo[m];
  }
}

// update_adv for one position in ONE pass over the whole array (src/advection_timestep.py:48-75):
// old <- the normal contravariant component; wind at t on the interior edge points; lat-lon ->
// contravariant everywhere (ghost points keep their lat-lon values).  Same expressions as
// velocity_kernel + copy2_kernel + convert_kernel, so the results are bit-identical.
__global__ void update_adv_kernel(Geo g, int vf, double t, int is_pu, int ni, int nj, const double* __restrict__ lon,
                                  const double* __restrict__ lat, Conv c, double* __restrict__ ulon,
                                  double* __restrict__ vlat, double* __restrict__ uc, double* __restrict__ vc,
                                  double* __restrict__ old) {
  int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y, p = blockIdx.z;
  if (j >= nj || i >= ni) return;
  long long id = gidx(g, p, i, j);
  old[id] = is_pu ? uc[id] : vc[id];
  int ih = is_pu ? g.hi : g.hi - 1, jh = is_pu ? g.hi - 1 : g.hi;
  double ul, vl;
  if (i >= g.lo && i <= ih && j >= g.lo && j <= jh) {
    velocity_point(vf, t, lon[id], lat[id], &ul, &vl);
    ulon[id] = ul;
    vlat[id] = vl;
  } else {
    ul = ulon[id];
    vl = vlat[id];
  }
  ll2contra(ul, vl, c, id, &uc[id], &vc[id]);
}

// ll -> contravariant on a rectangle [i0,i1) x [j0,j1)
__global__ void convert_kernel(Geo g, int i0, int i1, int j0, int j1, Conv c,
                               const double* __restrict__ ulon, const double* __restrict__ vlat,
                               double* __restrict__ uc, double* __restrict__ vc) {
  int j = j0 + blockIdx.x * BX + threadIdx.x, i = i0 + blockIdx.y, p = blockIdx.z;
  if (j >= j1 || i >= i1) return;
  long long id = gidx(g, p, i, j);
  ll2contra(ulon[id], vlat[id], c, id, &uc[id], &vc[id]);
}

__global__ void copy2_kernel(Geo g, int ni, int nj, double* __restrict__ d, const double* __restrict__ s) {
  int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y, p = blockIdx.z;
  if (j >= nj || i >= ni) return;
  long long id = gidx(g, p, i, j);
  d[id] = s[id];
}

// RK2 departure velocity at edges lo..hi along DIR (src/averaged_velocity.py:36-62)
template <int DIR>
__global__ void averaged_rk2_kernel(Geo g, const double* __restrict__ u, const double* __restrict__ uold,
                                    double* __restrict__ uavg, double dto2) {
  int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y, p = blockIdx.z;
  int ni = DIR == 0 ? g.P + 1 : g.P, nj = DIR == 0 ? g.P : g.P + 1;
  if (j >= nj || i >= ni) return;
  int s = DIR == 0 ? i : j;
  if (s < g.lo || s > g.hi) return;
  const long long st = DIR == 0 ? g.ld : 1;
  long long id = gidx(g, p, i, j);
  double uu = u[id];
  double a = uu * dto2 / (DIR == 0 ? g.dx : g.dy);
  double u0 = 1.5 * uu - 0.5 * uold[id];
  double r;
  if (uu >= 0) {
    double um = 1.5 * u[id - st] - 0.5 * uold[id - st];
    r = (1.0 - a) * u0 + a * um;
  } else {
    double up = 1.5 * u[id + st] - 0.5 * uold[id + st];
    r = -a * up + (1.0 + a) * u0;
  }
  uavg[id] = r;
}

// Centre winds on the boundary ring + lat-lon conversion (src/interpolation.py:359-425)
__global__ void ring_kernel(Geo g, const double* __restrict__ u, const double* __restrict__ v,
                            double* __restrict__ uc, double* __restrict__ vc,
                            double* __restrict__ ulon, double* __restrict__ vlat,
                            const double* __restrict__ exlon, const double* __restrict__ exlat,
                            const double* __restrict__ eylon, const double* __restrict__ eylat) {
  // one block per interior row: the 4 first / last rows entirely, otherwise the 4 + 4 frame columns
  const int i = g.lo + blockIdx.y, p = blockIdx.z;
  const int lo = g.lo, hi = g.hi;
  const bool full = i < lo + PYCS_NG || i >= hi - PYCS_NG;
  const int n = full ? g.N : 2 * PYCS_NG;
  const double a1 = 5.0 / 16.0, a2 = 15.0 / 16.0, a3 = -5.0 / 16.0, a4 = 1.0 / 16.0;
  const double b1 = -1.0 / 16.0, b2 = 9.0 / 16.0, b3 = 9.0 / 16.0, b4 = -1.0 / 16.0;
  const long long L = g.ld;
  for (int t = threadIdx.x; t < n; t += BX) {
    const int j = full ? lo + t : (t < PYCS_NG ? lo + t : hi - 2 * PYCS_NG + t);
    long long id = gidx(g, p, i, j);
    double x, y;
    if (i == lo) x = a1 * u[id] + a2 * u[id + L] + a3 * u[id + 2 * L] + a4 * u[id + 3 * L];
    else if (i == hi - 1) x = a4 * u[id - 2 * L] + a3 * u[id - L] + a2 * u[id] + a1 * u[id + L];
    else x = b1 * u[id - L] + b2 * u[id] + b3 * u[id + L] + b4 * u[id + 2 * L];
    if (j == lo) y = a1 * v[id] + a2 * v[id + 1] + a3 * v[id + 2] + a4 * v[id + 3];
    else if (j == hi - 1) y = a4 * v[id - 2] + a3 * v[id - 1] + a2 * v[id] + a1 * v[id + 1];
    else y = b1 * v[id - 1] + b2 * v[id] + b3 * v[id + 1] + b4 * v[id + 2];
    uc[id] = x;
    vc[id] = y;
    ulon[id] = exlon[id] * x + eylon[id] * y;               // src/sphgeo.py:131-132
    vlat[id] = exlat[id] * x + eylat[id] * y;
  }
}

// Ghost edges from ghost centres + conversion (src/interpolation.py:443-532).
// DIR = 0: U_pu, edges along i (S/N ghost rows for i in [lo,hi], lines lo-1 and hi+1 for all j);
// DIR = 1: U_pv, edges along j.
template <int DIR>
__global__ void ghost_edge_kernel(Geo g, const double* __restrict__ culon, const double* __restrict__ cvlat,
                                  double* __restrict__ ulon, double* __restrict__ vlat,
                                  double* __restrict__ uc, double* __restrict__ vc, Conv c) {
  // one block per row i; it visits only the points this kernel writes:
  //   DIR = 0 (s = i in [lo-1, hi+1]): the lines s = lo-1, hi+1 for all j, otherwise the ghost columns;
  //   DIR = 1 (s = j in [lo-1, hi+1]): ghost rows take all s, interior rows only s = lo-1, hi+1.
  const int i = blockIdx.y, p = blockIdx.z;
  int n;
  bool wide;
  if (DIR == 0) {
    if (i < g.lo - 1 || i > g.hi + 1) return;
    wide = (i == g.lo - 1 || i == g.hi + 1);
    n = wide ? g.P : 2 * PYCS_NG;
  } else {
    if (i >= g.P) return;
    wide = (i < g.lo || i >= g.hi);
    n = wide ? g.N + 3 : 2;
  }
  const long long st = DIR == 0 ? g.ld : 1;
  const double a1 = 9.0 / 16.0, a2 = -1.0 / 16.0;
  for (int t = threadIdx.x; t < n; t += BX) {
    int j;
    if (DIR == 0) j = wide ? t : (t < PYCS_NG ? t : g.hi - PYCS_NG + t);
    else j = wide ? g.lo - 1 + t : (t == 0 ? g.lo - 1 : g.hi + 1);
    long long id = gidx(g, p, i, j);
    double ul = a1 * (culon[id] + culon[id - st]) + a2 * (culon[id + st] + culon[id - 2 * st]);
    double vl = a1 * (cvlat[id] + cvlat[id - st]) + a2 * (cvlat[id + st] + cvlat[id - 2 * st]);
    ulon[id] = ul;
    vlat[id] = vl;
    ll2contra(ul, vl, c, id, &uc[id], &vc[id]);
  }
}

// src/edges_treatment.py:311-347: ghost line of the normal wind beyond each cube edge
__global__ void rk2_copy_kernel(Geo g, double* __restrict__ u, double* __restrict__ v) {
  int t = blockIdx.x * BX + threadIdx.x;
  if (t >= g.N) return;
  CubeEdge ce = cube_edge(blockIdx.y);
  double sg = ce.a.side == ce.b.side ? -1.0 : 1.0;
  for (int d = 0; d < 2; ++d) {
    EdgeEnd dst = d == 0 ? ce.a : ce.b, src = d == 0 ? ce.b : ce.a;
    int ts = ce.flip ? g.N - 1 - t : t;
    int di = dst.side == 0 ? g.lo - 1 : g.hi + 1;
    int si = src.side == 0 ? g.lo + 1 : g.hi - 1;
    double* dp = dst.dir == 0 ? u + gidx(g, dst.panel, di, g.lo + t) : v + gidx(g, dst.panel, g.lo + t, di);
    const double* sp = src.dir == 0 ? u + gidx(g, src.panel, si, g.lo + ts) : v + gidx(g, src.panel, g.lo + ts, si);
    *dp = sg * (*sp);
  }
}

inline dim3 grid_all(int ni, int nj) { return dim3((nj + BX - 1) / BX, ni, 6); }

}  // namespace

#define F(h, id, var)    \
  double* var = nullptr; \
  TRY(pycs_field_ptr(h, id, &var))

static int get_conv(pycs_handle h, int base, Conv* c) {
  double* p[5];
  for (int k = 0; k < 5; ++k) TRY(pycs_field_ptr(h, base + k, &p[k]));
  c->exlon = p[0]; c->exlat = p[1]; c->eylon = p[2]; c->eylat = p[3]; c->det = p[4];
  return 0;
}

int k_time_averaged_velocity(pycs_handle h) {
  const Geo& g = h->g;
  F(h, PYCS_F_PU_UCONTRA, u); F(h, PYCS_F_PU_UOLD, uo); F(h, PYCS_F_PU_UAVG, ua);
  F(h, PYCS_F_PV_VCONTRA, v); F(h, PYCS_F_PV_VOLD, vo); F(h, PYCS_F_PV_VAVG, va);
  if (h->prm.dp == 1) {                        // RK1: plain copies (:30-32)
    copy2_kernel<<<grid_all(g.P + 1, g.P), BX, 0, h->stream>>>(g, g.P + 1, g.P, ua, u);
    CKL(h);
    copy2_kernel<<<grid_all(g.P, g.P + 1), BX, 0, h->stream>>>(g, g.P, g.P + 1, va, v);
    CKL(h);
  } else {
    double dto2 = g.dt * 0.5;
    averaged_rk2_kernel<0><<<grid_all(g.P + 1, g.P), BX, 0, h->stream>>>(g, u, uo, ua, dto2);
    CKL(h);
    averaged_rk2_kernel<1><<<grid_all(g.P, g.P + 1), BX, 0, h->stream>>>(g, v, vo, va, dto2);
    CKL(h);
  }
  return 0;
}

// wind_edges2center_cubic_interpolation (src/interpolation.py:347-430): C-grid winds -> centres on the boundary
// ring, -> lat-lon, Lagrange ghost fill of both components
int k_wind_edges2center(pycs_handle h) {
  const Geo& g = h->g;
  F(h, PYCS_F_PU_UCONTRA, u); F(h, PYCS_F_PV_VCONTRA, v);
  F(h, PYCS_F_PC_UCONTRA, cu); F(h, PYCS_F_PC_VCONTRA, cv);
  F(h, PYCS_F_PC_ULON, cul); F(h, PYCS_F_PC_VLAT, cvl);
  F(h, PYCS_F_PC_EXLON, exlon); F(h, PYCS_F_PC_EXLAT, exlat);
  F(h, PYCS_F_PC_EYLON, eylon); F(h, PYCS_F_PC_EYLAT, eylat);
  ring_kernel<<<dim3(1, g.N, 6), BX, 0, h->stream>>>(g, u, v, cu, cv, cul, cvl, exlon, exlat, eylon, eylat);
  CKL(h);
  TRY(k_dg_fill_single(h, cul));
  return k_dg_fill_single(h, cvl);
}

// wind_center2ghostedge_cubic_interpolation (src/interpolation.py:436-532): ghost centres -> ghost edges, -> contravariant
int k_wind_center2ghostedge(pycs_handle h) {
  const Geo& g = h->g;
  F(h, PYCS_F_PU_UCONTRA, u); F(h, PYCS_F_PV_VCONTRA, v);
  F(h, PYCS_F_PC_ULON, cul); F(h, PYCS_F_PC_VLAT, cvl);
  F(h, PYCS_F_PU_ULON, uul); F(h, PYCS_F_PU_VLAT, uvl); F(h, PYCS_F_PU_VCONTRA, uvc);
  F(h, PYCS_F_PV_ULON, vul); F(h, PYCS_F_PV_VLAT, vvl); F(h, PYCS_F_PV_UCONTRA, vuc);
  Conv cpu, cpv;
  TRY(get_conv(h, PYCS_F_PU_EXLON, &cpu));
  TRY(get_conv(h, PYCS_F_PV_EXLON, &cpv));
  ghost_edge_kernel<0><<<dim3(1, g.P + 1, 6), BX, 0, h->stream>>>(g, cul, cvl, uul, uvl, u, uvc, cpu);
  CKL(h);
  ghost_edge_kernel<1><<<dim3(1, g.P, 6), BX, 0, h->stream>>>(g, cul, cvl, vul, vvl, vuc, v, cpv);
  CKL(h);
  return 0;
}

// edges_ghost_cell_treatment_vector (src/edges_treatment.py:296-347)
int k_wind_ghost_fill(pycs_handle h) {
  const Geo& g = h->g;
  if (h->prm.et == 3) {
    TRY(k_wind_edges2center(h));
    return k_wind_center2ghostedge(h);
  }
  if (h->prm.dp == 2) {
    F(h, PYCS_F_PU_UCONTRA, u); F(h, PYCS_F_PV_VCONTRA, v);
    rk2_copy_kernel<<<dim3((g.N + BX - 1) / BX, 12), BX, 0, h->stream>>>(g, u, v);
    CKL(h);
  }
  return 0;
}

// velocity at time t on the interior edge points; convert_interior_only = 1 for the
// init block (src/advection_vars.py:37-53), 0 for update_adv (whole arrays, :65-75)
int k_wind_interior(pycs_handle h, double t, int convert_interior_only, int do_velocity, int basis) {
  const Geo& g = h->g;
  double *ulon_ = nullptr, *ulat_ = nullptr, *vlon_ = nullptr, *vlat_ = nullptr;
  if (do_velocity) {
    TRY(pycs_field_ptr(h, PYCS_F_PU_LON, &ulon_)); TRY(pycs_field_ptr(h, PYCS_F_PU_LAT, &ulat_));
    TRY(pycs_field_ptr(h, PYCS_F_PV_LON, &vlon_)); TRY(pycs_field_ptr(h, PYCS_F_PV_LAT, &vlat_));
  }
  F(h, PYCS_F_PU_ULON, uul); F(h, PYCS_F_PU_VLAT, uvl); F(h, PYCS_F_PU_UCONTRA, uuc); F(h, PYCS_F_PU_VCONTRA, uvc);
  F(h, PYCS_F_PV_ULON, vul); F(h, PYCS_F_PV_VLAT, vvl); F(h, PYCS_F_PV_UCONTRA, vuc); F(h, PYCS_F_PV_VCONTRA, vvc);
  dim3 gr((g.N + 1 + BX - 1) / BX, g.N + 1, 6);
  if (do_velocity) {
    velocity_kernel<<<gr, BX, 0, h->stream>>>(g, h->prm.vf, t, 1, ulon_, ulat_, uul, uvl, basis);
    CKL(h);
    velocity_kernel<<<gr, BX, 0, h->stream>>>(g, h->prm.vf, t, 0, vlon_, vlat_, vul, vvl, basis);
    CKL(h);
  }
  Conv cpu, cpv;
  TRY(get_conv(h, PYCS_F_PU_EXLON, &cpu));
  TRY(get_conv(h, PYCS_F_PV_EXLON, &cpv));
  if (convert_interior_only) {
    convert_kernel<<<gr, BX, 0, h->stream>>>(g, g.lo, g.hi + 1, g.lo, g.hi, cpu, uul, uvl, uuc, uvc);
    CKL(h);
    convert_kernel<<<gr, BX, 0, h->stream>>>(g, g.lo, g.hi, g.lo, g.hi + 1, cpv, vul, vvl, vuc, vvc);
    CKL(h);
  } else {
    F(h, PYCS_F_PU_UOLD, uo); F(h, PYCS_F_PV_VOLD, vo);
    copy2_kernel<<<grid_all(g.P + 1, g.P), BX, 0, h->stream>>>(g, g.P + 1, g.P, uo, uuc);   // :61-62
    CKL(h);
    copy2_kernel<<<grid_all(g.P, g.P + 1), BX, 0, h->stream>>>(g, g.P, g.P + 1, vo, vvc);
    CKL(h);
    convert_kernel<<<grid_all(g.P + 1, g.P), BX, 0, h->stream>>>(g, 0, g.P + 1, 0, g.P, cpu, uul, uvl, uuc, uvc);
    CKL(h);
    convert_kernel<<<grid_all(g.P, g.P + 1), BX, 0, h->stream>>>(g, 0, g.P, 0, g.P + 1, cpv, vul, vvl, vuc, vvc);
    CKL(h);
  }
  return 0;
}

int k_update_adv(pycs_handle h, double t) {
  if (h->prm.vf < 2) return 0;               // src/advection_timestep.py:50
  if (getenv("PYCS_UNFUSED_WIND")) return k_wind_interior(h, t, 0, 1);
  // one pass per position: wind at t, old <- normal component, conversion of the whole array
  const Geo& g = h->g;
  F(h, PYCS_F_PU_LON, ulon_); F(h, PYCS_F_PU_LAT, ulat_); F(h, PYCS_F_PV_LON, vlon_); F(h, PYCS_F_PV_LAT, vlat_);
  F(h, PYCS_F_PU_ULON, uul); F(h, PYCS_F_PU_VLAT, uvl); F(h, PYCS_F_PU_UCONTRA, uuc); F(h, PYCS_F_PU_VCONTRA, uvc);
  F(h, PYCS_F_PV_ULON, vul); F(h, PYCS_F_PV_VLAT, vvl); F(h, PYCS_F_PV_UCONTRA, vuc); F(h, PYCS_F_PV_VCONTRA, vvc);
  F(h, PYCS_F_PU_UOLD, uo); F(h, PYCS_F_PV_VOLD, vo);
  Conv cpu, cpv;
  TRY(get_conv(h, PYCS_F_PU_EXLON, &cpu));
  TRY(get_conv(h, PYCS_F_PV_EXLON, &cpv));
  update_adv_kernel<<<grid_all(g.P + 1, g.P), BX, 0, h->stream>>>(g, h->prm.vf, t, 1, g.P + 1, g.P, ulon_, ulat_, cpu,
                                                                  uul, uvl, uuc, uvc, uo);
  CKL(h);
  update_adv_kernel<<<grid_all(g.P, g.P + 1), BX, 0, h->stream>>>(g, h->prm.vf, t, 0, g.P, g.P + 1, vlon_, vlat_, cpv,
                                                                  vul, vvl, vuc, vvc, vo);
  CKL(h);
  return 0;
}


// ---- basis winds (stepper.cu) ------------------------------------------------------------------------
int k_wind_basis_count(pycs_handle h) { return wind_basis_count(h->prm.vf); }

// basis field m through the wind pipeline of init_vars_adv (interior conversion + ghost fill,
// src/advection_vars.py:37-80): leaves it in U_pu.ucontra / U_pv.vcontra (and clobbers U_pu / U_pv / U_pc)
int k_wind_basis_build(pycs_handle h, int m) {
  TRY(k_wind_interior(h, 0.0, 1, 1, m));
  return k_wind_ghost_fill(h);
}

int k_wind_basis_combine(pycs_handle h, double* const* bu, double* const* bv, int nb, double* ua, double* um, double* va,
                         double* vm, const double* coef, int cmask, const long long* steps, int pdl) {
  const Geo& g = h->g;
  WindBasisArgs a;
  a.g = g;
  for (int m = 0; m < 4; ++m) {
    a.bu[m] = m < nb ? bu[m] : nullptr;
    a.bv[m] = m < nb ? bv[m] : nullptr;
  }
  a.ua = ua; a.um = um; a.va = va; a.vm = vm;
  a.coef = coef;
  a.steps = steps;
  a.cmask = cmask;
  a.nb = nb;
  a.rk2 = (h->prm.dp == 2) ? 1 : 0;
  a.dto2 = g.dt * 0.5;
  const dim3 gu((g.P + BX - 1) / BX, (g.N + 1 + WB_ROWS - 1) / WB_ROWS, 6), gv((g.P + 1 + BX - 1) / BX, g.P, 6);
  if (pdl) CK(pycs_launch_pdl(wind_basis_u_kernel, gu, dim3(BX), 0, h->stream, true, a));
  else wind_basis_u_kernel<<<gu, BX, 0, h->stream>>>(a);
  CKL(h);
  if (pdl) CK(pycs_launch_pdl(wind_basis_v_kernel, gv, dim3(BX), 0, h->stream, true, a));
  else wind_basis_v_kernel<<<gv, BX, 0, h->stream>>>(a);
  CKL(h);
  return 0;
}

int k_wind_coef_fill(pycs_handle h, double* tab, int cmask, long long s0, long long k0, int n) {
  wind_coef_fill_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(tab, cmask, s0, k0, n, h->g.dt, h->prm.vf);
  CKL(h);
  return 0;
}
