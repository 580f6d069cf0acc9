// Grid geometry, initial condition and diagnostics on the device ("next" rows f2 / f3 of SURVEY.md s8).
//
// Reference (file:line under /root/reference):
//   src/cs_datastruct.py:240-324   pc / pu / pv point generation            -> geometry_kernel
//   src/cs_transform.py:41-96      equiangular gnomonic map                 -> geometry_kernel
//   src/cs_transform.py:245-397    tangent vectors ex, ey                   -> geometry_kernel
//   src/cs_datastruct.py:407-446   sqrt(g) from the tangent vectors         -> geometry_kernel
//   src/cs_datastruct.py:448-493   lat-lon <-> contravariant coefficients   -> geometry_kernel
//   src/advection_ic.py:215-281    qexact_adv (initial condition / exact)   -> tracer_kernel
//   src/advection_vars.py:89-98    CFL = max over cx, cy (no abs inside)    -> rect_max_kernel
//
// The reference builds these arrays once per grid with whole-array numpy: 42 s and 14 GB at N = 1536
// (SURVEY.md s6).  Here one kernel per position writes the eight fields of a point straight into the
// device layout: same formulas in the same order, IEEE sqrt and division, no FMA contraction (this unit
// is compiled with -fmad=false).  The 1-D coordinate arrays and their tan / cos^2 come from the host
// (np.linspace and libm, as in the reference); what differs from the reference bits is the device's
// atan2 / sin / cos / x*x*x-for-pow, i.e. the last one or two ulps of the coefficients.  The host
// numpy grid (cs_datastruct.cubed_sphere) stays the bit-exact default of the parity tests; this path is
// selected with cubed_sphere(N, lean=True) and is what bench.py and large-N runs use.
#include <cmath>
#include <vector>
#include "pycs_common.cuh"

namespace {

constexpr int BX = 128;

// panel p = signed axis permutation of the panel-0 vector (src/cs_transform.py:64-92, :163-191)
__constant__ int c_rot_k[6][3] = {{0, 1, 2}, {1, 0, 2}, {0, 1, 2}, {1, 0, 2}, {2, 1, 0}, {2, 1, 0}};
__constant__ double c_rot_s[6][3] = {{1, 1, 1}, {-1, 1, 1}, {-1, -1, 1}, {1, -1, 1}, {-1, 1, 1}, {1, 1, -1}};

struct Axis { const double *x, *t, *c2; };      // coordinate, tan, cos^2 (per index)
struct GeoFields { double *sqrtg, *exlon, *exlat, *eylon, *eylat, *det, *lon, *lat; };

__global__ void geometry_kernel(Geo g, Axis ax, Axis ay, int ni, int nj, double half, GeoFields f) {
  const int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y;
  if (j >= nj || i >= ni) return;
  const double tx = ax.t[i], ty = ay.t[j], c2x = ax.c2[i], c2y = ay.c2[j];
  const double invD = 1.0 / sqrt(1.0 + tx * tx + ty * ty);
  const double v0[3] = {invD, invD * tx, invD * ty};
  const double X = half * tx, Y = half * ty;
  const double s = sqrt(half * half + X * X + Y * Y);
  const double invr3 = 1.0 / (s * s * s);                  // R = 1 (src/cs_datastruct.py:37)
  const double a2 = half * half, xy = X * Y;
  double ex0[3] = {-(half * X) * invr3, (a2 + Y * Y) * invr3, -(xy * invr3)};
  double ey0[3] = {-((half * Y) * invr3), -(xy * invr3), (a2 + X * X) * invr3};
  for (int c = 0; c < 3; ++c) {                            // chain rule of the equiangular map (:365-397)
    ex0[c] = half * ex0[c] / c2x;
    ey0[c] = half * ey0[c] / c2y;
  }
  const double dxy = ex0[0] * ey0[0] + ex0[1] * ey0[1] + ex0[2] * ey0[2];
  const double nx = ex0[0] * ex0[0] + ex0[1] * ex0[1] + ex0[2] * ex0[2];
  const double ny = ey0[0] * ey0[0] + ey0[1] * ey0[1] + ey0[2] * ey0[2];
  f.sqrtg[gidx(g, 0, i, j)] = sqrt(-(dxy * dxy) + nx * ny);
  for (int p = 0; p < 6; ++p) {
    double pos[3], ex[3], ey[3];
    for (int c = 0; c < 3; ++c) {
      const int k = c_rot_k[p][c];
      const double sg = c_rot_s[p][c];
      pos[c] = sg * v0[k];
      ex[c] = sg * ex0[k];
      ey[c] = sg * ey0[k];
    }
    const double lat = atan2(pos[2], hypot(pos[0], pos[1]));          // src/sphgeo.py:29-33
    const double lon = atan2(pos[1], pos[0]);
    const double sl = sin(lon), cl = cos(lon), st = sin(lat), ct = cos(lat);
    const double elon[3] = {-sl, cl, 0.0 * sl};
    const double elat[3] = {-st * cl, -st * sl, ct};
    const double exlon = ex[0] * elon[0] + ex[1] * elon[1] + ex[2] * elon[2];
    const double exlat = ex[0] * elat[0] + ex[1] * elat[1] + ex[2] * elat[2];
    const double eylon = ey[0] * elon[0] + ey[1] * elon[1] + ey[2] * elon[2];
    const double eylat = ey[0] * elat[0] + ey[1] * elat[1] + ey[2] * elat[2];
    const long long id = gidx(g, p, i, j);
    f.lon[id] = lon;
    f.lat[id] = lat;
    f.exlon[id] = exlon;
    f.exlat[id] = exlat;
    f.eylon[id] = eylon;
    f.eylat[id] = eylat;
    f.det[id] = exlon * eylat - eylon * exlat;
  }
}

// qexact_adv at one point (src/advection_ic.py:215-281)
__device__ double qexact_point(int ic, int vf, double t, double lon, double lat) {
  const double pi = PYCS_PI;
  if (ic == 1) return 1.0;
  const double X = cos(lat) * cos(lon), Y = cos(lat) * sin(lon), Z = sin(lat);      // src/sphgeo.py:18-22
  if (ic == 2) {
    if (vf == 1) {
      const double alpha = -45.0 * (1.0 / (180.0 / pi));     // deg2rad = 1 / rad2deg (src/constants.py)
      const double u0 = 2.0 * pi / 5.0;
      const double wt = (-u0) * t;
      const double cosa = cos(alpha), sina = sin(alpha);
      const double cos2a = cosa * cosa, sin2a = sina * sina;
      const double coswt = cos(wt), sinwt = sin(wt);
      const double rotX = (coswt * cos2a + sin2a) * X - sinwt * cosa * Y + (coswt * cosa * sina - cosa * sina) * Z;
      const double rotY = sinwt * cosa * X + coswt * Y + sina * sinwt * Z;
      const double rotZ = (coswt * sina * cosa - sina * cosa) * X - sinwt * sina * Y + (coswt * sin2a + cos2a) * Z;
      const double lon0 = pi / 4.0, lat0 = pi / 6.0;
      const double X0 = cos(lat0) * cos(lon0), Y0 = cos(lat0) * sin(lon0), Z0 = sin(lat0);
      const double dx = rotX - X0, dy = rotY - Y0, dz = rotZ - Z0;
      return exp(-10.0 * (dx * dx + dy * dy + dz * dz));
    }
    const double dx = X - 1.0, dy = Y - 0.0, dz = Z - 0.0;           // hill centred at lon = lat = 0
    return exp(-10.0 * (dx * dx + dy * dy + dz * dz));
  }
  if (ic == 3) {
    double lon1, lat1, lon2, lat2;
    if (vf == 1) { lon1 = 0; lat1 = pi / 3.0; lon2 = 0; lat2 = -pi / 3.0; }
    else { lon1 = -pi / 6.0; lat1 = 0; lon2 = pi / 6.0; lat2 = 0; }
    const double X1 = cos(lat1) * cos(lon1), Y1 = cos(lat1) * sin(lon1), Z1 = sin(lat1);
    const double X2 = cos(lat2) * cos(lon2), Y2 = cos(lat2) * sin(lon2), Z2 = sin(lat2);
    const double b0 = 5.0;
    const double d1 = (X - X1) * (X - X1) + (Y - Y1) * (Y - Y1) + (Z - Z1) * (Z - Z1);
    const double d2 = (X - X2) * (X - X2) + (Y - Y2) * (Y - Y2) + (Z - Z2) * (Z - Z2);
    return exp(-b0 * d1) + exp(-b0 * d2);
  }
  const double alpha = -45.0 * (1.0 / (180.0 / pi));
  const double fq = (-cos(lon) * cos(lat) * sin(alpha) + sin(lat) * cos(alpha));
  return 1.0 - fq * fq;
}

// interior <- qexact(t); ghost cells <- 0 (init_vars_adv fills a zero array, src/advection_vars.py:105)
__global__ void tracer_kernel(Geo g, int ic, int vf, double t, const double* __restrict__ lon,
                              const double* __restrict__ lat, double* __restrict__ q) {
  const int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y, p = blockIdx.z;
  if (j >= g.P || i >= g.P) return;
  const long long id = gidx(g, p, i, j);
  const bool inner = i >= g.lo && i < g.hi && j >= g.lo && j < g.hi;
  q[id] = inner ? qexact_point(ic, vf, t, lon[id], lat[id]) : 0.0;
}

// max over the rectangle [i0,i1) x [j0,j1) of all panels (NaN-free data): per-block partials
__global__ void rect_max_kernel(Geo g, const double* __restrict__ f, int i0, int j0, int j1, double* __restrict__ part) {
  __shared__ double sh[BX / 32];
  const int j = j0 + blockIdx.x * BX + threadIdx.x, i = i0 + blockIdx.y, p = blockIdx.z;
  double v = -INFINITY;
  if (j < j1) v = f[gidx(g, p, i, j)];
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < BX / 32; ++w) v = fmax(v, sh[w]);
    part[(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = v;
  }
}
__global__ void final_max_kernel2(const double* __restrict__ part, int n, double* __restrict__ out) {
  __shared__ double sh[32];
  double v = -INFINITY;
  for (int k = threadIdx.x; k < n; k += blockDim.x) v = fmax(v, part[k]);
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) v = fmax(v, sh[w]);
    *out = v;
  }
}

}  // namespace

// xc (P) / xe (P+1): the 1-D centre / edge coordinates of np.linspace (src/cs_datastruct.py:240-262)
int k_generate_geometry(pycs_handle h, const double* xc, const double* xe) {
  const Geo& g = h->g;
  const int P = g.P;
  // tan and cos^2 of the coordinates with the host libm, like the reference's numpy
  std::vector<double> tab((size_t)3 * (2 * P + 1));
  double* c = tab.data();                 // [x | tan | cos2] for centres, then for edges
  double* e = c + 3 * P;
  for (int k = 0; k < P; ++k) {
    c[k] = xc[k];
    c[P + k] = tan(xc[k]);
    const double cs = cos(xc[k]);
    c[2 * P + k] = cs * cs;
  }
  for (int k = 0; k <= P; ++k) {
    e[k] = xe[k];
    e[P + 1 + k] = tan(xe[k]);
    const double cs = cos(xe[k]);
    e[2 * (P + 1) + k] = cs * cs;
  }
  double* dtab = nullptr;
  CK(cudaMalloc(&dtab, tab.size() * sizeof(double)));
  CK(cudaMemcpyAsync(dtab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  const Axis ac{dtab, dtab + P, dtab + 2 * P};
  const Axis ae{dtab + 3 * P, dtab + 3 * P + (P + 1), dtab + 3 * P + 2 * (P + 1)};
  const double half = 1.0 / sqrt(3.0);                     // R / sqrt(3), R = 1
  struct Pos { int sqrtg, conv, lon, ni, nj; Axis ax, ay; };
  const Pos pos[3] = {{PYCS_F_SQRTG_PC, PYCS_F_PC_EXLON, PYCS_F_PC_LON, P, P, ac, ac},
                      {PYCS_F_SQRTG_PU, PYCS_F_PU_EXLON, PYCS_F_PU_LON, P + 1, P, ae, ac},
                      {PYCS_F_SQRTG_PV, PYCS_F_PV_EXLON, PYCS_F_PV_LON, P, P + 1, ac, ae}};
  for (const Pos& q : pos) {
    GeoFields f;
    TRY(pycs_field_ptr(h, q.sqrtg, &f.sqrtg));
    TRY(pycs_field_ptr(h, q.conv, &f.exlon));
    TRY(pycs_field_ptr(h, q.conv + 1, &f.exlat));
    TRY(pycs_field_ptr(h, q.conv + 2, &f.eylon));
    TRY(pycs_field_ptr(h, q.conv + 3, &f.eylat));
    TRY(pycs_field_ptr(h, q.conv + 4, &f.det));
    TRY(pycs_field_ptr(h, q.lon, &f.lon));
    TRY(pycs_field_ptr(h, q.lon + 1, &f.lat));
    geometry_kernel<<<dim3((q.nj + BX - 1) / BX, q.ni), BX, 0, h->stream>>>(g, q.ax, q.ay, q.ni, q.nj, half, f);
    CKL(h);
  }
  CK(cudaStreamSynchronize(h->stream));
  cudaFree(dtab);
  return 0;
}

int k_init_tracer(pycs_handle h, int field, double t) {
  const Geo& g = h->g;
  double *lon, *lat, *q;
  TRY(pycs_field_ptr(h, PYCS_F_PC_LON, &lon));
  TRY(pycs_field_ptr(h, PYCS_F_PC_LAT, &lat));
  TRY(pycs_field_ptr(h, field, &q));
  tracer_kernel<<<dim3((g.P + BX - 1) / BX, g.P, 6), BX, 0, h->stream>>>(g, h->prm.ic, h->prm.vf, t, lon, lat, q);
  CKL(h);
  return 0;
}

int k_rect_max(pycs_handle h, const double* f, int i0, int i1, int j0, int j1, double* out_host) {
  const Geo& g = h->g;
  dim3 gr((j1 - j0 + BX - 1) / BX, i1 - i0, 6);
  const int n = gr.x * gr.y * gr.z;
  if (h->red_blocks < n) {
    if (h->red_part) cudaFree(h->red_part);
    h->red_part = nullptr;
    h->red_blocks = 0;
    CK(cudaMalloc(&h->red_part, sizeof(double) * n));
    h->red_blocks = n;
  }
  rect_max_kernel<<<gr, BX, 0, h->stream>>>(g, f, i0, j0, j1, h->red_part);
  CKL(h);
  final_max_kernel2<<<1, 1024, 0, h->stream>>>(h->red_part, n, h->red_out + 6);
  CKL(h);
  CK(cudaMemcpyAsync(out_host, h->red_out + 6, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}
