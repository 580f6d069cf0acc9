// C ABI of libpycs_b200.so (declared in include/pycs_b200.h): handle lifetime,
// reference-layout <-> device-layout transfers and the operator / step drivers.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>
#include <nvtx3/nvToolsExt.h>
#include "pycs_common.cuh"
#include "fused_args.cuh"
#include "mgpu.cuh"

// NVTX range per reference-level operation (header-only NVTX 3: a no-op unless a tool is attached), so that
// a timeline shows adv_time_step / divergence / ghost fills the way the reference's call stack names them
namespace {
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
}  // namespace
#define PYCS_RANGE(name) NvtxRange nvtx_range__(name)

static thread_local std::string g_err;
void pycs_set_error(const std::string& msg) { g_err = msg; }
int pycs_cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof buf, "CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
  g_err = buf;
  return PYCS_ERR_CUDA;
}
static int arg_fail(const char* msg) {
  g_err = msg;
  return PYCS_ERR_ARG;
}

int pycs_field_shape(const Geo& g, int f, int* ni, int* nj, int* np) {
  if (f < 0 || f >= PYCS_F_COUNT) return arg_fail("unknown field id");
  *ni = pycs_field_is_u(f) ? g.P + 1 : g.P;
  *nj = pycs_field_is_v(f) ? g.P + 1 : g.P;
  *np = pycs_field_single_panel(f) ? 1 : 6;
  return 0;
}

int pycs_field_ptr(pycs_handle h, int f, double** out) {
  if (f < 0 || f >= PYCS_F_COUNT) return arg_fail("unknown field id");
  if (!h->f[f]) {
    size_t n = (size_t)(pycs_field_single_panel(f) ? 1 : 6) * h->g.ps;
    CK(cudaMalloc(&h->f[f], n * sizeof(double)));
    CK(cudaMemsetAsync(h->f[f], 0, n * sizeof(double), h->stream));
  }
  *out = h->f[f];
  return 0;
}

// --------------------------------------------------------------------------- layout kernels
namespace {
// host reference layout [i][j][6] (staged on the device) -> panel-major padded
__global__ void to_device_layout(Geo g, int i0, int nj, int np, const double* __restrict__ src,
                                 double* __restrict__ dst) {
  int j = blockIdx.x * blockDim.x + threadIdx.x, i = i0 + blockIdx.y;
  if (j >= nj) return;
  const double* s = src + ((long long)i * nj + j) * 6;
  for (int p = 0; p < np; ++p) dst[gidx(g, p, i, j)] = s[p];
}
__global__ void from_device_layout(Geo g, int i0, int nj, int np, const double* __restrict__ src,
                                   double* __restrict__ dst) {
  int j = blockIdx.x * blockDim.x + threadIdx.x, i = i0 + blockIdx.y;
  if (j >= nj) return;
  double* d = dst + ((long long)i * nj + j) * 6;
  for (int p = 0; p < 6; ++p) d[p] = src[gidx(g, np == 1 ? 0 : p, i, j)];
}
__global__ void fill_kernel(double* p, long long n, double v) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) p[k] = v;
}
}  // namespace

static int ensure_stage(pycs_handle h, size_t bytes) {
  if (h->stage_bytes < bytes) {
    if (h->stage_dev) cudaFree(h->stage_dev);
    h->stage_dev = nullptr;
    h->stage_bytes = 0;
    CK(cudaMalloc(&h->stage_dev, bytes));
    h->stage_bytes = bytes;
  }
  return 0;
}
static int ensure_pinned(pycs_handle h, size_t bytes) {
  if (h->pin_bytes < bytes) {
    if (h->stage_pin) cudaFreeHost(h->stage_pin);
    h->stage_pin = nullptr;
    h->pin_bytes = 0;
    CK(cudaMallocHost(&h->stage_pin, bytes));
    h->pin_bytes = bytes;
  }
  return 0;
}

// pycs_adv_time_step_host advances Q with the separable wind and leaves the exposed wind arrays
// behind (wind_stale_k); every other entry point that reads or writes wind state, or changes
// what it depends on, calls this first.
static int wind_sync(pycs_handle h) {
  if (h->wind_stale_k < 0) return 0;
  const long long k = h->wind_stale_k;
  h->wind_stale_k = -1;
  return k_wind_catch_up(h, k);
}

// The fused step leaves Q in whichever of the two ping-pong buffers it wrote last, with the MF-PR
// projection term of the last step still pending and a stale ghost ring (fused.cu).  pycs_run does
// not pay for that at the end of every call (the reference's loop has no per-run epilogue either,
// src/advection_sphere.py:45-57): every other entry point that reads or writes Q calls this first.
static int normalize_q(pycs_handle h) {
  TRY(k_fused_flush(h));
  if (h->qcur == 1) {
    double* t = h->f[PYCS_F_Q];
    h->f[PYCS_F_Q] = h->f[PYCS_F_Q_NEXT];
    h->f[PYCS_F_Q_NEXT] = t;
    h->qcur = 0;
  }
  return 0;
}
static int state_sync(pycs_handle h) {
  TRY(wind_sync(h));
  return normalize_q(h);
}
// the same, but only for what an access to `field` can observe: Q / Q_NEXT need the projection flush, the wind
// arrays (and cx / cy, which are formed from them) the lazy wind catch-up; everything else neither
static bool is_wind_field(int f) { return (f >= PYCS_F_PU_ULON && f <= PYCS_F_PC_VCONTRA) || f == PYCS_F_CX || f == PYCS_F_CY; }
static int field_sync(pycs_handle h, int field) {
  if (field == PYCS_F_Q || field == PYCS_F_Q_NEXT) return normalize_q(h);
  if (is_wind_field(field)) return wind_sync(h);
  return 0;
}

// --------------------------------------------------------------------------- lifetime
extern "C" const char* pycs_last_error(void) { return g_err.c_str(); }

extern "C" int pycs_create(const pycs_params* prm, pycs_handle* out) {
  if (!prm || !out) return arg_fail("null argument");
  if (prm->N < 8) return arg_fail("N must be >= 8");
  if (prm->recon < 1 || prm->recon > 4 || prm->dp < 1 || prm->dp > 2 || prm->opsplit < 1 || prm->opsplit > 3 ||
      prm->et < 1 || prm->et > 3 || prm->mt < 1 || prm->mt > 2 || prm->mf < 1 || prm->mf > 3 || prm->vf < 1 ||
      prm->vf > 4)
    return arg_fail("scheme selector out of range (src/advection_ic.py:85-149)");
  // the only combinations the reference can run (src/discrete_operators.py:72-81)
  bool ok = (prm->opsplit <= 2 && prm->mt == 1) || (prm->opsplit == 3 && prm->mt == 2);
  if (!ok) return arg_fail("invalid opsplit/mt combination: SP-AVLT|SP-L04 need MT-0, SP-PL07 needs MT-PL07");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_err = "no CUDA device: libpycs_b200 has no CPU fallback";
    return PYCS_ERR_CUDA;
  }
  if (prm->device < 0 || prm->device >= ndev) return arg_fail("bad device ordinal");
  CK(cudaSetDevice(prm->device));
  pycs_handle h = new (std::nothrow) pycs_handle_s();
  if (!h) return PYCS_ERR_NOMEM;
  memset(h, 0, sizeof *h);
  h->prm = *prm;
  h->device = prm->device;
  Geo& g = h->g;
  g.N = prm->N;
  g.P = prm->N + 2 * PYCS_NG;
  g.lo = PYCS_NG;
  g.hi = PYCS_NG + prm->N;
  g.ld = ((PYCS_JOFF + g.P + 1 + 15) / 16) * 16;
  g.ps = (long long)(g.P + 1) * g.ld;
  g.dx = prm->dx;
  g.dy = prm->dy;
  g.dt = prm->dt;
  h->row_lo = g.lo;
  h->row_hi = g.hi;
  h->wind_stale_k = -1;
  h->no_separable = getenv("PYCS_NO_SEPARABLE") ? 1 : 0;     // environment knobs are read once per handle
  h->dg_two_phase = getenv("PYCS_DG_TWO_PHASE") ? 1 : 0;
  cudaDeviceProp dp;
  CK(cudaGetDeviceProperties(&dp, prm->device));
  h->sm_count = dp.multiProcessorCount;
  {   // highest priority: in a split step the boundary CTAs run on this stream, ahead of the interior CTAs
    int lo_pri = 0, hi_pri = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
    CK(cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, hi_pri));
  }
  CK(cudaEventCreate(&h->ev0));
  CK(cudaEventCreate(&h->ev1));
  CK(cudaMalloc(&h->red_out, 16 * sizeof(double)));
  CK(cudaMemset(h->red_out, 0, 16 * sizeof(double)));
  CK(cudaMalloc(&h->halo_buf, sizeof(double) * 4 * 6 * 4 * g.P));
  pycs_build_halo_maps(g, &h->maps);
  *out = h;
  return 0;
}

extern "C" int pycs_destroy(pycs_handle h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  k_fused_release(h);
  k_mg_release(h);
  for (int k = 0; k < PYCS_F_COUNT; ++k)
    if (h->f[k]) cudaFree(h->f[k]);
  if (h->kminE) cudaFree(h->kminE);
  if (h->wE) cudaFree(h->wE);
  free(h->kmin_host);
  if (h->halo_buf) cudaFree(h->halo_buf);
  if (h->red_part) cudaFree(h->red_part);
  if (h->red_out) cudaFree(h->red_out);
  if (h->stage_dev) cudaFree(h->stage_dev);
  if (h->stage_pin) cudaFreeHost(h->stage_pin);
  cudaEventDestroy(h->ev0);
  cudaEventDestroy(h->ev1);
  cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

extern "C" int pycs_device_info(pycs_handle h, int32_t* sm_count, char* name, int32_t name_len) {
  cudaDeviceProp dp;
  CK(cudaGetDeviceProperties(&dp, h->device));
  if (sm_count) *sm_count = dp.multiProcessorCount;
  if (name && name_len > 0) {
    strncpy(name, dp.name, name_len - 1);
    name[name_len - 1] = 0;
  }
  return 0;
}

extern "C" int pycs_synchronize(pycs_handle h) {
  CK(cudaStreamSynchronize(h->stream));
  return k_mg_check(h);
}

extern "C" int pycs_set_dt(pycs_handle h, double dt) {
  TRY(wind_sync(h));
  h->g.dt = dt;
  h->prm.dt = dt;
  k_fused_invalidate_graphs(h);
  return 0;
}

extern "C" int pycs_launch_count(pycs_handle h, int64_t* count) {
  *count = h->launches;
  return 0;
}

// --------------------------------------------------------------------------- transfers
// rows [i0, i1) of a field (the reference layout keeps a row of all six panels contiguous); i1 < 0: all rows
static int upload_from(pycs_handle h, int field, const double* host, int i0 = 0, int i1 = -1) {
  int ni, nj, np;
  TRY(pycs_field_shape(h->g, field, &ni, &nj, &np));
  if (i1 < 0) i1 = ni;
  double* dst;
  TRY(pycs_field_ptr(h, field, &dst));
  const size_t row = (size_t)nj * 6, bytes = (size_t)(i1 - i0) * row * sizeof(double);
  TRY(ensure_stage(h, (size_t)ni * row * sizeof(double)));
  CK(cudaMemcpyAsync(h->stage_dev + i0 * row, host + i0 * row, bytes, cudaMemcpyHostToDevice, h->stream));
  dim3 grid((nj + 127) / 128, i1 - i0);
  to_device_layout<<<grid, 128, 0, h->stream>>>(h->g, i0, nj, np, h->stage_dev, dst);
  CKL(h);
  return 0;
}

extern "C" int pycs_upload_field(pycs_handle h, int32_t field, const double* host) {
  PYCS_RANGE("upload_field");
  // wind arrays, and the geometry the wind catch-up is computed from, must not be overwritten under a pending catch-up
  if (is_wind_field(field) || (field >= PYCS_F_SQRTG_PC && field <= PYCS_F_PV_LAT)) TRY(wind_sync(h));
  if (!host) return arg_fail("null host pointer");
  CK(cudaSetDevice(h->device));
  if (field == PYCS_F_Q) TRY(k_fused_discard(h));   // a new state: nothing of the old one is pending
  else if (field == PYCS_F_Q_NEXT) TRY(normalize_q(h));
  TRY(upload_from(h, field, host));
  CK(cudaStreamSynchronize(h->stream));
  if (field == PYCS_F_SQRTG_PC) h->a2_valid = 0;
  if (field >= PYCS_F_SQRTG_PC && field <= PYCS_F_PV_LAT) k_fused_invalidate(h);
  return 0;
}

static int download_to(pycs_handle h, int field, double* host, int i0 = 0, int i1 = -1) {
  int ni, nj, np;
  TRY(pycs_field_shape(h->g, field, &ni, &nj, &np));
  if (i1 < 0) i1 = ni;
  double* src;
  TRY(pycs_field_ptr(h, field, &src));
  const size_t row = (size_t)nj * 6, bytes = (size_t)(i1 - i0) * row * sizeof(double);
  TRY(ensure_stage(h, (size_t)ni * row * sizeof(double)));
  dim3 grid((nj + 127) / 128, i1 - i0);
  from_device_layout<<<grid, 128, 0, h->stream>>>(h->g, i0, nj, np, src, h->stage_dev);
  CKL(h);
  CK(cudaMemcpyAsync(host + i0 * row, h->stage_dev + i0 * row, bytes, cudaMemcpyDeviceToHost, h->stream));
  return 0;
}

extern "C" int pycs_download_field(pycs_handle h, int32_t field, double* host) {
  PYCS_RANGE("download_field");
  TRY(field_sync(h, field));
  if (!host) return arg_fail("null host pointer");
  CK(cudaSetDevice(h->device));
  TRY(download_to(h, field, host));
  CK(cudaStreamSynchronize(h->stream));
  return k_mg_check(h);
}

extern "C" int pycs_copy_field(pycs_handle h, int32_t dst, int32_t src) {
  TRY(field_sync(h, dst));
  TRY(field_sync(h, src));
  double *d, *s;
  TRY(pycs_field_ptr(h, dst, &d));
  TRY(pycs_field_ptr(h, src, &s));
  if (pycs_field_single_panel(dst) != pycs_field_single_panel(src)) return arg_fail("shape mismatch");
  size_t n = (size_t)(pycs_field_single_panel(dst) ? 1 : 6) * h->g.ps;
  CK(cudaMemcpyAsync(d, s, n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  return 0;
}

extern "C" int pycs_fill_field(pycs_handle h, int32_t field, double value) {
  TRY(field_sync(h, field));
  double* d;
  TRY(pycs_field_ptr(h, field, &d));
  long long n = (long long)(pycs_field_single_panel(field) ? 1 : 6) * h->g.ps;
  fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(d, n, value);
  CKL(h);
  return 0;
}

extern "C" int pycs_upload_lagrange(pycs_handle h, int32_t degree, const int32_t* kmin_east,
                                    const double* weights_east) {
  if (degree < 0 || degree > 7 || !kmin_east || !weights_east) return arg_fail("bad Lagrange table");
  CK(cudaSetDevice(h->device));
  TRY(normalize_q(h));                   // a pending projection term was defined with the old tables' ghost(sqrtg)
  int order = degree + 1, P = h->g.P;
  // validate the stencil range here: the kernels index neighbour strips with it
  for (int k = 0; k < 4 * P; ++k)
    if (kmin_east[k] < 0 || kmin_east[k] + order > P) return arg_fail("Kmin outside [0, P-order]");
  if (h->kminE) cudaFree(h->kminE);
  if (h->wE) cudaFree(h->wE);
  h->kminE = nullptr;
  h->wE = nullptr;
  CK(cudaMalloc(&h->kminE, sizeof(int) * 4 * P));
  CK(cudaMalloc(&h->wE, sizeof(double) * 4 * P * order));
  CK(cudaMemcpy(h->kminE, kmin_east, sizeof(int) * 4 * P, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->wE, weights_east, sizeof(double) * 4 * P * order, cudaMemcpyHostToDevice));
  free(h->kmin_host);
  h->kmin_host = (int*)malloc(sizeof(int) * 4 * P);
  if (!h->kmin_host) return PYCS_ERR_NOMEM;
  memcpy(h->kmin_host, kmin_east, sizeof(int) * 4 * P);
  h->degree = degree;
  h->order = order;
  k_fused_invalidate_ghost_metric(h);    // ghost(sqrtg) was filled with the old tables
  return k_mg_replan(h);                 // the exchange plan follows the stencils
}

// --------------------------------------------------------------------------- device-side set-up (f2 / f3)
extern "C" int pycs_generate_geometry(pycs_handle h, const double* xc, const double* xe) {
  PYCS_RANGE("generate_geometry");
  if (!xc || !xe) return arg_fail("null coordinate array");
  CK(cudaSetDevice(h->device));
  TRY(state_sync(h));
  TRY(k_generate_geometry(h, xc, xe));
  h->a2_valid = 0;
  k_fused_invalidate(h);
  return 0;
}

extern "C" int pycs_init_tracer(pycs_handle h, int32_t field, double t) {
  if (field < 0 || field >= PYCS_F_COUNT || pycs_field_is_u(field) || pycs_field_is_v(field) ||
      pycs_field_single_panel(field))
    return arg_fail("pycs_init_tracer: not a six-panel centre field");
  CK(cudaSetDevice(h->device));
  if (field == PYCS_F_Q) TRY(k_fused_discard(h));
  else TRY(normalize_q(h));
  return k_init_tracer(h, field, t);
}

extern "C" int pycs_field_max(pycs_handle h, int32_t field, int32_t i0, int32_t i1, int32_t j0, int32_t j1, double* out) {
  int ni, nj, np;
  TRY(pycs_field_shape(h->g, field, &ni, &nj, &np));
  if (np != 6 || i0 < 0 || i1 > ni || j0 < 0 || j1 > nj || i0 >= i1 || j0 >= j1 || !out) return arg_fail("pycs_field_max: bad rectangle");
  CK(cudaSetDevice(h->device));
  TRY(state_sync(h));
  double* f;
  TRY(pycs_field_ptr(h, field, &f));
  return k_rect_max(h, f, i0, i1, j0, j1, out);
}

// --------------------------------------------------------------------------- halo API
extern "C" int pycs_halo_gather(pycs_handle h, int32_t fx, int32_t fy, double* east, double* west,
                                double* north, double* south) {
  TRY(normalize_q(h));
  double *x, *y;
  TRY(pycs_field_ptr(h, fx, &x));
  TRY(pycs_field_ptr(h, fy, &y));
  TRY(k_halo_gather(h, x, y, h->halo_buf));
  int P = h->g.P, n = 4 * P;
  std::vector<double> tmp((size_t)4 * 6 * n);
  CK(cudaMemcpyAsync(tmp.data(), h->halo_buf, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  double* outs[4] = {east, west, north, south};
  for (int s = 0; s < 4; ++s)
    for (int p = 0; p < 6; ++p)
      for (int t = 0; t < n; ++t) outs[s][(size_t)t * 6 + p] = tmp[((size_t)s * 6 + p) * n + t];
  return 0;
}

extern "C" int pycs_halo_fill_dg(pycs_handle h, int32_t field) {
  PYCS_RANGE("ghost_cell_pc_lagrange_interpolation");
  TRY(normalize_q(h));
  double* q;
  TRY(pycs_field_ptr(h, field, &q));
  if (h->dg_two_phase) return k_dg_fill(h, q);     // the reference's two phases as two launches
  return k_dg_fill_single(h, q);
}

extern "C" int pycs_halo_fill_copy(pycs_handle h, int32_t fx, int32_t fy) {
  PYCS_RANGE("ghost_cells_adjacent_panels");
  TRY(normalize_q(h));
  double *x, *y;
  TRY(pycs_field_ptr(h, fx, &x));
  TRY(pycs_field_ptr(h, fy, &y));
  TRY(k_halo_gather(h, x, y, h->halo_buf));
  return k_halo_scatter_copy(h, x, y, h->halo_buf);
}

extern "C" int pycs_halo_fill_scalar(pycs_handle h, int32_t fx, int32_t fy) {
  if (h->prm.et == 3) return pycs_halo_fill_dg(h, fx);      // src/edges_treatment.py:289-290
  return pycs_halo_fill_copy(h, fx, fy);
}

extern "C" int pycs_halo_fill_vector(pycs_handle h) {
  PYCS_RANGE("edges_ghost_cell_treatment_vector");
  TRY(wind_sync(h));
  return k_wind_ghost_fill(h);
}

extern "C" int pycs_wind_edges2center(pycs_handle h) {
  TRY(wind_sync(h));
  if (!h->kminE) {
    g_err = "wind_edges2center needs pycs_upload_lagrange first";
    return PYCS_ERR_STATE;
  }
  return k_wind_edges2center(h);
}
extern "C" int pycs_wind_center2ghostedge(pycs_handle h) {
  TRY(wind_sync(h));
  return k_wind_center2ghostedge(h);
}

// --------------------------------------------------------------------------- operators
extern "C" int pycs_time_averaged_velocity(pycs_handle h) {
  PYCS_RANGE("time_averaged_velocity");
  TRY(wind_sync(h));
  return k_time_averaged_velocity(h);
}

extern "C" int pycs_cfl(pycs_handle h, int32_t dst, int32_t src, int32_t dir) {
  TRY(wind_sync(h));
  double *d, *s;
  TRY(pycs_field_ptr(h, dst, &d));
  TRY(pycs_field_ptr(h, src, &s));
  return k_cfl(h, d, s, dir);
}

extern "C" int pycs_ppm_reconstruction(pycs_handle h, int32_t fx, int32_t fy) {
  TRY(normalize_q(h));
  double *x, *y;
  TRY(pycs_field_ptr(h, fx, &x));
  TRY(pycs_field_ptr(h, fy, &y));
  TRY(k_recon(h, x, y));
  if (h->prm.et == 2) TRY(k_edges_extrapolation(h, x, y));   // src/reconstruction_1d.py:392-394
  return 0;
}

extern "C" int pycs_edges_extrapolation(pycs_handle h, int32_t fx, int32_t fy) {
  TRY(normalize_q(h));
  double *x, *y;
  TRY(pycs_field_ptr(h, fx, &x));
  TRY(pycs_field_ptr(h, fy, &y));
  return k_edges_extrapolation(h, x, y);
}

extern "C" int pycs_numerical_flux(pycs_handle h, int32_t fx, int32_t fy) {
  TRY(state_sync(h));
  double *x, *y;
  TRY(pycs_field_ptr(h, fx, &x));
  TRY(pycs_field_ptr(h, fy, &y));
  return k_flux(h, x, y);
}

extern "C" int pycs_compute_fluxes(pycs_handle h, int32_t fx, int32_t fy) {
  PYCS_RANGE("compute_fluxes");
  TRY(wind_sync(h));
  TRY(pycs_ppm_reconstruction(h, fx, fy));
  return pycs_numerical_flux(h, fx, fy);
}

extern "C" int pycs_F_operator(pycs_handle h) { return k_flux_diff(h, 0); }
extern "C" int pycs_G_operator(pycs_handle h) { return k_flux_diff(h, 1); }
extern "C" int pycs_average_flux_cube_edges(pycs_handle h) { return k_average_flux_edges(h); }

// divergence, operator by operator (src/discrete_operators.py:18-101)
extern "C" int pycs_divergence(pycs_handle h) {
  PYCS_RANGE("divergence");
  TRY(state_sync(h));
  double *Q, *gQ, *cx, *cy, *ua, *va;
  TRY(pycs_field_ptr(h, PYCS_F_Q, &Q));
  TRY(pycs_field_ptr(h, PYCS_F_GQ, &gQ));
  TRY(pycs_field_ptr(h, PYCS_F_CX, &cx));
  TRY(pycs_field_ptr(h, PYCS_F_CY, &cy));
  TRY(pycs_field_ptr(h, PYCS_F_PU_UAVG, &ua));
  TRY(pycs_field_ptr(h, PYCS_F_PV_VAVG, &va));
  TRY(k_mul_metric(h, gQ, Q));                                  // :31
  TRY(k_cfl(h, cx, ua, 0));                                     // :34-35
  TRY(k_cfl(h, cy, va, 1));
  TRY(pycs_compute_fluxes(h, PYCS_F_Q, PYCS_F_Q));              // :38
  TRY(k_flux_diff(h, 0));                                       // :42-43
  TRY(k_flux_diff(h, 1));
  TRY(k_inner_update(h));                                       // :45-73
  if (h->prm.et != 3) TRY(pycs_halo_fill_copy(h, PYCS_F_QX, PYCS_F_QY));   // :76-78
  TRY(pycs_compute_fluxes(h, PYCS_F_QY, PYCS_F_QX));            // :81 (swapped)
  if (h->prm.mf == 2) TRY(k_average_flux_edges(h));             // :85-86
  TRY(k_flux_diff(h, 0));                                       // :89-90
  TRY(k_flux_diff(h, 1));
  return k_div_and_fix(h);                                      // :95-101
}

extern "C" int pycs_adv_time_step(pycs_handle h, int64_t k, double t) {
  PYCS_RANGE("adv_time_step");
  TRY(wind_sync(h));
  (void)k; (void)t;
  CK(cudaSetDevice(h->device));
  TRY(normalize_q(h));
  TRY(pycs_halo_fill_scalar(h, PYCS_F_Q, PYCS_F_Q));            // src/advection_timestep.py:28
  if (h->prm.vf >= 2) {                                         // :31-37
    TRY(k_wind_ghost_fill(h));
    TRY(k_time_averaged_velocity(h));
  }
  TRY(pycs_divergence(h));                                      // :40
  return k_q_update(h);                                         // :43
}

extern "C" int pycs_update_adv(pycs_handle h, double t) {
  PYCS_RANGE("update_adv");
  TRY(wind_sync(h));
  return k_update_adv(h, t);
}
extern "C" int pycs_init_wind(pycs_handle h) {
  TRY(wind_sync(h));
  return k_wind_interior(h, 0.0, 1, 1);
}
extern "C" int pycs_convert_wind_interior(pycs_handle h) {
  TRY(wind_sync(h));
  return k_wind_interior(h, 0.0, 1, 0);
}

// How the fused step gets a time-dependent wind without the reference's per-step wind kernels (stepper.cu:
// k_fused_step): 1 = wind field 3 with RK1 is U(0) cos(pi t / T) exactly (src/advection_ic.py:301-305),
// scaled inside the step kernel; 2 = fields 2 and 3 are finite sums of static basis fields, combined by
// one kernel per step; 0 = steady winds (fields 1, 4) or PYCS_NO_SEPARABLE (wind kernels every step).
static int lazy_wind_mode(pycs_handle h) {
  if (h->no_separable || !(h->prm.vf == 2 || h->prm.vf == 3) || h->prm.et != 3) return 0;
  return (h->prm.vf == 3 && h->prm.dp == 1) ? 1 : 2;
}

static int run_steps(pycs_handle h, int64_t k0, int64_t nsteps, int fused) {
  CK(cudaSetDevice(h->device));
  if (fused && !k_fused_supported(h)) return arg_fail("no fused step kernel for this scheme tuple");
  if (h->mg && !fused) return arg_fail("multi-GPU handles run the fused step only");
  if (nsteps <= 0) return 0;
  // Lazy wind modes: the steps never touch U_pu / U_pv / U_pc, so those arrays are left behind and
  // caught up lazily (wind_sync) by whatever reads them next -- like pycs_adv_time_step_host, and
  // like the reference's loop, a run has no epilogue.  Stale arrays from an earlier run stay stale.
  const int wmode = fused ? lazy_wind_mode(h) : 0;
  if (!wmode) TRY(wind_sync(h));
  for (int64_t k = k0 + 1; k <= k0 + nsteps; ++k) {
    const double t = (double)k * h->g.dt;                       // t = k*dt, src/advection_sphere.py:47
    if (fused) {
      TRY(k_fused_step(h, k, t, wmode));
    } else {
      TRY(pycs_adv_time_step(h, k, t));
      TRY(k_update_adv(h, t));
    }
  }
  if (wmode) h->wind_stale_k = k0 + nsteps;
  if (fused) k_fused_profile_report(h);
  return 0;
}

extern "C" int pycs_fused_supported(pycs_handle h, int32_t* yes) {
  *yes = k_fused_supported(h);
  return 0;
}

extern "C" int pycs_run(pycs_handle h, int64_t k0, int64_t nsteps, int32_t fused) {
  PYCS_RANGE("pycs_run (adv_sphere time loop)");
  return run_steps(h, k0, nsteps, fused);
}

extern "C" int pycs_run_timed(pycs_handle h, int64_t k0, int64_t nsteps, int32_t fused, float* ms) {
  PYCS_RANGE("pycs_run_timed");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  // several GPUs: the ranks enter the timed steps together -- a barrier on the device, in stream order, so
  // that no rank's timed region contains the wait for a rank whose host got here later
  if (h->mg) TRY(k_mg_device_barrier(h, h->stream));
  CK(cudaEventRecord(h->ev0, h->stream));
  int r = run_steps(h, k0, nsteps, fused);
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaEventSynchronize(h->ev1));
  if (r) return r;
  CK(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return k_mg_check(h);
}

extern "C" int pycs_time_step_kernel(pycs_handle h, int32_t reps, int32_t separable, float* ms) {
  TRY(wind_sync(h));
  CK(cudaSetDevice(h->device));
  if (!k_fused_supported(h)) return arg_fail("no fused step kernel for this scheme tuple");
  TRY(normalize_q(h));
  return k_fused_time_kernel(h, reps, separable, ms);
}

extern "C" int pycs_step_kernel_info(pycs_handle h, int32_t* threads, int32_t* rows_per_chunk, int32_t* nblocks) {
  int a, b, c;
  TRY(k_fused_grid_info(h, &a, &b, &c));
  *threads = a; *rows_per_chunk = b; *nblocks = c;
  return 0;
}

extern "C" int pycs_step_kernel_name(pycs_handle h, char* name, int32_t name_len) {
  if (!name || name_len < 2) return arg_fail("bad name buffer");
  if (!k_fused_supported(h)) return arg_fail("no fused step kernel for this scheme tuple");
  return k_fused_kernel_name(h, name, name_len);
}

extern "C" int pycs_adv_time_step_host(pycs_handle h, double* Q, int64_t k, double t, int32_t fused) {
  PYCS_RANGE("adv_time_step (host buffers)");
  if (!Q) return arg_fail("null Q");
  CK(cudaSetDevice(h->device));
  if (h->mg && !fused) return arg_fail("multi-GPU handles run the fused step only");
  TRY(k_fused_discard(h));                  // the caller's Q replaces the device state
  // A sharded handle moves only its own slab: rows [row_lo, row_hi) of the caller's array go up, the
  // cells the peers read from them (halo rows, ghost-fill sources) are delivered by one extra exchange,
  // and the same rows come back.  The rest of the caller's array is neither read nor written.
  const int i0 = h->mg ? h->row_lo : 0, i1 = h->mg ? h->row_hi : -1;
  TRY(upload_from(h, PYCS_F_Q, Q, i0, i1));
  if (h->mg) TRY(k_fused_share_slab(h));
  if (fused) {
    if (!k_fused_supported(h)) return arg_fail("no fused step kernel for this scheme tuple");
    // lazy wind modes: the step does not touch the exposed wind arrays, they are caught up (wind_sync)
    // when something reads them
    const int wmode = fabs(t - (double)k * h->g.dt) <= 1e-12 * (1.0 + fabs(t)) ? lazy_wind_mode(h) : 0;   // t = k*dt as in adv_sphere
    if (wmode) {
      TRY(k_fused_step(h, k, t, wmode));
      h->wind_stale_k = k;
    } else {
      TRY(wind_sync(h));
      TRY(k_fused_step(h, k, t, 0));
    }
    TRY(normalize_q(h));
  } else {
    TRY(wind_sync(h));
    TRY(pycs_adv_time_step(h, k, t));
    TRY(k_update_adv(h, t));
  }
  TRY(download_to(h, PYCS_F_Q, Q, i0, i1));
  CK(cudaStreamSynchronize(h->stream));
  return k_mg_check(h);
}

// --------------------------------------------------------------------------- multi-GPU
extern "C" int pycs_mgpu_init(pycs_handle h, int32_t rank, int32_t world, unsigned char* handles_out) {
  TRY(wind_sync(h));
  if (!handles_out) return arg_fail("null handle buffer");
  CK(cudaSetDevice(h->device));
  TRY(normalize_q(h));
  return k_mg_init(h, rank, world, handles_out);
}
extern "C" int pycs_mgpu_connect(pycs_handle h, const unsigned char* all_handles) {
  if (!all_handles) return arg_fail("null handle buffer");
  CK(cudaSetDevice(h->device));
  return k_mg_connect(h, all_handles);
}
extern "C" int pycs_mgpu_row_range(pycs_handle h, int32_t* row_lo, int32_t* row_hi) {
  *row_lo = h->row_lo;
  *row_hi = h->row_hi;
  return 0;
}
extern "C" int pycs_mgpu_plan(int32_t N, int32_t world, int32_t rank, int32_t degree, const int32_t* kmin_east,
                              int32_t* row_lo, int32_t* row_hi, int32_t* rects6, int32_t max_rects, int32_t* nrects) {
  if (N < 8 || world < 1 || world > MG_MAX_WORLD || rank < 0 || rank >= world || degree < 0 || degree > 7 || !kmin_east)
    return arg_fail("bad plan arguments");
  const int P = N + 2 * PYCS_NG;
  for (int k = 0; k < 4 * P; ++k)
    if (kmin_east[k] < 0 || kmin_east[k] + degree + 1 > P) return arg_fail("Kmin outside [0, P-order]");
  int a, b;
  pycs_mgpu_rows(N, world, rank, &a, &b);
  *row_lo = a;
  *row_hi = b;
  std::vector<MgRect> rects;
  if (world > 1) pycs_mgpu_plan_rects(N, world, rank, kmin_east, degree + 1, &rects);
  *nrects = (int)rects.size();
  for (int k = 0; k < (int)rects.size() && k < max_rects; ++k) {
    const MgRect& r = rects[k];
    int32_t* o = rects6 + 6 * k;
    o[0] = r.peer; o[1] = r.panel; o[2] = r.i0; o[3] = r.i1; o[4] = r.j0; o[5] = r.j1;
  }
  return 0;
}

extern "C" int pycs_split_plan(int32_t row_lo, int32_t row_hi, int32_t nstrips, int32_t band, int32_t edge_rows,
                               int32_t rows, int32_t* ctas4, int32_t max_ctas, int32_t* n_ctas, int32_t* n_boundary) {
  if (row_hi <= row_lo || nstrips < 1 || !ctas4 || !n_ctas || !n_boundary) return arg_fail("bad split-plan arguments");
  std::vector<CtaDesc> tab;
  *n_boundary = pycs_plan_split_ctas(row_lo, row_hi, nstrips, band, edge_rows, rows, &tab);
  *n_ctas = (int)tab.size();
  for (int k = 0; k < (int)tab.size() && k < max_ctas; ++k) {
    ctas4[4 * k] = tab[k].r0; ctas4[4 * k + 1] = tab[k].r1; ctas4[4 * k + 2] = tab[k].strip; ctas4[4 * k + 3] = tab[k].panel;
  }
  return 0;
}

// --------------------------------------------------------------------------- diagnostics
// The whole-sphere reductions need every row: a sharded handle holds only its slab (the other rows are stale).
static int whole_sphere_only(pycs_handle h, const char* what) {
  if (!h->mg) return 0;
  g_err = std::string(what) + ": this handle is sharded (pycs_mgpu_init) and holds only rows [row_lo, row_hi); "
          "gather the field (parallel.gather_field) and reduce on an unsharded handle or on the host";
  return PYCS_ERR_STATE;
}

extern "C" int pycs_errors(pycs_handle h, const double* qexact_interior, double* out3) {
  PYCS_RANGE("compute_errors");
  TRY(whole_sphere_only(h, "pycs_errors"));
  // stage the (N,N,6) exact field into the interior of USER_B
  const Geo& g = h->g;
  TRY(normalize_q(h));
  std::vector<double> full((size_t)g.P * g.P * 6, 0.0);
  for (int i = 0; i < g.N; ++i)
    memcpy(&full[(((size_t)(i + g.lo)) * g.P + g.lo) * 6], qexact_interior + (size_t)i * g.N * 6,
           sizeof(double) * g.N * 6);
  TRY(pycs_upload_field(h, PYCS_F_USER_B, full.data()));
  double* qe;
  TRY(pycs_field_ptr(h, PYCS_F_USER_B, &qe));
  return k_errors(h, qe, out3);
}

extern "C" int pycs_errors_exact(pycs_handle h, double t, double* out3) {
  TRY(whole_sphere_only(h, "pycs_errors_exact"));
  CK(cudaSetDevice(h->device));
  TRY(normalize_q(h));
  TRY(k_init_tracer(h, PYCS_F_USER_B, t));
  double* qe;
  TRY(pycs_field_ptr(h, PYCS_F_USER_B, &qe));
  return k_errors(h, qe, out3);
}

extern "C" int pycs_mass(pycs_handle h, double* mass) {
  PYCS_RANGE("mass_computation");
  TRY(whole_sphere_only(h, "pycs_mass"));
  TRY(normalize_q(h));
  return k_mass(h, mass);
}
