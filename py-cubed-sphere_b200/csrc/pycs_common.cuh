// Internal definitions shared by the translation units of libpycs_b200.so.
//
// Device layout (DESIGN.md "Data layout in HBM"): every field is panel-major,
// [panel][i][j] with j contiguous.  A row has `ld` doubles, column j lives at
// offset j + JOFF so that the first interior column (j0 = 4) starts a 128-byte
// line; a panel has P+1 rows (enough for x-edge fields) and `ld` >= JOFF+P+1
// (enough for y-edge fields), so all fields share one geometry.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include "../../include/pycs_b200.h"

#define PYCS_JOFF 12
#define PYCS_NG 4          // ngl = ngr = 4  (src/cs_datastruct.py:225-231)
// Period T of the Nair-Lauritzen wind fields 2 and 3: a literal of velocity_adv (src/advection_ic.py:295,302),
// NOT the run length Tf.  wind.cu evaluates the fields with it, fused.cu the separable time factor cos(pi t / T).
#define PYCS_WIND_PERIOD 5.0
#define PYCS_PI 3.141592653589793

struct Geo {
  int N, P, ld, lo, hi;    // lo = i0 = j0 = 4, hi = iend = jend = N + 4
  long long ps;            // panel stride in doubles = (P + 1) * ld
  double dx, dy, dt;
};

__host__ __device__ __forceinline__ long long gidx(const Geo& g, int p, int i, int j) {
  return (long long)p * g.ps + (long long)i * g.ld + (j + PYCS_JOFF);
}

// Affine source map of one (panel, side) halo strip: entry [a][b] of the halo
// array comes from panel nb at (ci + ai*a + bi*b, cj + aj*a + bj*b); rot = the
// neighbour's axes are transposed w.r.t. ours (x/y field swap of
// src/halo_data.py:224,258,371,388).
struct SideMap { int nb, ci, ai, bi, cj, aj, bj, rot; };
struct HaloMaps { SideMap m[6][4]; };   // side: 0 E, 1 W, 2 N, 3 S

enum { SIDE_E = 0, SIDE_W = 1, SIDE_N = 2, SIDE_S = 3 };

struct MgpuState;
struct pycs_handle_s {
  pycs_params prm;
  Geo g;
  int device;
  int sm_count;
  cudaStream_t stream;
  double* f[PYCS_F_COUNT];
  // Lagrange tables (east), device
  int degree, order;
  int* kminE;
  double* wE;
  int* kmin_host;       // host copy of the stencil table: the multi-GPU exchange plan is derived from it
  HaloMaps maps;
  // halo gather buffers for the copy fill: E, W (4,P,6 each) then N, S (P,4,6)
  double* halo_buf;
  // reductions
  double* red_part;     // per-block partials
  double* red_out;      // small device result vector
  int red_blocks;
  // staging for reference-layout transfers
  double* stage_dev;
  size_t stage_bytes;
  double* stage_pin;
  size_t pin_bytes;
  long long launches;
  // fused step bookkeeping
  int qcur;             // which of Q / Q_NEXT currently holds the state
  cudaEvent_t ev0, ev1;
  float last_step_kernel_ms;
  long long last_step_kernel_launches;
  // deferred MF-PR correction scalar (device): m0 / a2 of the previous fused step
  double a2;            // sum sqrtg^2 over the interior (static)
  int a2_valid;
  // rows this handle updates: [lo, hi) or, with pycs_mgpu_init, this rank's slab of every panel
  int row_lo, row_hi;
  // >= 0: U_pu / U_pv / U_pc still hold an older step's winds; they must be brought to the state
  // after update_adv(t_k), k = wind_stale_k, before anything reads them (capi.cu: wind_sync)
  long long wind_stale_k;
  int no_separable;     // PYCS_NO_SEPARABLE: refresh the wind of field 3 every step instead of scaling it in-kernel
  int dg_two_phase;     // PYCS_DG_TWO_PHASE: pycs_halo_fill_dg as the reference's two phases (two launches)
  struct MgpuState* mg;   // multi-GPU state (mgpu.cu), null on a single GPU
};

// error plumbing ------------------------------------------------------------
void pycs_set_error(const std::string& msg);
int pycs_cuda_fail(cudaError_t e, const char* what, const char* file, int line);
// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------
// The kernels of a serial step (ghost fill -> [basis winds] -> step kernel -> next ghost fill ...) are chained in
// one stream.  Launched with the programmatic-stream-serialization attribute, a kernel's CTAs may become resident
// while its predecessor drains -- launch latency, CTA prologue (shared-memory set-up) and the predecessor's tail
// overlap -- and every such kernel calls pdl_wait() before its first global access: it returns once the
// predecessor grid has completed and flushed (transitively everything before it).  pdl_trigger() at the top of a
// kernel lets its own successor be scheduled as soon as all of this kernel's CTAs have started.  Without the
// attribute both calls are no-ops.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t pycs_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                                   Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

#define CK(call)                                                                 \
  do {                                                                           \
    cudaError_t e__ = (call);                                                    \
    if (e__ != cudaSuccess) return pycs_cuda_fail(e__, #call, __FILE__, __LINE__); \
  } while (0)
#define CKL(h)                                                        \
  do {                                                                \
    (h)->launches++;                                                  \
    cudaError_t e__ = cudaGetLastError();                             \
    if (e__ != cudaSuccess) return pycs_cuda_fail(e__, "kernel launch", __FILE__, __LINE__); \
  } while (0)
#define TRY(call)            \
  do {                       \
    int r__ = (call);        \
    if (r__ != 0) return r__; \
  } while (0)

int pycs_field_shape(const Geo& g, int field, int* ni, int* nj, int* npanels);
int pycs_field_ptr(pycs_handle h, int field, double** out);   // lazy allocation
inline bool pycs_field_is_u(int f) {
  return f == PYCS_F_CX || (f >= PYCS_F_PX_FL && f <= PYCS_F_PX_FUPW) ||
         (f >= PYCS_F_PU_ULON && f <= PYCS_F_PU_UOLD) || f == PYCS_F_SQRTG_PU ||
         (f >= PYCS_F_PU_EXLON && f <= PYCS_F_PU_DET) || f == PYCS_F_PU_LON || f == PYCS_F_PU_LAT;
}
inline bool pycs_field_is_v(int f) {
  return f == PYCS_F_CY || (f >= PYCS_F_PY_FL && f <= PYCS_F_PY_FUPW) ||
         (f >= PYCS_F_PV_ULON && f <= PYCS_F_PV_VOLD) || f == PYCS_F_SQRTG_PV ||
         (f >= PYCS_F_PV_EXLON && f <= PYCS_F_PV_DET) || f == PYCS_F_PV_LON || f == PYCS_F_PV_LAT;
}
inline bool pycs_field_single_panel(int f) { return f >= PYCS_F_SQRTG_PC && f <= PYCS_F_SQRTG_PV; }

// launchers implemented in the kernel translation units -----------------------
// halo.cu
void pycs_build_halo_maps(const Geo& g, HaloMaps* maps);
int k_halo_gather(pycs_handle h, const double* fx, const double* fy, double* buf);
int k_halo_scatter_copy(pycs_handle h, double* fx, double* fy, const double* buf);
int k_dg_fill(pycs_handle h, double* q);          // two launches (halo.cu)
int k_dg_fill_single(pycs_handle h, double* q);   // one launch (stepper.cu)
// ppm.cu
int k_cfl(pycs_handle h, double* dst, const double* src, int dir);
int k_mul_metric(pycs_handle h, double* gq, const double* q);
int k_recon(pycs_handle h, const double* qx, const double* qy);
int k_edges_extrapolation(pycs_handle h, const double* qx, const double* qy);
int k_flux(pycs_handle h, const double* qx, const double* qy);
int k_flux_diff(pycs_handle h, int dir);
int k_inner_update(pycs_handle h);
int k_average_flux_edges(pycs_handle h);
int k_div_and_fix(pycs_handle h);
int k_q_update(pycs_handle h);
int k_errors(pycs_handle h, const double* qexact_dev, double* out3_host);
int k_mass(pycs_handle h, double* out_host);
int k_sum_sq_metric(pycs_handle h, double* out_host);
// grid.cu
int k_generate_geometry(pycs_handle h, const double* xc, const double* xe);
int k_init_tracer(pycs_handle h, int field, double t);
int k_rect_max(pycs_handle h, const double* f, int i0, int i1, int j0, int j1, double* out_host);
// wind.cu
int k_time_averaged_velocity(pycs_handle h);
int k_wind_ghost_fill(pycs_handle h);
int k_wind_edges2center(pycs_handle h);
int k_wind_center2ghostedge(pycs_handle h);
int k_update_adv(pycs_handle h, double t);
int k_wind_interior(pycs_handle h, double t, int convert_interior_only, int do_velocity, int basis = -1);
int k_wind_basis_count(pycs_handle h);
int k_wind_basis_build(pycs_handle h, int m);
int k_wind_basis_combine(pycs_handle h, double* const* bu, double* const* bv, int nb, double* ua, double* um, double* va,
                         double* vm, const double* coef, int cmask, const long long* steps, int pdl = 0);
int k_wind_coef_fill(pycs_handle h, double* tab, int cmask, long long s0, long long k0, int n);
// stepper.cu
int k_fused_supported(pycs_handle h);
int k_fused_step(pycs_handle h, long long k, double t, int separable);
int k_wind_catch_up(pycs_handle h, long long k);
int k_fused_time_kernel(pycs_handle h, int reps, int separable, float* ms);
int k_fused_grid_info(pycs_handle h, int* tb, int* rows, int* nblocks);
int k_fused_kernel_name(pycs_handle h, char* out, int len);
int k_fused_flush(pycs_handle h);
int k_fused_share_slab(pycs_handle h);              // multi-GPU: deliver the freshly uploaded slab's boundary cells to the peers
int k_fused_discard(pycs_handle h);                 // a new Q was uploaded: drop what the fused path had pending
void k_fused_profile_report(pycs_handle h);         // PYCS_STEP_PROFILE: print and reset the per-kernel device times
void k_fused_release(pycs_handle h);
void k_fused_invalidate(pycs_handle h);
void k_fused_invalidate_ghost_metric(pycs_handle h); // the Lagrange tables changed
void k_fused_invalidate_graphs(pycs_handle h);       // dt changed: it is baked into the captured launches
// layout.cu (in capi.cu)
