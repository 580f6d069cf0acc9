/*
 * pycs_b200.h -- C ABI of the B200-native cubed-sphere PPM advection step.
 *
 * The reference (luanfs/py-cubed-sphere) has no FFI layer: its "plugin API" is
 * the Python function surface of src/*.py acting in place on numpy arrays held
 * by `cs_grid` and `simulation`.  Each entry point below names
 * the reference function it replaces (file:line under /root/reference).  The
 * Python shims in py-cubed-sphere_b200/*.py bind these with ctypes and keep the
 * reference's names and argument order; INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *  - every function returns 0 on success, a negative pycs_status otherwise;
 *    pycs_last_error() gives the message of the last failure on this thread;
 *  - host arrays are fp64, C order, in the REFERENCE layout [i][j][panel]
 *    (panel fastest) with the reference shapes: centre fields (P,P,6),
 *    x-edge fields (P+1,P,6), y-edge fields (P,P+1,6), P = N + 8;
 *  - the device owns the state: one handle = one GPU + one CUDA stream; a
 *    handle is confined to the thread that uses it (the reference is single
 *    threaded, no re-entrancy);
 *  - there is NO CPU fallback: without a CUDA device pycs_create fails.
 */
#ifndef PYCS_B200_H
#define PYCS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pycs_handle_s* pycs_handle;

typedef enum {
  PYCS_OK = 0,
  PYCS_ERR_ARG = -1,       /* bad argument / unknown field / bad scheme combination */
  PYCS_ERR_CUDA = -2,      /* CUDA runtime error (message has the CUDA string) */
  PYCS_ERR_STATE = -3,     /* call order (e.g. tables not uploaded for ET-DG) */
  PYCS_ERR_NOMEM = -4
} pycs_status;

/* Scheme selectors: the integers of par/advection.par, mapped to names in
 * adv_simulation_par.__init__ (src/advection_ic.py:85-149). */
typedef struct {
  int32_t N;        /* cells per panel edge                                   */
  int32_t recon;    /* 1 PPM-0, 2 PPM-CW84, 3 PPM-PL07, 4 PPM-L04             */
  int32_t dp;       /* 1 RK1, 2 RK2                                           */
  int32_t opsplit;  /* 1 SP-AVLT, 2 SP-L04, 3 SP-PL07                         */
  int32_t et;       /* 1 ET-S72, 2 ET-PL07, 3 ET-DG                           */
  int32_t mt;       /* 1 MT-0, 2 MT-PL07                                      */
  int32_t mf;       /* 1 MF-0, 2 MF-AF, 3 MF-PR                               */
  int32_t vf;       /* wind field 1..4 (src/advection_ic.py:287-313)          */
  int32_t ic;       /* initial condition 1..4 (src/advection_ic.py:215-281)   */
  int32_t device;   /* CUDA device ordinal                                    */
  double dt;        /* time step                                              */
  double dx, dy;    /* cs_grid.dx, cs_grid.dy (src/cs_datastruct.py:220-222)  */
} pycs_params;

/* Field ids.  Shapes: C = centre (P,P,6), U = x-edge (P+1,P,6), V = y-edge (P,P+1,6). */
enum {
  PYCS_F_Q = 0,          /* C simulation.Q                  src/advection_ic.py:178 */
  PYCS_F_Q_NEXT = 1,     /* C double buffer of the fused step (device only)         */
  PYCS_F_GQ = 2,         /* C simulation.gQ                 :179                    */
  PYCS_F_DIV = 3,        /* C simulation.div                :182                    */
  PYCS_F_CX = 4,         /* U simulation.cx                 :190                    */
  PYCS_F_CY = 5,         /* V simulation.cy                 :191                    */
  PYCS_F_QX = 6,         /* C inner-operator result Qx      src/discrete_operators.py:72 */
  PYCS_F_QY = 7,         /* C inner-operator result Qy                               */
  PYCS_F_PX_QL = 8, PYCS_F_PX_QR = 9, PYCS_F_PX_DQ = 10, PYCS_F_PX_Q6 = 11,   /* C  px.*  src/cs_datastruct.py:649-652 */
  PYCS_F_PX_FL = 12, PYCS_F_PX_FR = 13, PYCS_F_PX_FUPW = 14,                  /* U  :656-658 */
  PYCS_F_PX_DF = 15,                                                          /* C  :674 */
  PYCS_F_PY_QL = 16, PYCS_F_PY_QR = 17, PYCS_F_PY_DQ = 18, PYCS_F_PY_Q6 = 19, /* C  py.* */
  PYCS_F_PY_FL = 20, PYCS_F_PY_FR = 21, PYCS_F_PY_FUPW = 22,                  /* V  :665-667 */
  PYCS_F_PY_DF = 23,                                                          /* C */
  PYCS_F_PU_ULON = 24, PYCS_F_PU_VLAT = 25, PYCS_F_PU_UCONTRA = 26, PYCS_F_PU_VCONTRA = 27,
  PYCS_F_PU_UAVG = 28, PYCS_F_PU_UOLD = 29,                                   /* U  U_pu.* :707-713 */
  PYCS_F_PV_ULON = 30, PYCS_F_PV_VLAT = 31, PYCS_F_PV_UCONTRA = 32, PYCS_F_PV_VCONTRA = 33,
  PYCS_F_PV_VAVG = 34, PYCS_F_PV_VOLD = 35,                                   /* V  U_pv.* :717-722 */
  PYCS_F_PC_ULON = 36, PYCS_F_PC_VLAT = 37, PYCS_F_PC_UCONTRA = 38, PYCS_F_PC_VCONTRA = 39, /* C U_pc.* */
  PYCS_F_SQRTG_PC = 40,  /* C cs_grid.metric_tensor_pc (panel 0 is kept; identical on all panels, src/cs_datastruct.py:417) */
  PYCS_F_SQRTG_PU = 41,  /* U cs_grid.metric_tensor_pu */
  PYCS_F_SQRTG_PV = 42,  /* V cs_grid.metric_tensor_pv */
  PYCS_F_PC_EXLON = 43, PYCS_F_PC_EXLAT = 44, PYCS_F_PC_EYLON = 45, PYCS_F_PC_EYLAT = 46, PYCS_F_PC_DET = 47, /* C prod_e*_pc, determinant_ll2contra_pc :456-467 */
  PYCS_F_PU_EXLON = 48, PYCS_F_PU_EXLAT = 49, PYCS_F_PU_EYLON = 50, PYCS_F_PU_EYLAT = 51, PYCS_F_PU_DET = 52, /* U :469-480 */
  PYCS_F_PV_EXLON = 53, PYCS_F_PV_EXLAT = 54, PYCS_F_PV_EYLON = 55, PYCS_F_PV_EYLAT = 56, PYCS_F_PV_DET = 57, /* V :482-493 */
  PYCS_F_PC_LON = 58, PYCS_F_PC_LAT = 59,   /* C cs_grid.pc.lon/lat */
  PYCS_F_PU_LON = 60, PYCS_F_PU_LAT = 61,   /* U cs_grid.pu.lon/lat */
  PYCS_F_PV_LON = 62, PYCS_F_PV_LAT = 63,   /* V cs_grid.pv.lon/lat */
  PYCS_F_USER_A = 64, PYCS_F_USER_B = 65,   /* C caller-provided Qx / Qy for operator-level calls */
  PYCS_F_COUNT = 66
};

/* ---- lifetime -------------------------------------------------------------- */
int pycs_create(const pycs_params* params, pycs_handle* out);
int pycs_destroy(pycs_handle h);
const char* pycs_last_error(void);
/* Library / device facts for the harness (SM count, device name). */
int pycs_device_info(pycs_handle h, int32_t* sm_count, char* name, int32_t name_len);

/* ---- host <-> device (reference layout on the host side) -------------------- */
int pycs_upload_field(pycs_handle h, int32_t field, const double* host);
int pycs_download_field(pycs_handle h, int32_t field, double* host);
int pycs_copy_field(pycs_handle h, int32_t dst, int32_t src);
int pycs_fill_field(pycs_handle h, int32_t field, double value);
/* Lagrange tables of lagrange_poly_ghostcell_pc (src/lagrange.py:28-163): the
 * EAST stencil start Kmin (4,P) and weights (4,P,degree+1); the other three
 * sides are flips / transposes of these (src/lagrange.py:144-152). */
int pycs_upload_lagrange(pycs_handle h, int32_t degree, const int32_t* kmin_east,
                         const double* weights_east);
int pycs_set_dt(pycs_handle h, double dt);

/* ---- device-side set-up ("next" rows f2 / f3) ---------------------------------- */
/* The grid fields the path reads -- sqrt(g) at pc / pu / pv (src/cs_datastruct.py:407-446), the lat-lon <->
 * contravariant coefficients and determinant (:448-493) and the point coordinates lon / lat (:240-324,
 * src/cs_transform.py:41-96) -- generated on the device from the 1-D coordinate arrays of the equiangular
 * grid: xc = centres (P values), xe = edges (P + 1 values), as np.linspace gives them.  Replaces the 24
 * pycs_upload_field calls of the host-built grid; agrees with it to the last one or two ulps. */
int pycs_generate_geometry(pycs_handle h, const double* xc, const double* xe);
/* q0_adv / qexact_adv (src/advection_ic.py:215-281) at time t into the interior of a centre field
 * (ghost cells zero), evaluated on the device from pc.lon / pc.lat. */
int pycs_init_tracer(pycs_handle h, int32_t field, double t);
/* max over [i0,i1) x [j0,j1) x panels of a field, e.g. the CFL number of init_vars_adv
 * (src/advection_vars.py:89-98: amax of cx[i0:iend+1,:,:], no abs inside). */
int pycs_field_max(pycs_handle h, int32_t field, int32_t i0, int32_t i1, int32_t j0, int32_t j1, double* out);

/* ---- halo / edges (L1 of SURVEY.md s1) -------------------------------------- */
/* get_halo_data_interpolation (src/halo_data.py:15-185): out_* are (4,P,6),
 * (4,P,6), (P,4,6), (P,4,6) host arrays.  With fx != fy it is the _WE/_NS pair
 * (src/halo_data.py:191-400). */
int pycs_halo_gather(pycs_handle h, int32_t fx, int32_t fy, double* east, double* west,
                     double* north, double* south);
/* ghost_cell_pc_lagrange_interpolation (src/interpolation.py:154-314), in place. */
int pycs_halo_fill_dg(pycs_handle h, int32_t field);
/* ghost_cells_adjacent_panels (src/interpolation.py:320-340), in place on (fx, fy). */
int pycs_halo_fill_copy(pycs_handle h, int32_t fx, int32_t fy);
/* edges_ghost_cell_treatment_scalar (src/edges_treatment.py:284-290): dispatch on et. */
int pycs_halo_fill_scalar(pycs_handle h, int32_t fx, int32_t fy);
/* edges_ghost_cell_treatment_vector (src/edges_treatment.py:296-347) on U_pu, U_pv, U_pc. */
int pycs_halo_fill_vector(pycs_handle h);

/* The two stages of the duo-grid wind ghost fill on their own: wind_edges2center_cubic_interpolation
 * (src/interpolation.py:347-430: U_pu / U_pv -> U_pc on the boundary ring, lat-lon, Lagrange ghost fill) and
 * wind_center2ghostedge_cubic_interpolation (src/interpolation.py:436-532: U_pc ghost centres -> ghost edges of
 * U_pu / U_pv, contravariant). */
int pycs_wind_edges2center(pycs_handle h);
int pycs_wind_center2ghostedge(pycs_handle h);

/* ---- operators (L2) ---------------------------------------------------------- */
/* time_averaged_velocity (src/averaged_velocity.py:14-62). */
int pycs_time_averaged_velocity(pycs_handle h);
/* cfl_x / cfl_y (src/cfl.py:10-19): dst = src * dt / dx (dy). */
int pycs_cfl(pycs_handle h, int32_t dst, int32_t src, int32_t dir);
/* ppm_reconstruction (src/reconstruction_1d.py:387-394) of (fx, fy) into px / py. */
int pycs_ppm_reconstruction(pycs_handle h, int32_t fx, int32_t fy);
/* edges_extrapolation (src/edges_treatment.py:82-206, with average_parabola_cube_edges :31-76): the ET-PL07
 * one-sided edge values next to the cube edges of px / py reconstructed from (fx, fy), averaged across the
 * edge and copied into the ghost parabolas.  pycs_ppm_reconstruction calls it itself when et == 2. */
int pycs_edges_extrapolation(pycs_handle h, int32_t fx, int32_t fy);
/* numerical_flux_ppm_x / _y (src/flux.py:20-128) after a reconstruction. */
int pycs_numerical_flux(pycs_handle h, int32_t fx, int32_t fy);
/* compute_fluxes (src/flux.py:9-15) = reconstruction + both fluxes. */
int pycs_compute_fluxes(pycs_handle h, int32_t fx, int32_t fy);
/* F_operator / G_operator (src/discrete_operators.py:109-138). */
int pycs_F_operator(pycs_handle h);
int pycs_G_operator(pycs_handle h);
/* average_flux_cube_edges (src/edges_treatment.py:231-278). */
int pycs_average_flux_cube_edges(pycs_handle h);
/* divergence (src/discrete_operators.py:18-101): operator by operator. */
int pycs_divergence(pycs_handle h);

/* ---- time step (L3) ----------------------------------------------------------- */
/* adv_time_step (src/advection_timestep.py:19-43), operator-by-operator path. */
int pycs_adv_time_step(pycs_handle h, int64_t k, double t);
/* update_adv (src/advection_timestep.py:48-75); no-op for vf == 1. */
int pycs_update_adv(pycs_handle h, double t);
/* Wind at time t on the interior edges + conversion, the first block of
 * init_vars_adv (src/advection_vars.py:37-53). */
int pycs_init_wind(pycs_handle h);
/* Only the latlon_to_contravariant part of that block (src/advection_vars.py:44-53),
 * for callers that uploaded U_pu/U_pv ulon, vlat themselves. */
int pycs_convert_wind_interior(pycs_handle h);
/* The hot loop of adv_sphere (src/advection_sphere.py:45-57) without output:
 * for k = k0+1 .. k0+nsteps: adv_time_step(k, k*dt); update_adv(k*dt).
 * fused != 0 selects the fused step kernel (ET-DG schemes; falls back with
 * PYCS_ERR_ARG if the scheme has no fused kernel), 0 the operator path. */
int pycs_run(pycs_handle h, int64_t k0, int64_t nsteps, int32_t fused);
/* 1 in *yes when the scheme tuple of this handle has a fused step kernel. */
int pycs_fused_supported(pycs_handle h, int32_t* yes);
/* Same loop, timed on the handle's stream with CUDA events (ms for all steps). */
int pycs_run_timed(pycs_handle h, int64_t k0, int64_t nsteps, int32_t fused, float* ms);
/* One step through HOST buffers: upload Q (P,P,6), step k, download Q.  The
 * end-to-end entry point a numpy caller of adv_time_step sees. */
int pycs_adv_time_step_host(pycs_handle h, double* Q_inout, int64_t k, double t, int32_t fused);
int pycs_synchronize(pycs_handle h);

/* ---- multi-GPU (the reference is single process) ------------------------ */
/* One process per GPU.  Rank `rank` of `world` (2..8) owns a slab of rows of every panel;
 * in each fused step it stores the cells its peers read (halo rows, sources of their ghost cells)
 * straight into the peers' Q arrays (CUDA IPC over NVLink) beside its interior update and raises
 * flags the peers' kernels wait on.
 * pycs_mgpu_init: after the state and the Lagrange tables are uploaded; returns 3 cudaIpcMemHandle_t (192 bytes).  The
 * caller all-gathers them (torch.distributed) and passes the world*192 bytes to
 * pycs_mgpu_connect.  Afterwards pycs_run(fused=1) advances the own slab; pycs_download_field
 * returns an array whose rows [row_lo,row_hi) (pycs_mgpu_row_range, padded-panel index) are valid. */
int pycs_mgpu_init(pycs_handle h, int32_t rank, int32_t world, unsigned char* handles_out);
int pycs_mgpu_connect(pycs_handle h, const unsigned char* all_handles);
int pycs_mgpu_row_range(pycs_handle h, int32_t* row_lo, int32_t* row_hi);
/* Host-only: the slab of one rank and the rectangles (peer, panel, i0, i1, j0, j1) it stores into its
 * peers after every step -- the 3 rows next to their slabs and the interior cells their ghost cells are
 * interpolated from, derived from the halo index maps (src/halo_data.py:15-185) and the Lagrange stencil
 * table kmin_east (4, P) of the given degree (src/lagrange.py:28-163).  *nrects may exceed max_rects. */
int pycs_mgpu_plan(int32_t N, int32_t world, int32_t rank, int32_t degree, const int32_t* kmin_east,
                   int32_t* row_lo, int32_t* row_hi, int32_t* rects6, int32_t max_rects, int32_t* nrects);

/* Host-only: the CTA table of a split fused step (several GPUs, or PYCS_SPLIT=1) over the rows
 * [row_lo, row_hi) of every panel cut into nstrips column strips.  Entries (r0, r1, strip, panel); the
 * first *n_boundary are the boundary CTAs -- bands of `band` rows at both ends of the slab over all strips,
 * then the first and last strip in chunks of `edge_rows` rows: everything a ghost cell or a peer reads,
 * launched first and done early -- the rest are the interior CTAs (chunks of `rows` rows), which stage no
 * ghost cell and no row outside the slab and run beside the exchange and the ghost fill of
 * src/advection_timestep.py:28.  *n_ctas may exceed max_ctas. */
int pycs_split_plan(int32_t row_lo, int32_t row_hi, int32_t nstrips, int32_t band, int32_t edge_rows, int32_t rows,
                    int32_t* ctas4, int32_t max_ctas, int32_t* n_ctas, int32_t* n_boundary);

/* ---- diagnostics (next row f2) -------------------------------------------------- */
/* compute_errors (src/errors.py:99-113) of Q against a host reference field
 * qexact given on the interior (N,N,6): out = {Linf, L1, L2}. */
int pycs_errors(pycs_handle h, const double* qexact_interior, double* out3);
/* The same against qexact_adv(t) evaluated on the device (no host field, no transfer). */
int pycs_errors_exact(pycs_handle h, double t, double* out3);
/* mass_computation (src/diagnostics.py:14-26): sum Q*sqrtg*dx*dy over the interior. */
int pycs_mass(pycs_handle h, double* mass);
/* Number of kernels launched by this handle since creation (bench bookkeeping). */
int pycs_launch_count(pycs_handle h, int64_t* count);
/* Device time (ms) of `reps` back-to-back launches of the fused step kernel alone
 * (no ghost fill; ping-pong buffers), for the roofline figure.  separable != 0 times
 * the scaled-wind variant.  Leaves Q undefined: upload Q again afterwards. */
int pycs_time_step_kernel(pycs_handle h, int32_t reps, int32_t separable, float* ms);
/* Launch geometry of the fused step kernel: threads per CTA, rows per chunk, CTAs. */
int pycs_step_kernel_info(pycs_handle h, int32_t* threads, int32_t* rows_per_chunk, int32_t* nblocks);
/* Which fused step kernel (and tuning point) this handle launches. */
int pycs_step_kernel_name(pycs_handle h, char* name, int32_t name_len);

#ifdef __cplusplus
}
#endif
#endif /* PYCS_B200_H */
