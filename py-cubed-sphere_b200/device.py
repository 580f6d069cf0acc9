"""ctypes binding of libpycs_b200.so and the device-array proxies.

`DeviceArray` stands where the reference holds a numpy array inside
`simulation` / `px` / `py` / `U_pu` ...: it names a device field and converts to
numpy (reference layout `[i][j][panel]`) on demand, so `simulation.Q[i0:iend]`
or `np.asarray(simulation.px.q_L)` read like the reference while the state stays
in HBM.  Operators given plain numpy arrays upload them, run on the device and
write the result back into the caller's array (the reference mutates in place).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpycs_b200.so")


class PycsError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("N", "recon", "dp", "opsplit", "et", "mt", "mf", "vf", "ic", "device")] + \
               [("dt", C.c_double), ("dx", C.c_double), ("dy", C.c_double)]


# field ids of include/pycs_b200.h
F = dict(
    Q=0, Q_NEXT=1, GQ=2, DIV=3, CX=4, CY=5, QX=6, QY=7,
    PX_QL=8, PX_QR=9, PX_DQ=10, PX_Q6=11, PX_FL=12, PX_FR=13, PX_FUPW=14, PX_DF=15,
    PY_QL=16, PY_QR=17, PY_DQ=18, PY_Q6=19, PY_FL=20, PY_FR=21, PY_FUPW=22, PY_DF=23,
    PU_ULON=24, PU_VLAT=25, PU_UCONTRA=26, PU_VCONTRA=27, PU_UAVG=28, PU_UOLD=29,
    PV_ULON=30, PV_VLAT=31, PV_UCONTRA=32, PV_VCONTRA=33, PV_VAVG=34, PV_VOLD=35,
    PC_ULON=36, PC_VLAT=37, PC_UCONTRA=38, PC_VCONTRA=39,
    SQRTG_PC=40, SQRTG_PU=41, SQRTG_PV=42,
    PC_EXLON=43, PC_EXLAT=44, PC_EYLON=45, PC_EYLAT=46, PC_DET=47,
    PU_EXLON=48, PU_EXLAT=49, PU_EYLON=50, PU_EYLAT=51, PU_DET=52,
    PV_EXLON=53, PV_EXLAT=54, PV_EYLON=55, PV_EYLAT=56, PV_DET=57,
    PC_LON=58, PC_LAT=59, PU_LON=60, PU_LAT=61, PV_LON=62, PV_LAT=63,
    USER_A=64, USER_B=65,
)
_U_FIELDS = {4, 12, 13, 14, 24, 25, 26, 27, 28, 29, 41, 48, 49, 50, 51, 52, 60, 61}
_V_FIELDS = {5, 20, 21, 22, 30, 31, 32, 33, 34, 35, 42, 53, 54, 55, 56, 57, 62, 63}

_lib = None


def load_library():
    """Load libpycs_b200.so; raise (never fall back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PycsError(
            "libpycs_b200.so not built (%s). Run `python py-cubed-sphere_b200/build.py`; "
            "there is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    dp = C.POINTER(C.c_double)
    h = C.c_void_p
    sig = {
        "pycs_create": [C.POINTER(Params), C.POINTER(h)],
        "pycs_destroy": [h],
        "pycs_device_info": [h, C.POINTER(C.c_int32), C.c_char_p, C.c_int32],
        "pycs_upload_field": [h, C.c_int32, dp],
        "pycs_download_field": [h, C.c_int32, dp],
        "pycs_copy_field": [h, C.c_int32, C.c_int32],
        "pycs_fill_field": [h, C.c_int32, C.c_double],
        "pycs_upload_lagrange": [h, C.c_int32, C.POINTER(C.c_int32), dp],
        "pycs_set_dt": [h, C.c_double],
        "pycs_halo_gather": [h, C.c_int32, C.c_int32, dp, dp, dp, dp],
        "pycs_halo_fill_dg": [h, C.c_int32],
        "pycs_halo_fill_copy": [h, C.c_int32, C.c_int32],
        "pycs_halo_fill_scalar": [h, C.c_int32, C.c_int32],
        "pycs_halo_fill_vector": [h],
        "pycs_wind_edges2center": [h],
        "pycs_wind_center2ghostedge": [h],
        "pycs_edges_extrapolation": [h, C.c_int32, C.c_int32],
        "pycs_time_averaged_velocity": [h],
        "pycs_cfl": [h, C.c_int32, C.c_int32, C.c_int32],
        "pycs_ppm_reconstruction": [h, C.c_int32, C.c_int32],
        "pycs_numerical_flux": [h, C.c_int32, C.c_int32],
        "pycs_compute_fluxes": [h, C.c_int32, C.c_int32],
        "pycs_F_operator": [h],
        "pycs_G_operator": [h],
        "pycs_average_flux_cube_edges": [h],
        "pycs_divergence": [h],
        "pycs_adv_time_step": [h, C.c_int64, C.c_double],
        "pycs_update_adv": [h, C.c_double],
        "pycs_init_wind": [h],
        "pycs_convert_wind_interior": [h],
        "pycs_run": [h, C.c_int64, C.c_int64, C.c_int32],
        "pycs_fused_supported": [h, C.POINTER(C.c_int32)],
        "pycs_run_timed": [h, C.c_int64, C.c_int64, C.c_int32, C.POINTER(C.c_float)],
        "pycs_adv_time_step_host": [h, dp, C.c_int64, C.c_double, C.c_int32],
        "pycs_synchronize": [h],
        "pycs_errors": [h, dp, dp],
        "pycs_errors_exact": [h, C.c_double, dp],
        "pycs_generate_geometry": [h, dp, dp],
        "pycs_init_tracer": [h, C.c_int32, C.c_double],
        "pycs_field_max": [h, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, dp],
        "pycs_mass": [h, dp],
        "pycs_launch_count": [h, C.POINTER(C.c_int64)],
        "pycs_time_step_kernel": [h, C.c_int32, C.c_int32, C.POINTER(C.c_float)],
        "pycs_step_kernel_info": [h, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)],
        "pycs_step_kernel_name": [h, C.c_char_p, C.c_int32],
        "pycs_mgpu_init": [h, C.c_int32, C.c_int32, C.c_char_p],
        "pycs_mgpu_connect": [h, C.c_char_p],
        "pycs_mgpu_row_range": [h, C.POINTER(C.c_int32), C.POINTER(C.c_int32)],
    }
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.pycs_mgpu_plan.argtypes = [C.c_int32] * 4 + [C.POINTER(C.c_int32)] * 4 + [C.c_int32, C.POINTER(C.c_int32)]
    lib.pycs_mgpu_plan.restype = C.c_int
    lib.pycs_last_error.restype = C.c_char_p
    lib.pycs_last_error.argtypes = []
    _lib = lib
    return lib


EXPORTED_SYMBOLS = None  # filled by tests from include/pycs_b200.h


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Device:
    """One pycs handle (one GPU, one stream)."""

    def __init__(self, N, dx, dy, dt, recon=3, dp=1, opsplit=1, et=3, mt=1, mf=3, vf=1, ic=2, device=0):
        self.lib = load_library()
        self.N, self.P = N, N + 8
        self.params = Params(N, recon, dp, opsplit, et, mt, mf, vf, ic, device, dt, dx, dy)
        self.h = C.c_void_p()
        self._check(self.lib.pycs_create(C.byref(self.params), C.byref(self.h)))

    def _check(self, rc):
        if rc != 0:
            raise PycsError("pycs error %d: %s" % (rc, self.lib.pycs_last_error().decode()))

    def call(self, name, *args):
        self._check(getattr(self.lib, name)(self.h, *args))

    def close(self):
        if self.h:
            self.lib.pycs_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- transfers ---------------------------------------------------------
    def shape(self, fid):
        P = self.P
        return (P + 1 if fid in _U_FIELDS else P, P + 1 if fid in _V_FIELDS else P, 6)

    def upload(self, fid, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        if a.shape != self.shape(fid):
            raise PycsError("field %d expects shape %s, got %s" % (fid, self.shape(fid), a.shape))
        self.call("pycs_upload_field", fid, _dptr(a))

    def download(self, fid, out=None):
        if out is None:
            out = np.empty(self.shape(fid))
        if not (out.flags.c_contiguous and out.dtype == np.float64 and out.shape == self.shape(fid)):
            tmp = np.empty(self.shape(fid))
            self.call("pycs_download_field", fid, _dptr(tmp))
            out[...] = tmp
            return out
        self.call("pycs_download_field", fid, _dptr(out))
        return out

    def array(self, fid):
        return DeviceArray(self, fid)

    def field_max(self, fid, i0, i1, j0, j1):
        out = C.c_double()
        self.call("pycs_field_max", fid, int(i0), int(i1), int(j0), int(j1), C.byref(out))
        return out.value

    def sm_count(self):
        n = C.c_int32()
        name = C.create_string_buffer(128)
        self.call("pycs_device_info", C.byref(n), name, 128)
        return n.value, name.value.decode()

    def fused_supported(self):
        y = C.c_int32()
        self.call("pycs_fused_supported", C.byref(y))
        return bool(y.value)

    # -- multi-GPU (one process per GPU; see include/pycs_b200.h) ------------
    def mgpu_setup(self, rank, world, all_gather_bytes):
        """Shard the fused step over `world` ranks.  `all_gather_bytes(b)` must return the list of
        every rank's byte string in rank order (e.g. torch.distributed.all_gather_object)."""
        buf = C.create_string_buffer(192)
        self.call("pycs_mgpu_init", rank, world, buf)
        parts = all_gather_bytes(buf.raw)
        if len(parts) != world or any(len(p) != 192 for p in parts):
            raise PycsError("mgpu_setup: expected %d handle blocks of 192 bytes" % world)
        self.call("pycs_mgpu_connect", b"".join(parts))
        return self.row_range()

    def row_range(self):
        a, b = C.c_int32(), C.c_int32()
        self.call("pycs_mgpu_row_range", C.byref(a), C.byref(b))
        return a.value, b.value

    def step_kernel_name(self):
        buf = C.create_string_buffer(256)
        self.call("pycs_step_kernel_name", buf, 256)
        return buf.value.decode()

    def launches(self):
        n = C.c_int64()
        self.call("pycs_launch_count", C.byref(n))
        return n.value


def mgpu_plan(N, world, rank, kmin_east, degree=3):
    """Host-only view of the decomposition: (row_lo, row_hi, [(peer, panel, i0, i1, j0, j1), ...]) -- the
    rectangles `rank` stores into its peers after every step.  kmin_east: the (4, P) stencil-start table
    of lagrange_poly_ghostcell_pc (simulation.stencil_ghost_pc[0][0])."""
    lib = load_library()
    km = np.ascontiguousarray(kmin_east, dtype=np.int32)
    a, b, n = C.c_int32(), C.c_int32(), C.c_int32()
    cap = 4096
    rects = (C.c_int32 * (6 * cap))()
    rc = lib.pycs_mgpu_plan(N, world, rank, degree, km.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(a), C.byref(b),
                            rects, cap, C.byref(n))
    if rc != 0:
        raise PycsError(lib.pycs_last_error().decode())
    if n.value > cap:
        raise PycsError("mgpu_plan: %d rectangles" % n.value)
    return a.value, b.value, [tuple(rects[6 * k:6 * k + 6]) for k in range(n.value)]


class DeviceArray:
    """Lazy numpy view of a device field (reference layout)."""

    def __init__(self, dev, fid):
        self.dev, self.fid = dev, fid

    @property
    def shape(self):
        return self.dev.shape(self.fid)

    def numpy(self):
        return self.dev.download(self.fid)

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, idx):
        return self.numpy()[idx]

    def __setitem__(self, idx, value):
        if idx == Ellipsis or idx == (slice(None),) * 3:
            full = np.empty(self.shape)
            full[...] = value
        else:
            full = self.numpy()
            full[idx] = value
        self.dev.upload(self.fid, full)

    def __repr__(self):
        return "DeviceArray(field=%d, shape=%s)" % (self.fid, self.shape)


class staged:
    """Context manager: make sure `arr` lives on the device for an operator.

    DeviceArray -> its own field; numpy array -> uploaded into a scratch field
    and, on exit, downloaded back into the caller's array (in-place semantics).
    """

    def __init__(self, dev, arr, scratch, writeback=True):
        self.dev, self.arr, self.scratch, self.writeback = dev, arr, scratch, writeback

    def __enter__(self):
        if isinstance(self.arr, DeviceArray):
            self.fid = self.arr.fid
            self.host = None
        else:
            self.fid = self.scratch
            self.host = self.arr
            self.dev.upload(self.fid, self.host)
        return self.fid

    def __exit__(self, *exc):
        if self.host is not None and self.writeback and exc[0] is None:
            self.dev.download(self.fid, self.host)
        return False
