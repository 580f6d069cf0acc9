"""Advection set-up: parameters -> scheme names, state, analytic fields.

Mirror of src/advection_ic.py.  `adv_simulation_par` keeps the reference's
constructor signature and attribute names; its arrays (Q, gQ, div, cx, cy) are
`DeviceArray`s because the device owns the state.  The analytic fields are
host numpy (used once for the initial condition, the initial wind and the
error norms); the per-step wind refresh of update_adv runs on the device.
"""
import numpy as np

from .constants import pi, deg2rad, nbfaces
from .sphgeo import sph2cart
from .device import Device, F, PycsError

_RECON = {1: 'PPM-0', 2: 'PPM-CW84', 3: 'PPM-PL07', 4: 'PPM-L04'}
_DP = {1: 'RK1', 2: 'RK2'}
_SPLIT = {1: 'SP-AVLT', 2: 'SP-L04', 3: 'SP-PL07'}
_ET = {1: 'ET-S72', 2: 'ET-PL07', 3: 'ET-DG'}
_MT = {1: 'MT-0', 2: 'MT-PL07'}
_MF = {1: 'MF-0', 2: 'MF-AF', 3: 'MF-PR'}

_GEOM = (
    ("SQRTG_PC", "metric_tensor_pc"), ("SQRTG_PU", "metric_tensor_pu"), ("SQRTG_PV", "metric_tensor_pv"),
    ("PC_EXLON", "prod_ex_elon_pc"), ("PC_EXLAT", "prod_ex_elat_pc"), ("PC_EYLON", "prod_ey_elon_pc"),
    ("PC_EYLAT", "prod_ey_elat_pc"), ("PC_DET", "determinant_ll2contra_pc"),
    ("PU_EXLON", "prod_ex_elon_pu"), ("PU_EXLAT", "prod_ex_elat_pu"), ("PU_EYLON", "prod_ey_elon_pu"),
    ("PU_EYLAT", "prod_ey_elat_pu"), ("PU_DET", "determinant_ll2contra_pu"),
    ("PV_EXLON", "prod_ex_elon_pv"), ("PV_EXLAT", "prod_ex_elat_pv"), ("PV_EYLON", "prod_ey_elon_pv"),
    ("PV_EYLAT", "prod_ey_elat_pv"), ("PV_DET", "determinant_ll2contra_pv"),
)


def _fail(msg):
    """The reference prints and exit()s on bad parameters (src/advection_ic.py:61-63 ...)."""
    print(msg)
    raise SystemExit(1)


class adv_simulation_par:
    def __init__(self, cs_grid, dt, Tf, ic, vf, tc, recon, dp, opsplit, et, mt, mf, device=0):
        self.ic, self.vf, self.tc = ic, vf, tc
        self.dt, self.dto2, self.twodt = dt, dt * 0.5, dt * 2.0
        self.recon, self.dp, self.opsplit = recon, dp, opsplit
        self.Tf = Tf
        self.degree = 3                                    # src/advection_ic.py:50
        if ic not in (1, 2, 3, 4):
            _fail("Error in adv_simulation_par - invalid initial condition")
        if vf not in (1, 2, 3, 4):
            _fail("Error in adv_simulation_par - invalid vector field")
        if recon not in _RECON:
            _fail("Error in adv_simulation_par - invalid reconstruction method")
        if dp not in _DP:
            _fail("Error in simulation_adv_par - invalid departure point scheme")
        if opsplit not in _SPLIT:
            _fail("Error in adv_simulation_par - invalid operator splitting method")
        if et not in _ET:
            _fail('ERROR in adv_simulation_par: invalid ET')
        if mt not in _MT:
            _fail('ERROR in adv_simulation_par: invalid MT')
        if mf not in _MF:
            _fail('ERROR in adv_simulation_par: invalid MF')
        if tc not in (1, 2):
            _fail("Error in adv_simulation_par- invalid test case")
        self.recon_name, self.dp_name, self.opsplit_name = _RECON[recon], _DP[dp], _SPLIT[opsplit]
        self.et_name, self.mt_name, self.mf_name = _ET[et], _MT[mt], _MF[mf]
        self.edge_treatment, self.metric_tensor, self.mass_fixer = et, mt, mf
        self.title = '2D Advection ' if tc == 1 else '2D advection errors '
        # the reference raises UnboundLocalError inside divergence() for the other
        # combinations (src/discrete_operators.py:72-81); refuse them up front
        if not ((opsplit in (1, 2) and mt == 1) or (opsplit == 3 and mt == 2)):
            raise PycsError("invalid scheme: %s needs %s" % (
                self.opsplit_name, 'MT-0' if opsplit in (1, 2) else 'MT-PL07'))
        if et == 3 and cs_grid.projection != "gnomonic_equiangular":
            raise PycsError("ET-DG needs the equiangular grid (src/lagrange.py:40-45)")

        self.dev = Device(cs_grid.N, cs_grid.dx, cs_grid.dy, dt, recon, dp, opsplit, et, mt, mf, vf, ic, device)
        cs_grid.dev = self.dev
        if getattr(cs_grid, "lean", False):
            # geometry generated on the device from the 1-D coordinates (csrc/grid.cu)
            import ctypes as C
            xc = np.ascontiguousarray(cs_grid.x_centres, dtype=np.float64)
            xe = np.ascontiguousarray(cs_grid.x_edges, dtype=np.float64)
            dp = C.POINTER(C.c_double)
            self.dev.call("pycs_generate_geometry", xc.ctypes.data_as(dp), xe.ctypes.data_as(dp))
        else:
            for fname, attr in _GEOM:
                self.dev.upload(F[fname], getattr(cs_grid, attr))
            for pos in ("pc", "pu", "pv"):
                pts = getattr(cs_grid, pos)
                self.dev.upload(F[pos.upper() + "_LON"], pts.lon)
                self.dev.upload(F[pos.upper() + "_LAT"], pts.lat)

        self.px = None
        self.py = None
        self.Q = self.dev.array(F["Q"])
        self.gQ = self.dev.array(F["GQ"])
        self.div = self.dev.array(F["DIV"])
        from .cs_datastruct import velocity
        self.U_pu = velocity(cs_grid, 'pu', self)
        self.U_pv = velocity(cs_grid, 'pv', self)
        self.U_pc = velocity(cs_grid, 'pc', self)
        self.cx = self.dev.array(F["CX"])
        self.cy = self.dev.array(F["CY"])
        self.CFL = 0.0
        self.total_mass0 = self.total_mass = self.mass_change = 0.0
        self.error_linf = self.error_l1 = self.error_l2 = None
        self.lagrange_poly_ghost_pc = self.stencil_ghost_pc = None
        # run the fused step kernel inside adv_sphere when the scheme has one
        self.fused = True


def q0_adv(lon, lat, simulation):
    return qexact_adv(lon, lat, 0, simulation)


def qexact_adv(lon, lat, t, simulation):
    """src/advection_ic.py:215-281."""
    ic, vf = simulation.ic, simulation.vf
    if ic == 1:
        return np.ones(np.shape(lon))
    if ic == 2:
        X, Y, Z = sph2cart(lon, lat)
        if vf == 1:
            alpha = -45.0 * deg2rad
            u0 = 2.0 * pi / 5.0
            wt = (-u0) * t
            cosa, sina = np.cos(alpha), np.sin(alpha)
            cos2a, sin2a = cosa * cosa, sina * sina
            coswt, sinwt = np.cos(wt), np.sin(wt)
            rotX = (coswt * cos2a + sin2a) * X - sinwt * cosa * Y + (coswt * cosa * sina - cosa * sina) * Z
            rotY = sinwt * cosa * X + coswt * Y + sina * sinwt * Z
            rotZ = (coswt * sina * cosa - sina * cosa) * X - sinwt * sina * Y + (coswt * sin2a + cos2a) * Z
            X0, Y0, Z0 = sph2cart(np.pi / 4.0, np.pi / 6.0)
            return np.exp(-10.0 * ((rotX - X0) ** 2 + (rotY - Y0) ** 2 + (rotZ - Z0) ** 2))
        X0, Y0, Z0 = sph2cart(0.0, 0.0)
        return np.exp(-10.0 * ((X - X0) ** 2 + (Y - Y0) ** 2 + (Z - Z0) ** 2))
    if ic == 3:
        X, Y, Z = sph2cart(lon, lat)
        (lon1, lat1), (lon2, lat2) = ((0, pi / 3.0), (0, -pi / 3.0)) if vf == 1 else ((-pi / 6.0, 0), (pi / 6.0, 0))
        X1, Y1, Z1 = sph2cart(lon1, lat1)
        X2, Y2, Z2 = sph2cart(lon2, lat2)
        b0 = 5.0
        return np.exp(-b0 * ((X - X1) ** 2 + (Y - Y1) ** 2 + (Z - Z1) ** 2)) + \
            np.exp(-b0 * ((X - X2) ** 2 + (Y - Y2) ** 2 + (Z - Z2) ** 2))
    if ic == 4:
        alpha = -45.0 * deg2rad
        f = (-np.cos(lon) * np.cos(lat) * np.sin(alpha) + np.sin(lat) * np.cos(alpha))
        return 1.0 - f * f
    print('Invalid initial condition.\n')
    raise SystemExit(1)


def velocity_adv(lon, lat, t, simulation):
    """Host evaluation of the analytic winds (src/advection_ic.py:287-313); the
    device evaluates the same expressions in update_adv."""
    vf = simulation.vf
    cos, sin = np.cos, np.sin
    if vf == 1:
        alpha = -45.0 * deg2rad
        u0 = 2.0 * pi / 5.0
        return u0 * (cos(lat) * cos(alpha) + sin(lat) * cos(lon) * sin(alpha)), -u0 * sin(lon) * sin(alpha)
    if vf == 2:
        T, k = 5.0, 2.0
        lonp = lon - 2 * pi * t / T
        return (k * (sin((lonp + pi)) ** 2) * (sin(2. * lat)) * (cos(pi * t / T)) + 2. * pi * cos(lat) / T,
                k * (sin(2 * (lonp + pi))) * (cos(lat)) * (cos(pi * t / T)))
    if vf == 3:
        T, k = 5.0, 1.0
        return (-k * (sin((lon + pi) / 2.0) ** 2) * (sin(2.0 * lat)) * (cos(lat) ** 2) * (cos(pi * t / T)),
                (k / 2.0) * (sin((lon + pi))) * (cos(lat) ** 3) * (cos(pi * t / T)))
    m = n = 1
    return (-m * (sin(lon) * sin(m * lon) * cos(n * lat) ** 3),
            -4 * n * (cos(n * lat) ** 3) * sin(n * lat) * cos(m * lon) * sin(lon))


def div_exact(lon, lat, simulation):
    """src/advection_ic.py:318-329."""
    if simulation.vf <= 2:
        return np.zeros(np.shape(lon))
    m = n = 1
    return (-np.cos(lon) * np.sin(m * lon) * m * np.cos(n * lat) ** 4 / np.cos(lat) -
            np.sin(lon) * np.cos(m * lon) * m ** 2 * np.cos(n * lat) ** 4 / np.cos(lat) +
            12.0 * np.sin(lon) * np.cos(m * lon) * np.cos(n * lat) ** 2 * np.sin(n * lat) ** 2 * n ** 2 * np.cos(lat) -
            4.0 * np.sin(lon) * np.cos(m * lon) * np.cos(n * lat) ** 4 * n ** 2 * np.cos(lat) +
            4.0 * np.sin(lon) * np.cos(m * lon) * np.cos(n * lat) ** 3 * np.sin(n * lat) * n * np.sin(lat)) / np.cos(lat)
