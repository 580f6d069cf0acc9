"""Host grid object: the inputs the device path reads from `cs_grid`.

Mirror of `cubed_sphere` (src/cs_datastruct.py:24-541) restricted to the
attributes listed in SURVEY.md s8b, equiangular or equidistant gnomonic grid.
Grid generation is one-time numpy set-up (out of scope as GPU work); it is
written from the geometry rather than panel by panel: every panel's position
and tangent vectors are the panel-0 vectors under a signed axis permutation
R_p, which is exact in floating point, so the arrays equal the reference's.
"""
import numpy as np

from .constants import pi, pio2, pio4, nbfaces
from .sphgeo import point, sph2cart

# R_p as (source component, sign) for X, Y, Z: panel p position = R_p (v0, v1, v2)
# with (v0, v1, v2) the panel-0 position (src/cs_transform.py:64-92, :163-191).
_ROT = (
    ((0, 1), (1, 1), (2, 1)),
    ((1, -1), (0, 1), (2, 1)),
    ((0, -1), (1, -1), (2, 1)),
    ((1, 1), (0, -1), (2, 1)),
    ((2, -1), (1, 1), (0, 1)),
    ((2, 1), (1, 1), (0, -1)),
)


def _rotate(vec0, shape):
    out = [np.empty(shape + (nbfaces,)) for _ in range(3)]
    for p, rot in enumerate(_ROT):
        for c, (k, s) in enumerate(rot):
            out[c][:, :, p] = vec0[k] if s > 0 else -vec0[k]
    return out


class cubed_sphere:
    def __init__(self, N, transformation="gnomonic_equiangular", showonscreen=False, gridload=False,
                 centres_only=False, lean=False):
        """centres_only (not in the reference): build only the cell-centre coordinates (pc), which is
        all the Lagrange ghost-cell tables and the standalone halo fill read (src/lagrange.py:60-69);
        lets the interpolation.par path run at N = 3072 without the 50 GB full grid.

        lean (not in the reference; equiangular grid only): build nothing but the scalars and the two 1-D
        coordinate arrays; the fields the path reads (sqrt(g), conversion coefficients, lon / lat) are then
        generated on the device by adv_simulation_par (pycs_generate_geometry, csrc/grid.cu), and
        init_vars_adv evaluates the wind and the initial condition there too.  The host arrays
        (cs_grid.pc.lon, metric_tensor_pc ...) do not exist on a lean grid."""
        if transformation not in ("gnomonic_equiangular", "gnomonic_equidistant"):
            print("ERROR: invalid grid transformation.")
            raise SystemExit(1)
        self.R = 1.0
        self.projection = transformation
        self.N = N
        self.name = self.projection + "_cs_" + str(N)
        equiang = transformation == "gnomonic_equiangular"
        a = pio4 if equiang else self.R / np.sqrt(3.0)
        self.a = a
        dx = (a - (-a)) / N
        self.dx = self.dy = dx
        self.ngl = self.ngr = 4
        self.ng = 8
        self.i0 = self.j0 = 4
        self.iend = self.jend = 4 + N
        P = N + 8
        edges = np.linspace(-a - 4 * dx, a + 4 * dx, N + 1 + 8)
        cents = np.linspace(-a + dx / 2.0 - 4 * dx, a - dx / 2.0 + 4 * dx, P)
        half = self.R / np.sqrt(3.0)
        self.x_centres, self.x_edges = cents, edges
        self.lean = bool(lean)
        if lean:
            if not equiang:
                print("ERROR: the lean (device-generated) grid is equiangular only.")
                raise SystemExit(1)
            return
        positions = {"pc": (cents, cents), "pu": (edges, cents), "pv": (cents, edges)}
        if centres_only:
            positions = {"pc": (cents, cents)}
        for pos, (xs, ys) in positions.items():
            x, y = np.meshgrid(xs, ys, indexing="ij")
            shape = x.shape
            if equiang:
                tx, ty = np.tan(x), np.tan(y)
                invD = 1.0 / np.sqrt(1.0 + tx**2 + ty**2)
                v0 = (invD, invD * tx, invD * ty)
                X, Y = np.meshgrid(half * np.tan(xs), half * np.tan(ys), indexing="ij")
            else:
                invr = 1.0 / np.sqrt(half**2 + x**2 + y**2)
                v0 = (invr * half, invr * x, invr * y)
                X, Y = x, y
            pts = point()
            pts.X, pts.Y, pts.Z = _rotate(v0, shape)
            pts.lat = np.arctan2(pts.Z, np.hypot(pts.X, pts.Y))
            pts.lon = np.arctan2(pts.Y, pts.X)
            setattr(self, pos, pts)
            if centres_only:
                continue
            # panel-0 tangent vectors (src/cs_transform.py:245-354), chain rule for the
            # equiangular map (:365-397)
            invr3 = self.R / np.sqrt(half**2 + X**2 + Y**2) ** 3
            a2, xy = half * half, X * Y
            ex0 = [-(half * X) * invr3, (a2 + Y * Y) * invr3, -(xy * invr3)]
            ey0 = [-((half * Y) * invr3), -(xy * invr3), (a2 + X * X) * invr3]
            if equiang:
                c2x, c2y = np.cos(x) * np.cos(x), np.cos(y) * np.cos(y)
                ex0 = [half * c / c2x for c in ex0]
                ey0 = [half * c / c2y for c in ey0]
            g = -(ex0[0] * ey0[0] + ex0[1] * ey0[1] + ex0[2] * ey0[2]) ** 2 \
                + (ex0[0] ** 2 + ex0[1] ** 2 + ex0[2] ** 2) * (ey0[0] ** 2 + ey0[1] ** 2 + ey0[2] ** 2)
            setattr(self, "metric_tensor_" + pos, np.repeat(np.sqrt(g)[:, :, None], nbfaces, axis=2))
            ex, ey = _rotate(ex0, shape), _rotate(ey0, shape)
            sl, cl, st, ct = np.sin(pts.lon), np.cos(pts.lon), np.sin(pts.lat), np.cos(pts.lat)
            elon = (-sl, cl, 0.0 * sl)
            elat = (-st * cl, -st * sl, ct)
            dot = lambda u, v: u[0] * v[0] + u[1] * v[1] + u[2] * v[2]
            exlon, exlat, eylon, eylat = dot(ex, elon), dot(ex, elat), dot(ey, elon), dot(ey, elat)
            setattr(self, "prod_ex_elon_" + pos, exlon)
            setattr(self, "prod_ex_elat_" + pos, exlat)
            setattr(self, "prod_ey_elon_" + pos, eylon)
            setattr(self, "prod_ey_elat_" + pos, eylat)
            setattr(self, "determinant_ll2contra_" + pos, exlon * eylat - eylon * exlat)


class latlon_grid:
    """Regular lat-lon grid used for plots and the .npy error dumps (src/cs_datastruct.py:569-582);
    `ix, jy, mask` are filled by interpolation.ll2cs."""

    def __init__(self, Nlat, Nlon):
        self.Nlat, self.Nlon = Nlat, Nlon
        self.lat, self.lon = np.meshgrid(np.linspace(-pio2, pio2, Nlat), np.linspace(-pi, pi, Nlon))
        self.X, self.Y, self.Z = sph2cart(self.lon, self.lat)
        self.ix, self.jy, self.mask = [], [], []


class scalar_field:
    """src/cs_datastruct.py:546-564."""

    def __init__(self, grid, name, position):
        self.name, self.position = name, position
        self.N = grid.N + 1 if position == "vertex" else grid.N
        self.f = np.zeros((self.N, self.N, nbfaces))


class ppm_parabola:
    """Device-backed px / py (src/cs_datastruct.py:635-684)."""

    def __init__(self, cs_grid, simulation, direction):
        from .device import F
        self.recon_name = simulation.recon_name
        self.direction = direction
        pre = "PX_" if direction == "x" else "PY_"
        dev = simulation.dev
        for attr, f in (("q_L", "QL"), ("q_R", "QR"), ("dq", "DQ"), ("q6", "Q6"),
                        ("f_L", "FL"), ("f_R", "FR"), ("f_upw", "FUPW"), ("dF", "DF")):
            setattr(self, attr, dev.array(F[pre + f]))


class velocity:
    """Device-backed U_pu / U_pv / U_pc (src/cs_datastruct.py:701-735)."""

    def __init__(self, cs_grid, pos, simulation):
        from .device import F
        dev = simulation.dev
        self.pos = pos
        self._grid = cs_grid
        pre = {"pu": "PU_", "pv": "PV_", "pc": "PC_"}[pos]
        for attr in ("ulon", "vlat", "ucontra", "vcontra"):
            setattr(self, attr, dev.array(F[pre + attr.upper()]))
        if pos == "pu":
            self.ucontra_averaged = dev.array(F["PU_UAVG"])
            self.ucontra_old = dev.array(F["PU_UOLD"])
        elif pos == "pv":
            self.vcontra_averaged = dev.array(F["PV_VAVG"])
            self.vcontra_old = dev.array(F["PV_VOLD"])

    # upwind masks are the sign of the instantaneous wind (src/averaged_velocity.py:21-27)
    @property
    def upos(self):
        g = self._grid
        return np.asarray(self.ucontra)[g.i0:g.iend + 1, :, :] >= 0

    @property
    def uneg(self):
        return ~self.upos

    @property
    def vpos(self):
        g = self._grid
        return np.asarray(self.vcontra)[:, g.j0:g.jend + 1, :] >= 0

    @property
    def vneg(self):
        return ~self.vpos
