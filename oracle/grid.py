"""Oracle: lean equiangular cubed-sphere grid (the *inputs* of the hot path).

Restates only the part of `cubed_sphere.__init__` the advection path reads
(SURVEY.md s8b): point coordinates at pc/pu/pv, sqrt(g) at pc/pu/pv and the
lat-lon <-> contravariant conversion coefficients.  Reference:
  src/cs_datastruct.py:197-237  (extents, ghost width 4+4, index ranges)
  src/cs_datastruct.py:240-324  (pc / pu / pv point generation)
  src/cs_datastruct.py:407-446  (sqrt(g) from the tangent vectors, panel 0 copied)
  src/cs_datastruct.py:448-493  (conversion coefficients and determinant)
  src/cs_transform.py:41-96     (equiangular gnomonic map)
  src/cs_transform.py:245-397   (tangent vectors)
  src/sphgeo.py:29-33, 76-92    (cart2sph, lat-lon unit tangent vectors)
The operation order of each formula is kept so that the arrays agree with the
reference bit for bit (checked in tests/golden/make_golden.py).
"""
import math

import numpy as np

NPANELS = 6  # src/constants.py:27


class Points:
    """Container with X, Y, Z, lon, lat of shape (n, m, 6) (src/sphgeo.py:138-147)."""
    __slots__ = ("X", "Y", "Z", "lon", "lat")


# Per panel: which of (invD, XoD, YoD) and sign goes to X, Y, Z
# (src/cs_transform.py:64-92).  0 = invD, 1 = XoD, 2 = YoD.
_POS = (
    ((0, 1), (1, 1), (2, 1)),
    ((1, -1), (0, 1), (2, 1)),
    ((0, -1), (1, -1), (2, 1)),
    ((1, 1), (0, -1), (2, 1)),
    ((2, -1), (1, 1), (0, 1)),
    ((2, 1), (1, 1), (0, -1)),
)
# Tangent vector in x: components of (ax, a2+y2, xy)*invr (src/cs_transform.py:266-295)
_TGX = (
    ((0, -1), (1, 1), (2, -1)),
    ((1, -1), (0, -1), (2, -1)),
    ((0, 1), (1, -1), (2, -1)),
    ((1, 1), (0, 1), (2, -1)),
    ((2, 1), (1, 1), (0, -1)),
    ((2, -1), (1, 1), (0, 1)),
)
# Tangent vector in y: components of (ay, xy, a2+x2)*invr (src/cs_transform.py:325-354)
_TGY = (
    ((0, -1), (1, -1), (2, 1)),
    ((1, 1), (0, -1), (2, 1)),
    ((0, 1), (1, 1), (2, 1)),
    ((1, -1), (0, 1), (2, 1)),
    ((2, -1), (1, -1), (0, -1)),
    ((2, 1), (1, -1), (0, 1)),
)


def _assemble(table, comps, shape):
    out = [np.empty(shape + (NPANELS,)) for _ in range(3)]
    for p in range(NPANELS):
        for axis in range(3):
            k, s = table[p][axis]
            out[axis][:, :, p] = comps[k] if s > 0 else -comps[k]
    return out


def gnomonic_points(x1d, y1d):
    """Equiangular map of the tensor grid x1d x y1d (src/cs_transform.py:41-96)."""
    x, y = np.meshgrid(x1d, y1d, indexing="ij")
    tanx = np.tan(x)
    tany = np.tan(y)
    D2 = 1.0 + tanx**2 + tany**2
    invD = 1.0 / np.sqrt(D2)
    comps = (invD, invD * tanx, invD * tany)
    pts = Points()
    pts.X, pts.Y, pts.Z = _assemble(_POS, comps, x.shape)
    pts.lat = np.arctan2(pts.Z, np.hypot(pts.X, pts.Y))  # src/sphgeo.py:29-33
    pts.lon = np.arctan2(pts.Y, pts.X)
    return pts


def tangent_vectors(x1d, y1d, R=1.0):
    """(ex, ey) Cartesian components, each a 3-list of (n,m,6) arrays.

    src/cs_transform.py:365-397 calling :245-354 with (a tan x, a tan y).
    """
    a = R / np.sqrt(3.0)
    xe, ye = np.meshgrid(a * np.tan(x1d), a * np.tan(y1d), indexing="ij")
    r2 = a**2 + xe**2 + ye**2
    invr = R / np.sqrt(r2) ** 3
    a2 = a * a
    xy = xe * ye
    ex = _assemble(_TGX, ((a * xe) * invr, (a2 + ye * ye) * invr, xy * invr), xe.shape)
    ey = _assemble(_TGY, ((a * ye) * invr, xy * invr, (a2 + xe * xe) * invr), xe.shape)
    x, y = np.meshgrid(x1d, y1d, indexing="ij")
    cos2x = np.cos(x) * np.cos(x)
    cos2y = np.cos(y) * np.cos(y)
    for c in range(3):
        for p in range(NPANELS):
            ex[c][:, :, p] = a * ex[c][:, :, p] / cos2x
            ey[c][:, :, p] = a * ey[c][:, :, p] / cos2y
    return ex, ey


class LeanGrid:
    """Duck-typed stand-in for `cubed_sphere` holding only what the path reads."""

    @classmethod
    def centres_only(cls, N):
        """Grid with only pc coordinates -- enough for the Lagrange tables and
        the halo fill at sizes where the full grid is not needed (N=3072)."""
        return cls(N, positions=("pc",), with_conversion=False, with_metric=False)

    def __init__(self, N, with_conversion=True, positions=("pc", "pu", "pv"), with_metric=True):
        self.N = N
        self.R = 1.0
        self.projection = "gnomonic_equiangular"
        a = math.pi * 0.5 * 0.5  # pio4, src/constants.py:20-22
        self.a = a
        x_min, x_max, y_min, y_max = -a, a, -a, a
        dx = (x_max - x_min) / N
        dy = (y_max - y_min) / N
        self.dx, self.dy = dx, dy
        ngl = ngr = 4                                   # src/cs_datastruct.py:225-231
        ng = ngl + ngr
        self.ngl, self.ngr, self.ng = ngl, ngr, ng
        self.i0 = self.j0 = ngl
        self.iend = self.jend = ngl + N
        P = N + ng

        x_e = np.linspace(x_min - ngl * dx, x_max + ngr * dx, N + 1 + ng)          # edges
        y_e = np.linspace(y_min - ngl * dx, y_max + ngr * dy, N + 1 + ng)
        x_c = np.linspace(x_min + dx / 2.0 - ngl * dx, x_max - dx / 2.0 + ngr * dx, P)  # centres
        y_c = np.linspace(y_min + dy / 2.0 - ngl * dy, y_max - dy / 2.0 + ngr * dy, P)
        self._axes = {"pc": (x_c, y_c), "pu": (x_e, y_c), "pv": (x_c, y_e)}

        for pos, (xs, ys) in self._axes.items():
            if pos not in positions:
                continue
            pts = gnomonic_points(xs, ys)
            setattr(self, pos, pts)
            if not with_metric:
                continue
            ex, ey = tangent_vectors(xs, ys, self.R)
            # sqrt(g) on panel 0, copied to all panels (src/cs_datastruct.py:407-446)
            e0 = [c[:, :, 0] for c in ex]
            f0 = [c[:, :, 0] for c in ey]
            g = -(e0[0] * f0[0] + e0[1] * f0[1] + e0[2] * f0[2]) ** 2 \
                + (e0[0] ** 2 + e0[1] ** 2 + e0[2] ** 2) * (f0[0] ** 2 + f0[1] ** 2 + f0[2] ** 2)
            g = np.sqrt(g)
            setattr(self, "metric_tensor_" + pos, np.repeat(g[:, :, None], NPANELS, axis=2))
            if not with_conversion:
                continue
            # lat-lon unit vectors (src/sphgeo.py:76-92) and the 2x2 conversion
            # matrix entries (src/cs_datastruct.py:456-493)
            sl, cl = np.sin(pts.lon), np.cos(pts.lon)
            st, ct = np.sin(pts.lat), np.cos(pts.lat)
            elon = (-sl, cl, np.zeros_like(sl))
            elat = (-st * cl, -st * sl, ct)
            dot = lambda u, v: u[0] * v[0] + u[1] * v[1] + u[2] * v[2]
            exlon, exlat = dot(ex, elon), dot(ex, elat)
            eylon, eylat = dot(ey, elon), dot(ey, elat)
            setattr(self, "prod_ex_elon_" + pos, exlon)
            setattr(self, "prod_ex_elat_" + pos, exlat)
            setattr(self, "prod_ey_elon_" + pos, eylon)
            setattr(self, "prod_ey_elat_" + pos, eylat)
            setattr(self, "determinant_ll2contra_" + pos, exlon * eylat - eylon * exlat)
