"""Host precompute of the ghost-cell Lagrange tables (once per grid).

Mirror of `lagrange_poly_ghostcell_pc` (src/lagrange.py:28-163): stencil start /
end indices and weights for the ghost-cell centres of panel 0's east side,
expressed in panel 1's coordinates; the other sides are flips / transposes.
The float expressions that feed the integer truncation (`K`, the clamps) are
kept verbatim so that Kmin / Kmax are bit-exact; the tables are then uploaded
once through pycs_upload_lagrange.
"""
from math import ceil

import numpy as np


def lagrange_basis(x, nodes, N, j):
    """src/lagrange.py:18-23 (vectorised over leading axes)."""
    Lj = np.ones_like(np.asarray(x, dtype=float))
    for i in range(0, N + 1):
        if i != j:
            Lj = Lj * (x - nodes[..., i]) / (nodes[..., j] - nodes[..., i])
    return Lj


def ghost_tables(cs_grid, degree):
    """(Kmin_east, Kmax_east, weights_east): (4,P) int, (4,P) int, (4,P,degree+1)."""
    if cs_grid.projection != "gnomonic_equiangular":
        print('ERROR in lagrange_poly_ghostcells: grid is not gnomonic_equiangular.')
        raise SystemExit(1)
    N, ng, ngl, ngr = cs_grid.N, cs_grid.ng, cs_grid.ngl, cs_grid.ngr
    i0, iend = cs_grid.i0, cs_grid.iend
    P = N + ng
    order = degree + 1
    if getattr(cs_grid, "lean", False):
        # lean grid: the two four-row strips of cell centres from the 1-D coordinates, with the expressions
        # of cs_datastruct.cubed_sphere (src/cs_transform.py:41-96), hence the same bits as the full grid's
        def strip(rows, panel):
            x, y = np.meshgrid(cs_grid.x_centres[rows], cs_grid.x_centres, indexing="ij")
            tx, ty = np.tan(x), np.tan(y)
            invD = 1.0 / np.sqrt(1.0 + tx**2 + ty**2)
            v0 = (invD, invD * tx, invD * ty)
            return (v0[1], v0[2]) if panel == 0 else (v0[0], v0[2])     # (Y, Z): panel 1 has Y = invD
        Yg, Zg = strip(slice(iend, iend + ngr), 0)
        Y1, Z1 = strip(slice(i0, i0 + ngr), 1)
    else:
        pc = cs_grid.pc
        Yg, Zg = pc.Y[iend:iend + ngr, :, 0], pc.Z[iend:iend + ngr, :, 0]
        Y1, Z1 = pc.Y[i0:i0 + ngr, :, 1], pc.Z[i0:i0 + ngr, :, 1]
    # inverse equiangular map on panel 1: y = arctan(Z/Y) (src/cs_transform.py:105-107)
    y_ghost = np.arctan(Zg / Yg)
    y = np.arctan(Z1 / Y1)
    K = (y_ghost - y[:, 0:1]) / cs_grid.dy
    Kmax = K + ceil(order / 2)
    Kmin = Kmax - order + 1
    col = np.arange(P)[None, :]
    interior = (col >= i0) & (col < iend)
    over = interior & (Kmax >= iend)
    under = interior & ~over & (Kmin < i0)
    Kmax = np.where(over, iend - 1.0, Kmax)
    Kmin = np.where(over, Kmax - order + 1, Kmin)
    Kmin = np.where(under, float(i0), Kmin)
    Kmax = np.where(under, Kmin + order - 1, Kmax)
    top = (col >= iend) & (Kmax >= N + ng)
    Kmax = np.where(top, N + ng - 1.0, Kmax)
    Kmin = np.where(top, Kmax - order + 1, Kmin)
    bot = (col < i0) & (Kmin < 0)
    Kmin = np.where(bot, 0.0, Kmin)
    Kmax = np.where(bot, Kmin + order - 1, Kmax)
    Kmin, Kmax = Kmin.astype(int), Kmax.astype(int)
    ki, ka = Kmin[:, i0:iend], Kmax[:, i0:iend]
    rows = np.arange(ngl)[:, None]
    if np.any(ka - ki != degree):
        print('ERROR in lagrange_poly_ghostcells', degree)
        raise SystemExit(1)
    if np.any(ki < i0) or np.any(ka > iend):
        print('Error in lagrange_poly_ghostcells')
        raise SystemExit(1)
    if order > 1 and (np.any(y_ghost[:, i0:iend] < y[rows, ki]) or np.any(y_ghost[:, i0:iend] > y[rows, ka])):
        print('ERROR in lagrange_poly_ghostcells')
        raise SystemExit(1)
    nodes = y[np.arange(ngl)[:, None, None], Kmin[:, :, None] + np.arange(order)[None, None, :]]
    w = np.zeros((ngr, P, order))
    for l in range(order):
        w[:, :, l] = lagrange_basis(y_ghost, nodes, degree, l)
    return Kmin, Kmax, w


def lagrange_poly_ghostcell_pc(cs_grid, simulation):
    """Sets simulation.stencil_ghost_pc / lagrange_poly_ghost_pc (reference layout)
    and uploads the east tables to the device."""
    Kmin, Kmax, w = ghost_tables(cs_grid, simulation.degree)
    kmin = [Kmin, np.flip(Kmin, axis=0), Kmin.T, np.flip(Kmin.T, axis=1)]
    kmax = [Kmax, np.flip(Kmax, axis=0), Kmax.T, np.flip(Kmax.T, axis=1)]
    wn = np.transpose(w, (1, 0, 2))
    simulation.stencil_ghost_pc = [kmin, kmax]
    simulation.lagrange_poly_ghost_pc = [w, np.flip(w, axis=0), wn, np.flip(wn, axis=1)]
    dev = getattr(simulation, "dev", None)
    if dev is not None:
        upload_tables(dev, simulation.degree, Kmin, w)


def upload_tables(dev, degree, Kmin, w):
    import ctypes as C
    k32 = np.ascontiguousarray(Kmin, dtype=np.int32)
    wc = np.ascontiguousarray(w, dtype=np.float64)
    dev.call("pycs_upload_lagrange", degree, k32.ctypes.data_as(C.POINTER(C.c_int32)),
             wc.ctypes.data_as(C.POINTER(C.c_double)))
