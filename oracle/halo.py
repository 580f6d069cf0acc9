"""Oracle: panel-edge halo gather, Lagrange (duo-grid) ghost fill, copy fill.

Restates, table-driven instead of panel by panel:
  src/halo_data.py:15-185    get_halo_data_interpolation      -> gather()
  src/halo_data.py:191-400   ..._NS / ..._WE (x/y field swap)  -> gather(Qx, Qy)
  src/lagrange.py:18-163     lagrange_poly_ghostcell_pc        -> lagrange_tables()
  src/interpolation.py:154-314  ghost_cell_pc_lagrange_interpolation -> dg_fill()
  src/interpolation.py:320-340  ghost_cells_adjacent_panels    -> copy_fill()
"""
from math import ceil

import numpy as np

E, W, N_, S = 0, 1, 2, 3

# (neighbour panel, strip, ops) for every panel and side, from
# src/halo_data.py:33-182.  strip: 'ilo' = Q[i0:i0+4,:], 'ihi' = Q[iend-4:iend,:],
# 'jlo' = Q[:,j0:j0+4], 'jhi' = Q[:,jend-4:jend];  ops applied left to right:
# 'T' transpose, 'F0'/'F1' flip along axis 0/1.
HALO_TABLE = {
    0: {E: (1, "ilo", ()), W: (3, "ihi", ()), N_: (4, "jlo", ()), S: (5, "jhi", ())},
    1: {E: (2, "ilo", ()), W: (0, "ihi", ()),
        N_: (4, "ihi", ("T", "F1")), S: (5, "ihi", ("T", "F0"))},
    2: {E: (3, "ilo", ()), W: (1, "ihi", ()),
        N_: (4, "jhi", ("F0", "F1")), S: (5, "jlo", ("F1", "F0"))},
    3: {E: (0, "ilo", ()), W: (2, "ihi", ()),
        N_: (4, "ilo", ("T", "F0")), S: (5, "ilo", ("T", "F1"))},
    4: {E: (1, "jhi", ("F1", "T")), W: (3, "jhi", ("T", "F1")),
        N_: (2, "jhi", ("F0", "F1")), S: (0, "jhi", ())},
    5: {E: (1, "jlo", ("T", "F1")), W: (3, "jlo", ("T", "F0")),
        N_: (0, "jlo", ()), S: (2, "jlo", ("F0", "F1"))},
}


def _strip(A, kind, g):
    if kind == "ilo":
        return A[g.i0:g.i0 + g.ngr, :]
    if kind == "ihi":
        return A[g.iend - g.ngl:g.iend, :]
    if kind == "jlo":
        return A[:, g.j0:g.j0 + g.ngr]
    return A[:, g.jend - g.ngl:g.jend]


def _apply(A, ops):
    for op in ops:
        A = A.T if op == "T" else np.flip(A, axis=0 if op == "F0" else 1)
    return A


def gather(Qx, Qy, g):
    """Neighbour strips re-oriented into each panel's frame: (E, W, N, S).

    With Qx is Qy this is get_halo_data_interpolation (src/halo_data.py:15);
    otherwise the strip comes from Qy when the neighbour's axes are not rotated
    w.r.t. an E/W exchange (or from Qx for N/S) and from the other field when a
    transpose is involved -- exactly the choices at src/halo_data.py:213-297
    (NS) and :327-396 (WE).
    """
    P = g.N + g.ng
    out = [np.zeros((g.ngl, P, 6)), np.zeros((g.ngl, P, 6)),
           np.zeros((P, g.ngl, 6)), np.zeros((P, g.ngl, 6))]
    for p in range(6):
        for side in (E, W, N_, S):
            nb, kind, ops = HALO_TABLE[p][side]
            rotated = "T" in ops
            if side in (E, W):
                src = Qx if rotated else Qy
            else:
                src = Qy if rotated else Qx
            out[side][:, :, p] = _apply(_strip(src[:, :, nb], kind, g), ops)
    return out


def index_maps(g):
    """Integer source maps of gather(): for side s, (panel, i, j) of every entry.

    Obtained by pushing index arrays through the same strip/transpose/flip
    table, so they are exact by construction.  Returns [(nb, I, J)] per side
    with nb, I, J of the halo-array shape (.., .., 6).
    """
    P = g.N + g.ng
    ii, jj = np.meshgrid(np.arange(P), np.arange(P), indexing="ij")
    I6 = np.repeat(ii[:, :, None], 6, 2).astype(float)
    J6 = np.repeat(jj[:, :, None], 6, 2).astype(float)
    P6 = np.broadcast_to(np.arange(6.0), (P, P, 6)).copy()
    gi, gj, gp = gather(I6, I6, g), gather(J6, J6, g), gather(P6, P6, g)
    return [(gp[s].astype(np.int64), gi[s].astype(np.int64), gj[s].astype(np.int64))
            for s in range(4)]


# ----------------------------------------------------------------------------
def lagrange_basis(x, nodes, degree, j):
    """src/lagrange.py:18-23, vectorised over leading axes (same op order)."""
    L = np.ones_like(x)
    for i in range(degree + 1):
        if i != j:
            L = L * (x - nodes[..., i]) / (nodes[..., j] - nodes[..., i])
    return L


def lagrange_tables(g, degree=3):
    """Stencil start/end and weights for the ghost-cell centres.

    Returns (stencil, poly) shaped like simulation.stencil_ghost_pc /
    lagrange_poly_ghost_pc: stencil = [[Kmin_E,W,N,S],[Kmax_E,W,N,S]],
    poly = [E, W, N, S].  src/lagrange.py:28-163.
    """
    if g.projection != "gnomonic_equiangular":
        raise ValueError("ET-DG needs the equiangular grid (src/lagrange.py:40-45)")
    i0, iend, ngl, ngr = g.i0, g.iend, g.ngl, g.ngr
    P = g.N + g.ng
    order = degree + 1
    east = 1
    # inverse equiangular map onto panel 1 (src/cs_transform.py:105-107):
    inv_y = lambda X, Y, Z: np.arctan(Z / Y)
    y_ghost = inv_y(g.pc.X[iend:iend + ngr, :, 0], g.pc.Y[iend:iend + ngr, :, 0],
                    g.pc.Z[iend:iend + ngr, :, 0])                     # :57-60
    y = inv_y(g.pc.X[i0:i0 + ngr, :, east], g.pc.Y[i0:i0 + ngr, :, east],
              g.pc.Z[i0:i0 + ngr, :, east])                            # :63-66

    K = (y_ghost - y[:, 0:1]) / g.dy                                   # :82
    Kmax = K + ceil(order / 2)                                         # :83
    Kmin = Kmax - order + 1                                            # :84
    idx = np.arange(P)[None, :]
    inner = (idx >= i0) & (idx < iend)
    hi = idx >= iend
    lo = idx < i0
    # shifts, float comparisons as in :87-103
    m = inner & (Kmax >= iend)
    m2 = inner & ~(Kmax >= iend) & (Kmin < i0)
    Kmax = np.where(m, iend - 1.0, Kmax); Kmin = np.where(m, Kmax - order + 1, Kmin)
    Kmin = np.where(m2, float(i0), Kmin); Kmax = np.where(m2, Kmin + order - 1, Kmax)
    m = hi & (Kmax >= P)
    Kmax = np.where(m, P - 1.0, Kmax); Kmin = np.where(m, Kmax - order + 1, Kmin)
    m = lo & (Kmin < 0)
    Kmin = np.where(m, 0.0, Kmin); Kmax = np.where(m, Kmin + order - 1, Kmax)
    Kmin = Kmin.astype(int)                                            # :105-107
    Kmax = Kmax.astype(int)

    # consistency checks of :109-126 (interior columns only)
    ki, xi = Kmin[:, i0:iend], Kmax[:, i0:iend]
    if np.any(xi - ki != degree) or np.any(ki < i0) or np.any(xi > iend):
        raise RuntimeError("lagrange_tables: bad stencil")
    if order > 1:
        rows = np.arange(ngl)[:, None]
        yg = y_ghost[:, i0:iend]
        if np.any(yg < y[rows, ki]) or np.any(yg > y[rows, xi]):
            raise RuntimeError("lagrange_tables: ghost point not bracketed")

    rows = np.arange(ngl)[:, None, None]
    nodes = y[rows, Kmin[:, :, None] + np.arange(order)[None, None, :]]   # :131-133
    poly = np.zeros((ngr, P, order))
    for l in range(order):                                               # :136-139
        poly[:, :, l] = lagrange_basis(y_ghost, nodes, degree, l)

    poly_E = poly                                                        # :141-152
    poly_W = np.flip(poly, axis=0)
    poly_N = np.transpose(poly_E, (1, 0, 2))
    poly_S = np.flip(poly_N, axis=1)
    kmin = [Kmin, np.flip(Kmin, 0), Kmin.T, np.flip(Kmin.T, 1)]
    kmax = [Kmax, np.flip(Kmax, 0), Kmax.T, np.flip(Kmax.T, 1)]
    return [kmin, kmax], [poly_E, poly_W, poly_N, poly_S]


def _interp_rows(strip, kmin, poly):
    """ghost[g,k] = sum_l poly[g,k,l] * strip[g, kmin[g,k]+l]  (E/W orientation).

    src/interpolation.py:203-209: gather, multiply, np.sum over the last axis
    (numpy adds 4 terms left to right).
    """
    order = poly.shape[2]
    rows = np.arange(strip.shape[0])[:, None, None]
    cols = kmin[:, :, None] + np.arange(order)[None, None, :]
    return np.sum(strip[rows, cols] * poly, axis=2)


def dg_fill(Q, g, tables):
    """In-place duo-grid ghost fill of Q (src/interpolation.py:154-314)."""
    (kmin, _kmax), poly = tables
    i0, iend, j0, jend, ngl, ngr = g.i0, g.iend, g.j0, g.jend, g.ngl, g.ngr
    # phase 1: the four edge strips, interior extent only (:200-248)
    hE, hW, hN, hS = gather(Q, Q, g)
    for p in range(6):
        Q[iend:iend + ngr, j0:jend, p] = _interp_rows(hE[:, :, p], kmin[E], poly[E])[:, j0:jend]
        Q[0:i0, j0:jend, p] = _interp_rows(hW[:, :, p], kmin[W], poly[W])[:, j0:jend]
        # N/S: same operation on the transposed strip
        gn = _interp_rows(hN[:, :, p].T, kmin[N_].T, np.transpose(poly[N_], (1, 0, 2))).T
        Q[i0:iend, jend:jend + ngr, p] = gn[i0:iend, :]
        gs = _interp_rows(hS[:, :, p].T, kmin[S].T, np.transpose(poly[S], (1, 0, 2))).T
        Q[i0:iend, 0:j0, p] = gs[i0:iend, :]
    # phase 2: corners from the re-gathered E/W strips (:250-314)
    hE, hW, _, _ = gather(Q, Q, g)
    for p in range(6):
        e = _interp_rows(hE[:, :, p], kmin[E], poly[E])
        w = _interp_rows(hW[:, :, p], kmin[W], poly[W])
        Q[iend:iend + ngr, jend:jend + ngr, p] = e[:, jend:jend + ngr]
        Q[iend:iend + ngr, 0:j0, p] = e[:, 0:j0]
        Q[0:i0, jend:jend + ngr, p] = w[:, jend:jend + ngr]
        Q[0:i0, 0:j0, p] = w[:, 0:j0]


def copy_fill(Qx, Qy, g):
    """ET-S72 / ET-PL07 ghost fill (src/interpolation.py:320-340)."""
    hE, hW, hN, hS = gather(Qx, Qy, g)
    Qy[0:g.i0, :, :] = hW
    Qy[g.iend:, :, :] = hE
    Qx[:, 0:g.j0, :] = hS
    Qx[:, g.jend:, :] = hN
