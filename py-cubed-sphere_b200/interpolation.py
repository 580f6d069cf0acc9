"""Ghost-cell fills (hot part of src/interpolation.py) and the host-side lat-lon regridding
used by the output routines (src/interpolation.py:38-149)."""
import numpy as np

from .constants import nbfaces
from .cs_transform import inverse_equiangular_gnomonic_map, inverse_equidistant_gnomonic_map
from .device import staged, F
from .halo_data import _dev_of


def ll2cs(cs_grid, latlon_grid):
    """Panel and cell indices (i, j, panel), each [Nlon, Nlat] uint32, of the cubed-sphere cell
    every lat-lon point falls in (src/interpolation.py:38-141; Lauritzen et al. 2015).  The panel is
    the one whose axis carries the largest |coordinate|; on ties the later test of the reference's
    sequence (+X, +Y, -X, -Y, +Z, -Z) wins.  No netCDF cache: the map is recomputed (milliseconds)."""
    X, Y, Z = latlon_grid.X, latlon_grid.Y, latlon_grid.Z
    absv = (np.abs(X), np.abs(Y), np.abs(Z))
    largest = np.maximum(np.maximum(absv[0], absv[1]), absv[2])
    panel = np.zeros(X.shape, dtype=np.uint32)
    for p, (k, positive) in enumerate(((0, X > 0), (1, Y > 0), (0, X < 0), (1, Y < 0), (2, Z > 0), (2, Z <= 0))):
        panel[(largest == absv[k]) & positive] = p
    a = cs_grid.a
    d = 2 * a / cs_grid.N
    inverse = {"gnomonic_equiangular": inverse_equiangular_gnomonic_map,
               "gnomonic_equidistant": inverse_equidistant_gnomonic_map}[cs_grid.projection]
    i = np.zeros(X.shape, dtype=np.uint32)
    j = np.zeros(X.shape, dtype=np.uint32)
    for p in range(nbfaces):
        m = panel == p
        x, y = inverse(X[m], Y[m], Z[m], p)
        i[m] = np.array(np.floor((x + a) / d), dtype=np.uint32)
        j[m] = np.array(np.floor((y + a) / d), dtype=np.uint32)
    return i, j, panel


def nearest_neighbour(Q, cs_grid, latlon_grid):
    """Values of the scalar_field Q (interior cells, [N, N, 6]) at the lat-lon points
    (src/interpolation.py:143-149)."""
    f = np.asarray(Q.f)
    return f[latlon_grid.ix, latlon_grid.jy, latlon_grid.mask]


def ghost_cell_pc_lagrange_interpolation(Q, cs_grid, simulation):
    """Duo-grid Lagrange ghost fill, in place (src/interpolation.py:154-314)."""
    dev = simulation.dev
    with staged(dev, Q, F["USER_A"]) as f:
        dev.call("pycs_halo_fill_dg", f)


def ghost_cells_adjacent_panels(Qx, Qy, cs_grid, simulation):
    """Copy fill from the adjacent panels, in place (src/interpolation.py:320-340)."""
    dev = simulation.dev if simulation is not None else _dev_of(cs_grid, Qx, Qy)
    with staged(dev, Qx, F["USER_A"]) as fx:
        if Qy is Qx:
            dev.call("pycs_halo_fill_copy", fx, fx)
        else:
            with staged(dev, Qy, F["USER_B"]) as fy:
                dev.call("pycs_halo_fill_copy", fx, fy)


def wind_edges2center_cubic_interpolation(U_pc, U_pu, U_pv, cs_grid, simulation):
    """C-grid winds to the cell centres of the boundary ring, to lat-lon, Lagrange ghost fill of both
    components, on the device-resident U_pu / U_pv / U_pc (src/interpolation.py:347-430)."""
    simulation.dev.call("pycs_wind_edges2center")


def wind_center2ghostedge_cubic_interpolation(U_pc, U_pu, U_pv, cs_grid, simulation):
    """Ghost centres to ghost edges and back to contravariant components, on the device-resident
    U_pc / U_pu / U_pv (src/interpolation.py:436-532)."""
    simulation.dev.call("pycs_wind_center2ghostedge")
