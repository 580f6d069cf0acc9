"""Ghost-cell interpolation experiment (par/interpolation.par, test case 1 of
src/interpolation_test.py:65-176): the Lagrange duo-grid fill of an analytic scalar field for
degrees 0..4 on N = 16, 32, ..., error in the ghost cells against the field itself.

Only the scalar test (tc = 1) is on the accelerated path; the vector-field and reconstruction
experiments (tc = 2, 3, 4) and all plots are not provided.  The fill runs on the GPU
(`pycs_halo_fill_dg`); the tables are the host set-up of `lagrange.py`."""
import types

import numpy as np

from .advection_ic import div_exact
from .configuration import get_interpolation_parameters
from .cs_datastruct import cubed_sphere
from .device import Device
from .errors import compute_errors, print_errors_simul
from .interpolation import ghost_cell_pc_lagrange_interpolation
from .lagrange import lagrange_poly_ghostcell_pc
from .sphgeo import sph2cart


class interpolation_simulation_par:
    """src/interpolation_test.py:28-50."""

    def __init__(self, ic, degree):
        self.ic = self.vf = ic
        self.degree = degree
        self.title = 'Interpolation'


def q_scalar_field(lon, lat, simulation):
    """ic 1: Gaussian hill centred at (pi/4, pi/6); ic 2: the trigonometric field that is also the
    exact divergence of wind field 3 (src/interpolation_test.py:55-71, src/advection_ic.py:321-329)."""
    if simulation.ic == 1:
        X0, Y0, Z0 = sph2cart(np.pi / 4.0, np.pi / 6.0)
        X, Y, Z = sph2cart(lon, lat)
        return np.exp(-10.0 * ((X - X0) ** 2 + (Y - Y0) ** 2 + (Z - Z0) ** 2))
    if simulation.ic == 2:
        return div_exact(lon, lat, types.SimpleNamespace(vf=3))
    print("Error - invalid scalar field")
    raise SystemExit(1)


def error_analysis_sf_interpolation(ic, map_projection, transformation, showonscreen, gridload, Ntest=6,
                                    degrees=(0, 1, 2, 3, 4)):
    """Linf error of the ghost-cell fill, [Ntest, len(degrees)] (src/interpolation_test.py:104-176)."""
    Nc = 16 * 2 ** np.arange(Ntest)
    error_linf = np.zeros((Ntest, len(degrees)))
    for d, degree in enumerate(degrees):
        for i in range(Ntest):
            N = int(Nc[i])
            simulation = interpolation_simulation_par(ic, degree)
            print('\nParameters: N = ' + str(N) + ', degree = ' + str(degree))
            cs_grid = cubed_sphere(N, transformation, False, gridload, centres_only=True)
            i0, iend, j0, jend = cs_grid.i0, cs_grid.iend, cs_grid.j0, cs_grid.jend
            Q_exact = q_scalar_field(cs_grid.pc.lon, cs_grid.pc.lat, simulation)
            Q_numerical = np.zeros(np.shape(Q_exact))
            Q_numerical[i0:iend, j0:jend, :] = Q_exact[i0:iend, j0:jend, :]
            simulation.dev = Device(N, cs_grid.dx, cs_grid.dy, 0.01)
            lagrange_poly_ghostcell_pc(cs_grid, simulation)
            ghost_cell_pc_lagrange_interpolation(Q_numerical, cs_grid, simulation)
            error_linf[i, d], _, _ = compute_errors(Q_numerical, Q_exact)
            print_errors_simul(error_linf[:, d], error_linf[:, d], error_linf[:, d], i)
            simulation.dev.close()
        print()
    return Nc, error_linf


def interpolation_test(map_projection, transformation, showonscreen, gridload, pardir=None, Ntest=6):
    tc, ic, vf = get_interpolation_parameters(pardir)
    if tc == 1:
        print("Test case 1: Interpolation of scalar field test case.\n")
        return error_analysis_sf_interpolation(ic, map_projection, transformation, showonscreen, gridload, Ntest)
    if tc in (2, 3, 4):
        print("Interpolation test case %d (vector field / reconstruction experiments) is not provided." % tc)
        raise SystemExit(1)
    print('ERROR in interpolation_test: invalid test case ', tc)
    raise SystemExit(1)
