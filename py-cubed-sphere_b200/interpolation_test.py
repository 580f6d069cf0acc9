"""Ghost-cell interpolation experiment (par/interpolation.par, test case 1 of
src/interpolation_test.py:65-176): the Lagrange duo-grid fill of an analytic scalar field for
degrees 0..4 on N = 16, 32, ..., error in the ghost cells against the field itself.

Test case 4 (:506-617) is the reconstruction experiment: ghost fill (ET-S72 / ET-PL07 / ET-DG) + PPM edge values
(PPM-PL07, PPM-L04) of the same analytic fields against the field at the cell edges.

Test case 3 (:354-470) is the vector-field ghost-cell experiment: wind at the cell edges -> centres (cubic) ->
Lagrange ghost fill of the lat-lon wind -> ghost edges, relative Linf error of the contravariant wind there.

The scalar test (tc = 1), the ghost-edge wind test (tc = 3) and the reconstruction test (tc = 4) run on the
GPU operators (`pycs_halo_fill_*`, `pycs_halo_fill_vector`, `pycs_ppm_reconstruction`); tc = 2 (error at the
centres: the reference's own routine only fills the ring next to the cube edges, src/interpolation.py:359-418,
so its interior error is the field itself) and all plots are not provided.  The tables are the host set-up of `lagrange.py`."""
import types

import numpy as np

from .advection_ic import div_exact, velocity_adv, adv_simulation_par
from .advection_vars import init_vars_adv
from .configuration import get_interpolation_parameters
from .cs_datastruct import cubed_sphere, ppm_parabola
from .device import Device
from .errors import compute_errors, print_errors_simul
from .edges_treatment import edges_ghost_cell_treatment_scalar
from .interpolation import ghost_cell_pc_lagrange_interpolation
from .reconstruction_1d import ppm_reconstruction
from .lagrange import lagrange_poly_ghostcell_pc
from .sphgeo import sph2cart, latlon_to_contravariant


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


def ghost_edge_wind_errors(cs_grid, simulation):
    """The eight relative errors of src/interpolation_test.py:441-456 (east u, v; west; north; south) of the
    device-resident winds of `simulation` against the analytic wind at t = 0."""
    i0, iend, j0, jend = cs_grid.i0, cs_grid.iend, cs_grid.j0, cs_grid.jend
    exact = {}
    for pos in ("pu", "pv"):
        pts = getattr(cs_grid, pos)
        ulon, vlat = velocity_adv(pts.lon, pts.lat, 0.0, simulation)
        exact[pos] = latlon_to_contravariant(ulon, vlat, getattr(cs_grid, "prod_ex_elon_" + pos),
                                             getattr(cs_grid, "prod_ex_elat_" + pos),
                                             getattr(cs_grid, "prod_ey_elon_" + pos),
                                             getattr(cs_grid, "prod_ey_elat_" + pos),
                                             getattr(cs_grid, "determinant_ll2contra_" + pos))
    got = {"pu": (np.asarray(simulation.U_pu.ucontra), np.asarray(simulation.U_pu.vcontra)),
           "pv": (np.asarray(simulation.U_pv.ucontra), np.asarray(simulation.U_pv.vcontra))}
    regions = (("pv", np.s_[iend:, j0 - 1:jend + 2, :]), ("pv", np.s_[:i0, j0 - 1:jend + 2, :]),
               ("pu", np.s_[i0 - 1:iend + 2, jend:, :]), ("pu", np.s_[i0 - 1:iend + 2, :j0, :]))
    errs = []
    for pos, R in regions:
        for c in (0, 1):
            errs.append(np.amax(abs(got[pos][c][R] - exact[pos][c][R])) / np.amax(abs(exact[pos][c][R])))
    return np.array(errs)


def error_analysis_vf_interpolation_ghost_cells(vf, map_projection, transformation, showonscreen, gridload, Ntest=7,
                                                degrees=(0, 1, 2, 3, 4)):
    """Linf error [Ntest, len(degrees)] of the wind on the ghost edges (src/interpolation_test.py:354-470).  The
    reference sets the edge winds and calls the two interpolation routines; here init_vars_adv does exactly that
    on the device (RK2 build, so that the lines next to the panel are filled as well)."""
    Nc = 16 * 2 ** np.arange(Ntest)
    error_linf = np.zeros((Ntest, len(degrees)))
    for d, degree in enumerate(degrees):
        for i in range(Ntest):
            N = int(Nc[i])
            print('\nParameters: N = ' + str(N) + ', degree = ' + str(degree))
            cs_grid = cubed_sphere(N, transformation, False, gridload)
            simulation = adv_simulation_par(cs_grid, 0.01, 5.0, 1, vf, 1, 3, 2, 1, 3, 1, 1)
            simulation.degree = degree
            init_vars_adv(cs_grid, simulation)
            error_linf[i, d] = np.max(ghost_edge_wind_errors(cs_grid, simulation))
            print_errors_simul(error_linf[:, d], error_linf[:, d], error_linf[:, d], i)
            simulation.dev.close()
        print()
    return Nc, error_linf


class recon_simulation_par:
    """src/interpolation_test.py:660-702; `attach` creates the device handle for one grid."""
    _RECON = {1: 'PPM-0', 2: 'PPM-CW84', 3: 'PPM-PL07', 4: 'PPM-L04'}
    _ET = {1: 'ET-S72', 2: 'ET-PL07', 3: 'ET-DG'}

    def __init__(self, ic, recon, et):
        if ic not in (1, 2, 3):
            print("Error - invalid scalar field")
            raise SystemExit(1)
        if et not in self._ET:
            print('ERROR in recon_simulation_par: invalid ET')
            raise SystemExit(1)
        self.ic, self.recon, self.edge_treatment = ic, recon, et
        self.recon_name, self.et_name = self._RECON.get(recon), self._ET[et]
        self.degree = 3
        self.title = 'Reconstruction'
        self.dev = None

    def attach(self, cs_grid):
        self.dev = Device(cs_grid.N, cs_grid.dx, cs_grid.dy, 0.01, recon=self.recon, et=self.edge_treatment,
                          mf=1, ic=min(self.ic, 2))
        return self


def error_analysis_recon(ic, map_projection, transformation, showonscreen, gridload, Ntest=7, recons=(3, 4)):
    """Error norms [Ntest, len(ets), len(recons), 3] of the reconstructed edge values
    (src/interpolation_test.py:506-617); ET-DG only on the equiangular grid."""
    ets = (1, 2, 3) if transformation == 'gnomonic_equiangular' else (1, 2)
    Nc = 16 * 2 ** np.arange(Ntest)
    errors = np.zeros((Ntest, len(ets), len(recons), 3))
    for e, et in enumerate(ets):
        for r, recon in enumerate(recons):
            for i in range(Ntest):
                N = int(Nc[i])
                cs_grid = cubed_sphere(N, transformation, False, True)
                simulation = recon_simulation_par(ic, recon, et).attach(cs_grid)
                i0, iend, j0, jend = cs_grid.i0, cs_grid.iend, cs_grid.j0, cs_grid.jend
                Q = np.zeros((N + cs_grid.ng, N + cs_grid.ng, 6))
                Qexact = q_scalar_field(cs_grid.pc.lon, cs_grid.pc.lat, simulation)
                q_pu = q_scalar_field(cs_grid.pu.lon, cs_grid.pu.lat, simulation)
                q_pv = q_scalar_field(cs_grid.pv.lon, cs_grid.pv.lat, simulation)
                Q[i0:iend, j0:jend, :] = Qexact[i0:iend, j0:jend, :]
                print('\nParameters: N = ' + str(N) + ', et = ' + str(et) + ' , recon = ', recon)
                if cs_grid.projection == 'gnomonic_equiangular':
                    lagrange_poly_ghostcell_pc(cs_grid, simulation)
                edges_ghost_cell_treatment_scalar(Q, Q, cs_grid, simulation)
                px, py = ppm_parabola(cs_grid, simulation, 'x'), ppm_parabola(cs_grid, simulation, 'y')
                ppm_reconstruction(Q, Q, px, py, cs_grid, simulation)
                I = np.s_[i0:iend, j0:jend, :]
                err = abs(q_pu[i0:iend, j0:jend, :] - np.asarray(px.q_L)[I])
                err = np.maximum(err, abs(q_pu[i0 + 1:iend + 1, j0:jend, :] - np.asarray(px.q_R)[I]))
                err = np.maximum(err, abs(q_pv[i0:iend, j0:jend, :] - np.asarray(py.q_L)[I]))
                err = np.maximum(err, abs(q_pv[i0:iend, j0 + 1:jend + 1, :] - np.asarray(py.q_R)[I]))
                errors[i, e, r] = compute_errors(err, 0 * err)
                print_errors_simul(errors[:, e, r, 0], errors[:, e, r, 1], errors[:, e, r, 2], i)
                simulation.dev.close()
    return Nc, errors


def interpolation_test(map_projection, transformation, showonscreen, gridload, pardir=None, Ntest=6):
    tc, ic, vf = get_interpolation_parameters(pardir)
    if tc == 1:
        print("Test case 1: Interpolation of scalar field test case.\n")
        return error_analysis_sf_interpolation(ic, map_projection, transformation, showonscreen, gridload, Ntest)
    if tc == 4:
        print("Test case 4: Reconstruction test case.\n")
        return error_analysis_recon(ic, map_projection, transformation, showonscreen, gridload, Ntest)
    if tc == 3:
        print("Test case 3: Interpolation of vector field at ghost cells test case.\n")
        return error_analysis_vf_interpolation_ghost_cells(vf, map_projection, transformation, showonscreen, gridload,
                                                           Ntest)
    if tc == 2:
        print("Interpolation test case 2 (vector field at the centres) is not provided.")
        raise SystemExit(1)
    print('ERROR in interpolation_test: invalid test case ', tc)
    raise SystemExit(1)
