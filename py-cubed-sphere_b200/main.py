"""Driver with the reference's main.py structure (main.py:30-99) for the accelerated test cases:
read par/configuration.par, build the grid, dispatch on test_case.  Test cases 5 (advection,
par/advection.par), 4 (divergence convergence, src/operator_accuracy.py) and 3 (ghost-cell
interpolation experiment, par/interpolation.par) run on the GPU; 1 and 2 (grid plots, grid quality)
are not part of this library.

    python -m pycs_b200.main [pardir]        # from the repository root (pycs_b200.py shim on sys.path)
"""
import sys



def main(pardir=None):
    from .configuration import get_parameters, get_advection_parameters
    from .cs_datastruct import cubed_sphere, latlon_grid
    from .constants import Nlat, Nlon
    from .interpolation import ll2cs
    from .operator_accuracy import error_analysis_div
    from .advection_ic import adv_simulation_par
    from .advection_sphere import adv_sphere
    from .advection_error import error_analysis_adv
    N, transformation, showonscreen, gridload, test_case, map_projection = get_parameters(pardir)
    if test_case in (1, 2):
        print("Test case %d (grid plots / grid quality) is outside the accelerated advection path." % test_case)
        raise SystemExit(1)
    if test_case == 3:
        from .interpolation_test import interpolation_test
        return interpolation_test(map_projection, transformation, showonscreen, True, pardir=pardir)
    if test_case not in (4, 5):
        print("ERROR: invalid testcase.")
        raise SystemExit(1)
    cs_grid = cubed_sphere(N, transformation, showonscreen, gridload)
    dt, Tf, tc, ic, vf, recon, dp, opsplit, et, mt, mf = get_advection_parameters(pardir)
    if test_case == 4:
        # src/operator_accuracy.py:29-147: convergence of div for Q = 1 over N = 16 ... 1024
        print("Test case 4: Divergence test case.\n")
        return error_analysis_div(vf, map_projection, False, transformation, showonscreen, gridload, pardir=pardir)
    # lat-lon grid of the error dump (main.py:57-60 of the reference)
    ll_grid = latlon_grid(Nlat, Nlon)
    ll_grid.ix, ll_grid.jy, ll_grid.mask = ll2cs(cs_grid, ll_grid)
    print("Test case 5: Advection test case.\n")
    simulation = adv_simulation_par(cs_grid, dt, Tf, ic, vf, tc, recon, dp, opsplit, et, mt, mf)
    if simulation.tc == 1:
        simulation.fused = True
        adv_sphere(cs_grid, ll_grid, simulation, map_projection, False, False)
    elif simulation.tc == 2:
        error_analysis_adv(simulation, map_projection, False, transformation, showonscreen, gridload)
    else:
        print('Invalid advection testcase.\n')
        raise SystemExit(1)
    return simulation


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
