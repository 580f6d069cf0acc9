"""Diagnostics part of output_adv (src/output.py:22-50); plots and file dumps are out of scope."""
import numpy as np

from .errors import compute_errors
from .advection_ic import qexact_adv
from .diagnostics import mass_computation


def print_diagnostics_adv(error_linf, error_l1, error_l2, mass_change, t, Nsteps):
    print('\nStep', t, 'from', Nsteps)
    print('Error (Linf, L1, L2) :', "{:.2e}".format(error_linf), "{:.2e}".format(error_l1), "{:.2e}".format(error_l2))
    print('Total mass variation:', "{:.2e}".format(mass_change))


def output_adv(cs_grid, ll_grid, simulation, plot, k, t, Nsteps, plotstep, map_projection, divtest_flag):
    if plot or k == Nsteps:
        i0, iend, j0, jend = cs_grid.i0, cs_grid.iend, cs_grid.j0, cs_grid.jend
        I = np.s_[i0:iend, j0:jend, :]
        q_exact = qexact_adv(cs_grid.pc.lon[I], cs_grid.pc.lat[I], t, simulation)
        q = simulation.Q[I]
        simulation.error_linf[k], simulation.error_l1[k], simulation.error_l2[k] = compute_errors(q, q_exact)
        if simulation.error_linf[k] > 100.0:
            print('Stopping due to large errors.')
            print('The CFL number is:', simulation.CFL)
            raise SystemExit(1)
        simulation.total_mass, simulation.mass_change = \
            mass_computation(simulation.Q, cs_grid, simulation.total_mass0)
        if k > 0 and (not divtest_flag):
            print_diagnostics_adv(simulation.error_linf[k], simulation.error_l1[k], simulation.error_l2[k],
                                  simulation.mass_change, k, Nsteps)
