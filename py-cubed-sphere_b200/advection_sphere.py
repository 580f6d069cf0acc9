"""Time loop of the advection solver (src/advection_sphere.py:13-61)."""
import numpy as np

from .diagnostics import mass_computation
from .output import output_adv
from .advection_vars import init_vars_adv
from .advection_timestep import adv_time_step, update_adv, run_steps


def adv_sphere(cs_grid, ll_grid, simulation, map_projection, plot, divtest_flag):
    dt, Tf = simulation.dt, simulation.Tf
    Nsteps = int(Tf / dt)
    plotstep = int(Nsteps / 5)
    if divtest_flag:
        Nsteps = 1
        plotstep = 1
    simulation.error_linf, simulation.error_l1, simulation.error_l2 = \
        np.zeros(Nsteps + 1), np.zeros(Nsteps + 1), np.zeros(Nsteps + 1)
    init_vars_adv(cs_grid, simulation)
    simulation.total_mass0, _ = mass_computation(simulation.Q, cs_grid, 1.0)
    if not divtest_flag:
        output_adv(cs_grid, ll_grid, simulation, plot, 0, 0.0, Nsteps, plotstep, map_projection, divtest_flag)
    if plot or divtest_flag:
        # diagnostics wanted after every step: step by step like the reference
        for k in range(1, Nsteps + 1):
            t = k * dt
            adv_time_step(cs_grid, simulation, k, t)
            output_adv(cs_grid, ll_grid, simulation, plot, k, t, Nsteps, plotstep, map_projection, divtest_flag)
            update_adv(cs_grid, simulation, t)
    else:
        # output_adv only acts at k == Nsteps (src/output.py:26): run the loop on the device
        fused = bool(getattr(simulation, "fused", False)) and simulation.dev.fused_supported()
        run_steps(cs_grid, simulation, 0, Nsteps, fused=fused)
        output_adv(cs_grid, ll_grid, simulation, plot, Nsteps, Nsteps * dt, Nsteps, plotstep, map_projection,
                   divtest_flag)
    return simulation.error_linf[Nsteps], simulation.error_l1[Nsteps], simulation.error_l2[Nsteps]
