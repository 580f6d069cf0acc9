"""One advection step and the wind refresh (src/advection_timestep.py)."""


def adv_time_step(cs_grid, simulation, k, t):
    """Ghost fill -> (wind ghost fill + departure velocity) -> divergence ->
    Q update, operator by operator (src/advection_timestep.py:19-43)."""
    simulation.dev.call("pycs_adv_time_step", int(k), float(t))


def update_adv(cs_grid, simulation, t):
    """Wind at time t for the next step; no-op for vf == 1 (src/advection_timestep.py:48-75)."""
    simulation.dev.call("pycs_update_adv", float(t))


def run_steps(cs_grid, simulation, k0, nsteps, fused=True):
    """The hot loop of adv_sphere for k = k0+1 .. k0+nsteps without host round trips:
    adv_time_step(k, k*dt); update_adv(k*dt) (src/advection_sphere.py:45-57)."""
    simulation.dev.call("pycs_run", int(k0), int(nsteps), 1 if fused else 0)
