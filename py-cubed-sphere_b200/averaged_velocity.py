"""Time-averaged (departure point) velocity (src/averaged_velocity.py:14-62)."""


def time_averaged_velocity(cs_grid, simulation):
    simulation.dev.call("pycs_time_averaged_velocity")
