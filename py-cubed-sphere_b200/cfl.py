"""CFL numbers (src/cfl.py:10-19)."""
import numpy as np

from .device import DeviceArray, F


def _cfl(u_edges, simulation, step, dst, direction):
    if isinstance(u_edges, DeviceArray):
        u_edges.dev.call("pycs_cfl", F[dst], u_edges.fid, direction)
        return u_edges.dev.array(F[dst])
    return np.asarray(u_edges) * simulation.dt / step          # host arrays: plain numpy scalar op


def cfl_x(u_edges, cs_grid, simulation):
    return _cfl(u_edges, simulation, cs_grid.dx, "CX", 0)


def cfl_y(v_edges, cs_grid, simulation):
    return _cfl(v_edges, simulation, cs_grid.dy, "CY", 1)
