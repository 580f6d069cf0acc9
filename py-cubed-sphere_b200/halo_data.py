"""Halo strips of the four neighbouring panels (src/halo_data.py:15-400).

The 24 strip orientations are a table inside the library (csrc/halo.cu); these
wrappers return the same 4-tuples / 2-tuples of numpy arrays as the reference.
"""
import ctypes as C

import numpy as np

from .device import staged, F, DeviceArray


def _dev_of(cs_grid, *arrs):
    for a in arrs:
        if isinstance(a, DeviceArray):
            return a.dev
    dev = getattr(cs_grid, "dev", None)
    if dev is None:
        raise RuntimeError("no device bound: pass DeviceArrays or bind cs_grid.dev (pycs_b200.device.Device)")
    return dev


def _gather(Qx, Qy, cs_grid):
    dev = _dev_of(cs_grid, Qx, Qy)
    P = cs_grid.N + cs_grid.ng
    E, W = np.empty((4, P, 6)), np.empty((4, P, 6))
    N_, S = np.empty((P, 4, 6)), np.empty((P, 4, 6))
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    with staged(dev, Qx, F["USER_A"], writeback=False) as fx:
        if Qy is Qx:
            dev.call("pycs_halo_gather", fx, fx, p(E), p(W), p(N_), p(S))
        else:
            with staged(dev, Qy, F["USER_B"], writeback=False) as fy:
                dev.call("pycs_halo_gather", fx, fy, p(E), p(W), p(N_), p(S))
    return E, W, N_, S


def get_halo_data_interpolation(Q, cs_grid):
    """src/halo_data.py:15-185 -> (east, west, north, south)."""
    return _gather(Q, Q, cs_grid)


def get_halo_data_interpolation_NS(Qx, Qy, cs_grid):
    """src/halo_data.py:191-299 -> (north, south)."""
    return _gather(Qx, Qy, cs_grid)[2:]


def get_halo_data_interpolation_WE(Qx, Qy, cs_grid):
    """src/halo_data.py:305-400 -> (east, west)."""
    return _gather(Qx, Qy, cs_grid)[:2]
