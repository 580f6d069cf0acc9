"""PPM edge reconstruction (src/reconstruction_1d.py)."""
from .device import staged, F


def ppm_reconstruction(Qx, Qy, px, py, cs_grid, simulation):
    """q_L, q_R of px from Qx and of py from Qy, plus the ET-PL07 edge
    extrapolation when selected (src/reconstruction_1d.py:387-394)."""
    dev = simulation.dev
    with staged(dev, Qx, F["USER_A"], writeback=False) as fx:
        if Qy is Qx:
            dev.call("pycs_ppm_reconstruction", fx, fx)
        else:
            with staged(dev, Qy, F["USER_B"], writeback=False) as fy:
                dev.call("pycs_ppm_reconstruction", fx, fy)
