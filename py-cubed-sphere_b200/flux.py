"""PPM upwind fluxes (src/flux.py)."""
from .device import staged, F


def _both(simulation, Qx, Qy, entry):
    dev = simulation.dev
    with staged(dev, Qx, F["USER_A"], writeback=False) as fx:
        if Qy is Qx:
            dev.call(entry, fx, fx)
        else:
            with staged(dev, Qy, F["USER_B"], writeback=False) as fy:
                dev.call(entry, fx, fy)


def compute_fluxes(Qx, Qy, px, py, U_pu, U_pv, cx, cy, cs_grid, simulation):
    """Reconstruction + x and y upwind fluxes (src/flux.py:9-15).  cx, cy, U_pu, U_pv
    are the device-resident simulation fields."""
    _both(simulation, Qx, Qy, "pycs_compute_fluxes")


def numerical_flux_ppm(Qx, Qy, px, py, U_pu, U_pv, cx, cy, cs_grid, simulation):
    """numerical_flux_ppm_x and _y after a reconstruction (src/flux.py:20-128)."""
    _both(simulation, Qx, Qy, "pycs_numerical_flux")
