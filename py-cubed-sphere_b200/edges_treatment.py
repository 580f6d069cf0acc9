"""Ghost-cell / cube-edge treatments (src/edges_treatment.py)."""
from .device import staged, F


def edges_ghost_cell_treatment_scalar(Qx, Qy, cs_grid, simulation):
    """Dispatch on simulation.et_name (src/edges_treatment.py:284-290)."""
    dev = simulation.dev
    with staged(dev, Qx, F["USER_A"]) as fx:
        if Qy is Qx:
            dev.call("pycs_halo_fill_scalar", fx, fx)
        else:
            with staged(dev, Qy, F["USER_B"]) as fy:
                dev.call("pycs_halo_fill_scalar", fx, fy)


def edges_ghost_cell_treatment_vector(U_pu, U_pv, U_pc, cs_grid, simulation):
    """Wind ghost fill on the device-resident U_pu / U_pv / U_pc
    (src/edges_treatment.py:296-347; DG path = src/interpolation.py:347-532)."""
    simulation.dev.call("pycs_halo_fill_vector")


def average_flux_cube_edges(px, py, cs_grid):
    """MF-AF: average f_upw on the 12 cube edges (src/edges_treatment.py:231-278)."""
    px.f_upw.dev.call("pycs_average_flux_cube_edges")


def edges_extrapolation(Qx, Qy, px, py, cs_grid, simulation):
    """ET-PL07 edge extrapolation and parabola averaging at the cube edges (src/edges_treatment.py:82-206,
    :31-76) on the device-resident px / py; reconstruction_1d.ppm_reconstruction runs it itself when
    et_name == 'ET-PL07'."""
    dev = simulation.dev
    with staged(dev, Qx, F["USER_A"], writeback=False) as fx:
        if Qy is Qx:
            dev.call("pycs_edges_extrapolation", fx, fx)
        else:
            with staged(dev, Qy, F["USER_B"], writeback=False) as fy:
                dev.call("pycs_edges_extrapolation", fx, fy)
