"""Ghost-cell fills (hot part of src/interpolation.py)."""
from .device import staged, F
from .halo_data import _dev_of


def ghost_cell_pc_lagrange_interpolation(Q, cs_grid, simulation):
    """Duo-grid Lagrange ghost fill, in place (src/interpolation.py:154-314)."""
    dev = simulation.dev
    with staged(dev, Q, F["USER_A"]) as f:
        dev.call("pycs_halo_fill_dg", f)


def ghost_cells_adjacent_panels(Qx, Qy, cs_grid, simulation):
    """Copy fill from the adjacent panels, in place (src/interpolation.py:320-340)."""
    dev = simulation.dev if simulation is not None else _dev_of(cs_grid, Qx, Qy)
    with staged(dev, Qx, F["USER_A"]) as fx:
        if Qy is Qx:
            dev.call("pycs_halo_fill_copy", fx, fx)
        else:
            with staged(dev, Qy, F["USER_B"]) as fy:
                dev.call("pycs_halo_fill_copy", fx, fy)


def wind_edges2center_cubic_interpolation(U_pc, U_pu, U_pv, cs_grid, simulation):
    """src/interpolation.py:347-430 -- fused with the next routine on the device;
    call edges_treatment.edges_ghost_cell_treatment_vector."""
    raise NotImplementedError("use edges_ghost_cell_treatment_vector (both stages run in one call)")


wind_center2ghostedge_cubic_interpolation = wind_edges2center_cubic_interpolation
