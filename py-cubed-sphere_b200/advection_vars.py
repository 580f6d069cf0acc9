"""One-time initialisation of the advection variables (src/advection_vars.py:19-107)."""
import numpy as np

from .advection_ic import velocity_adv, q0_adv
from .cs_datastruct import ppm_parabola, velocity
from .lagrange import lagrange_poly_ghostcell_pc
from .edges_treatment import edges_ghost_cell_treatment_vector
from .averaged_velocity import time_averaged_velocity
from .device import F


def init_vars_adv(cs_grid, simulation):
    i0, iend, j0, jend = cs_grid.i0, cs_grid.iend, cs_grid.j0, cs_grid.jend
    P = cs_grid.N + cs_grid.ng
    dev = simulation.dev
    simulation.U_pu = velocity(cs_grid, 'pu', simulation)
    simulation.U_pv = velocity(cs_grid, 'pv', simulation)
    simulation.U_pc = velocity(cs_grid, 'pc', simulation)

    lean = getattr(cs_grid, "lean", False)
    if lean:
        # lean grid: wind at t = 0 and its conversion on the device (the kernels of update_adv)
        for f in ("PU_ULON", "PU_VLAT", "PV_ULON", "PV_VLAT"):
            dev.call("pycs_fill_field", F[f], 0.0)
        dev.call("pycs_init_wind")
    else:
        # winds at t = 0 on the interior edge points (:37-41), evaluated on the host so
        # that the steady wind (vf = 1) enters the device bit-identical to the reference
        for pos, idx, shape in (("pu", np.s_[i0:iend + 1, j0:jend, :], (P + 1, P, 6)),
                                ("pv", np.s_[i0:iend, j0:jend + 1, :], (P, P + 1, 6))):
            pts = getattr(cs_grid, pos)
            ulon, vlat = np.zeros(shape), np.zeros(shape)
            ulon[idx], vlat[idx] = velocity_adv(pts.lon[idx], pts.lat[idx], 0.0, simulation)
            dev.upload(F[pos.upper() + "_ULON"], ulon)
            dev.upload(F[pos.upper() + "_VLAT"], vlat)
        # latlon -> contravariant on the interior (:44-53)
        dev.call("pycs_convert_wind_interior")

    if cs_grid.projection == "gnomonic_equiangular":          # :76-77
        lagrange_poly_ghostcell_pc(cs_grid, simulation)

    edges_ghost_cell_treatment_vector(simulation.U_pu, simulation.U_pv, simulation.U_pc, cs_grid, simulation)  # :80
    dev.call("pycs_copy_field", F["PU_UOLD"], F["PU_UCONTRA"])   # :83-84
    dev.call("pycs_copy_field", F["PV_VOLD"], F["PV_VCONTRA"])
    time_averaged_velocity(cs_grid, simulation)                  # :87

    # CFL (:89-98): cx, cy of the instantaneous wind, max WITHOUT abs inside
    dev.call("pycs_cfl", F["CX"], F["PU_UCONTRA"], 0)
    dev.call("pycs_cfl", F["CY"], F["PV_VCONTRA"], 1)
    CFL_x = dev.field_max(F["CX"], i0, iend + 1, 0, P)            # reduced on the device (a max is exact)
    CFL_y = dev.field_max(F["CY"], 0, P, j0, jend + 1)
    simulation.CFL = max(abs(CFL_x), abs(CFL_y))

    simulation.px = ppm_parabola(cs_grid, simulation, 'x')      # :101-102
    simulation.py = ppm_parabola(cs_grid, simulation, 'y')

    if lean:
        dev.call("pycs_init_tracer", F["Q"], 0.0)                # :105 on the device
    else:
        Q = np.zeros((P, P, 6))                                  # :105
        I = np.s_[i0:iend, j0:jend, :]
        Q[I] = q0_adv(cs_grid.pc.lon[I], cs_grid.pc.lat[I], simulation)
        simulation.Q[...] = Q
