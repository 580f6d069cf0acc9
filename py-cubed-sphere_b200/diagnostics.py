"""Mass diagnostic (src/diagnostics.py:14-26), reduced on the device."""
import ctypes as C

from .device import DeviceArray


def mass_computation(Q, cs_grid, total_mass0):
    from .device import F
    if not isinstance(Q, DeviceArray) or Q.fid != F["Q"]:
        raise TypeError("mass_computation expects the device-resident simulation.Q (the reduction runs on F_Q)")
    m = C.c_double()
    Q.dev.call("pycs_mass", C.byref(m))
    total_mass = m.value
    if abs(total_mass0) > 10 ** (-10):
        mass_change = abs(total_mass0 - total_mass) / abs(total_mass0)
    else:
        mass_change = abs(total_mass0 - total_mass)
    return total_mass, mass_change
