"""Error convergence of the advection solver (numeric part of src/advection_error.py:21-188).

The reference sweeps N = 16 ... 512 for four scheme tuples and plots; here the sweep runs the
scheme tuple of `simulation` over `Ns` with the reference's time steps dt(N) = dt16 * 16 / N
(src/advection_error.py:43-58) on the device and prints the error table with the
convergence orders.  Plots and .npy/.txt dumps are out of scope."""
import numpy as np

from .advection_ic import adv_simulation_par
from .advection_sphere import adv_sphere
from .cs_datastruct import cubed_sphere

DT16 = {1: 0.025, 2: 0.0125, 3: 0.00625, 4: 0.0125}     # src/advection_error.py:43-50


def error_analysis_adv(simulation, map_projection, plot, transformation, showonscreen, gridload,
                       Ns=(16, 32, 64, 128), fused=True):
    s = simulation
    errors = np.zeros((len(Ns), 3))
    for n, N in enumerate(Ns):
        dt = DT16[s.vf] * 16.0 / N
        grid = cubed_sphere(N, transformation, showonscreen, gridload)
        sim = adv_simulation_par(grid, dt, 5.0, s.ic, s.vf, 1, s.recon, s.dp, s.opsplit, s.et, s.mt, s.mf)
        sim.fused = fused
        errors[n] = adv_sphere(grid, None, sim, map_projection, False, False)
        sim.dev.close()
    print("\n%6s %12s %6s %12s %6s %12s %6s" % ("N", "Linf", "ord", "L1", "ord", "L2", "ord"))
    for n, N in enumerate(Ns):
        orders = [np.log2(errors[n - 1, c] / errors[n, c]) if n else float("nan") for c in range(3)]
        print("%6d %12.4e %6.2f %12.4e %6.2f %12.4e %6.2f" % (N, errors[n, 0], orders[0], errors[n, 1], orders[1],
                                                              errors[n, 2], orders[2]))
    return np.asarray(Ns), errors
