"""Convergence of the divergence operator (numeric part of src/operator_accuracy.py:29-147).

Q = 1, one step of each of the reference's four scheme tuples on N = 16, 32, ... with
dt(N) = dt16 * 16 / N, the error being div against the exact divergence of the wind
(src/output.py:152-169).  The reference goes on to plot the table; here it is printed with the
reference's per-run lines (errors.print_errors_simul) and returned."""
import numpy as np

from .advection_ic import adv_simulation_par
from .advection_sphere import adv_sphere
from .configuration import get_advection_parameters
from .cs_datastruct import cubed_sphere
from .errors import print_errors_simul

DT16 = {1: 0.025, 2: 0.0125, 3: 0.00625, 4: 0.0125}       # src/operator_accuracy.py:39-46
# (recon, dp, opsplit, et, mt, mf) of the four columns (src/operator_accuracy.py:57-62)
SCHEMES = {"PL07-RK1": (3, 1, 3, 2, 2, 1), "PL07-RK1-DG-PR": (3, 1, 3, 3, 2, 3),
           "AVLT-RK2-DG-AF": (3, 2, 1, 3, 1, 2), "AVLT-RK2-DG-PR": (3, 2, 1, 3, 1, 3)}


def error_analysis_div(vf, map_projection, plot, transformation, showonscreen, gridload, Ntest=7, pardir=None):
    if vf not in DT16:
        print('ERROR: invalid vector field, ', vf)
        raise SystemExit(1)
    Nc = 16 * 2 ** np.arange(Ntest)
    dts = DT16[vf] * 0.5 ** np.arange(Ntest)
    _, Tf, tc, _, _, _, _, _, _, _, _ = get_advection_parameters(pardir)
    errors = np.zeros((3, Ntest, len(SCHEMES)))            # norm (Linf, L1, L2) x resolution x scheme
    for d, (name, (recon, dp, opsplit, et, mt, mf)) in enumerate(SCHEMES.items()):
        for i in range(Ntest):
            N = int(Nc[i])
            cs_grid = cubed_sphere(N, transformation, False, gridload)
            simulation = adv_simulation_par(cs_grid, dts[i], Tf, 1, vf, tc, recon, dp, opsplit, et, mt, mf)
            print('\nParameters: N=' + str(N) + ', dt=' + str(dts[i]), ', recon=', simulation.recon_name, ', split=',
                  simulation.opsplit_name, ', dp=', simulation.dp_name, ', et=', simulation.et_name)
            errors[:, i, d] = adv_sphere(cs_grid, None, simulation, map_projection, False, True)
            print_errors_simul(errors[0, :, d], errors[1, :, d], errors[2, :, d], i)
            simulation.dev.close()
    return Nc, errors
