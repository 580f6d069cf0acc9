"""Small fused runs for compute-sanitizer (memcheck / racecheck / synccheck): both shapes of the step at N = 50 and
N = 130, separable and RK2 winds, several run calls with a flush in between.
    compute-sanitizer --tool racecheck python scripts/sanitize_case.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pycs_b200  # noqa
from pycs_b200 import cs_datastruct, advection_ic, advection_vars, advection_timestep

for onek, split in (("1", "0"), ("1", "1"), ("0", "0"), ("0", "1")):     # one-kernel step (uniform grid / CTA table), serial, split
    os.environ["PYCS_ONEKERNEL"] = onek
    os.environ["PYCS_SPLIT"] = split
    for N, vf, tup in ((50, 3, (3, 1, 1, 3, 1, 3)), (130, 2, (3, 2, 1, 3, 1, 3)), (130, 1, (4, 1, 1, 3, 1, 3))):
        g = cs_datastruct.cubed_sphere(N)
        s = advection_ic.adv_simulation_par(g, 0.00625 * 16 / N, 5, 2, vf, 1, *tup)
        advection_vars.init_vars_adv(g, s)
        k = 0
        for n in (2, 3):
            advection_timestep.run_steps(g, s, k, n, fused=True)
            k += n
            q = np.asarray(s.Q)
        print("onekernel=%s split=%s N=%d vf=%d tuple=%s: max|Q| = %.6f" % (onek, split, N, vf, tup, float(np.max(np.abs(q)))), flush=True)
        s.dev.close()
