"""ms per launch of the step kernel alone at the bench size (pycs_time_step_kernel: CUDA events around back-to-back
launches over the whole grid), for the environment it is started with.  Usage: [PYCS_VARIANT=v] python scripts/time_kernel.py [N]"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pycs_b200  # noqa
from pycs_b200 import cs_datastruct, advection_ic, advection_vars

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1536
g = cs_datastruct.cubed_sphere(N, lean=True)
sim = advection_ic.adv_simulation_par(g, 0.00625 * 16 / N, 5, 2, 3, 1, 3, 1, 1, 3, 1, 3)
advection_vars.init_vars_adv(g, sim)
ms = C.c_float()
sim.dev.call("pycs_time_step_kernel", 5, 1, C.byref(ms))
best = 1e9
for _ in range(4):
    sim.dev.call("pycs_time_step_kernel", 50, 1, C.byref(ms))
    best = min(best, ms.value / 50)
env = {k: v for k, v in os.environ.items() if k.startswith("PYCS_")}
print("%s N=%d %s : %.4f ms per launch (best of 4 bursts of 50)" % (sim.dev.step_kernel_name(), N, env, best), flush=True)
