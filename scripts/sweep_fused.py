"""Tuning sweep of the fused step kernel (threads per CTA x ring depth x rows per chunk)
at the bench size.  Usage: python scripts/sweep_fused.py [N] ; prints ms per launch."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pycs_b200  # noqa
from pycs_b200 import cs_datastruct, advection_ic, advection_vars

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1536
g = cs_datastruct.cubed_sphere(N)
tup = (3, 1, 1, 3, 1, 3)
configs = [(tb, d, rows) for tb in (128, 160, 256) for d in (5, 6, 7) for rows in (0, 48, 96, 192)
           if not (tb == 256 and d == 7)]
for tb, d, rows in configs:
    os.environ["PYCS_FUSED_TB"] = str(tb)
    os.environ["PYCS_FUSED_DEPTH"] = str(d)
    os.environ["PYCS_FUSED_ROWS"] = str(rows)
    sim = advection_ic.adv_simulation_par(g, 0.00625 * 16 / N, 5, 2, 3, 1, *tup)
    advection_vars.init_vars_adv(g, sim)
    ms = C.c_float()
    sim.dev.call("pycs_time_step_kernel", 5, 1, C.byref(ms))
    sim.dev.call("pycs_time_step_kernel", 40, 1, C.byref(ms))
    a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
    sim.dev.call("pycs_step_kernel_info", C.byref(a), C.byref(b), C.byref(c))
    print("tb=%d depth=%d rows=%d ctas=%d : %.4f ms" % (a.value, d, b.value, c.value, ms.value / 40), flush=True)
    sim.dev.close()
