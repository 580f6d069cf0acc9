"""Tuning sweep of the fused step kernels at the bench size: prints ms per launch.
Usage: python scripts/sweep_fused.py [N] [v2|v3]"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pycs_b200  # noqa
from pycs_b200 import cs_datastruct, advection_ic, advection_vars

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1536
which = sys.argv[2] if len(sys.argv) > 2 else "v3"
g = cs_datastruct.cubed_sphere(N)
tup = (3, 1, 1, 3, 1, 3)
if which == "v2":
    configs = [dict(IMPL=2, TB=tb, DEPTH=d, ROWS=rows) for tb in (128, 160) for d in (5, 6) for rows in (0, 48, 96)]
elif which == "cs":      # const-slot march against the default point
    configs = [dict(IMPL=4, TB=tb, PF=2, MINB=mb, ROWS=0) for tb, mb in ((160, 14), (160, 34), (128, 35), (160, 53), (128, 54), (160, 64), (160, 33), (160, 34), (160, 64))]
elif which == "v2b":
    pts = ((160, 2, 14), (160, 2, 4), (160, 2, 34), (128, 2, 35), (160, 2, 53), (160, 2, 33), (160, 2, 64))
    configs = [dict(IMPL=4, TB=tb, PF=pf, MINB=mb, ROWS=rows) for tb, pf, mb in pts for rows in (0, 96)]
else:
    pts = ((3, 3, 3), (3, 2, 3), (3, 4, 3), (3, 2, 4), (3, 3, 4), (3, 3, 13), (3, 4, 13), (2, 3, 4), (2, 4, 4), (2, 4, 5),
           (4, 2, 2), (4, 3, 2), (4, 2, 3))
    configs = [dict(IMPL=3, NW=nw, PF=pf, MINB=mb, ROWS=rows) for nw, pf, mb in pts for rows in (0, 64, 128)]
for cfg in configs:
    for k, v in cfg.items():
        os.environ["PYCS_FUSED_" + k] = str(v)
    sim = advection_ic.adv_simulation_par(g, 0.00625 * 16 / N, 5, 2, 3, 1, *tup)
    advection_vars.init_vars_adv(g, sim)
    ms = C.c_float()
    sim.dev.call("pycs_time_step_kernel", 5, 1, C.byref(ms))
    sim.dev.call("pycs_time_step_kernel", 40, 1, C.byref(ms))
    a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
    sim.dev.call("pycs_step_kernel_info", C.byref(a), C.byref(b), C.byref(c))
    print("%s threads=%d rows=%d ctas=%d : %.4f ms" % (cfg, a.value, b.value, c.value, ms.value / 40), flush=True)
    sim.dev.close()
