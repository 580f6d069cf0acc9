"""One-shot check of a step-kernel tuning point on the GPU: parity against the operator path at small N
(several run calls: separable wind, pending projection), then ms per launch at N=1536 next to the default.
Usage: python scripts/check_variant.py MINB [N_parity ...] [--no-time]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pycs_b200  # noqa
from pycs_b200 import cs_datastruct, advection_ic, advection_vars, advection_timestep

args = [x for x in sys.argv[1:] if x != "--no-time"]
timing = "--no-time" not in sys.argv
minb = args[0] if args else "53"
sizes = [int(x) for x in args[1:]] or [130]
tup = (3, 1, 1, 3, 1, 3)


def sim_of(g, vf, t=tup):
    s = advection_ic.adv_simulation_par(g, 0.00625 * 16 / g.N, 5, 2, vf, 1, *t)
    advection_vars.init_vars_adv(g, s)
    return s


for N in sizes:
    g = cs_datastruct.cubed_sphere(N)
    for vf, t in ((3, tup), (2, (3, 2, 1, 3, 1, 3))):
        os.environ["PYCS_FUSED_MINB"] = minb
        a = sim_of(g, vf, t)
        k = 0
        for n in (1, 4, 7, 9):
            advection_timestep.run_steps(g, a, k, n, fused=True)
            k += n
        name = a.dev.step_kernel_name()
        b = sim_of(g, vf, t)
        advection_timestep.run_steps(g, b, 0, k, fused=False)
        qa, qb = np.asarray(a.Q), np.asarray(b.Q)
        err = np.max(np.abs(qa - qb)) / np.max(np.abs(qb))
        print("parity N=%d vf=%d %s: rel err %.3e %s" % (N, vf, name, err, "OK" if err <= 1e-12 else "FAIL"), flush=True)
        a.dev.close()
        b.dev.close()

if not timing:
    sys.exit(0)
g = cs_datastruct.cubed_sphere(1536)
for mb in (minb, "34", minb):
    os.environ["PYCS_FUSED_MINB"] = mb
    s = sim_of(g, 3)
    ms = C.c_float()
    s.dev.call("pycs_time_step_kernel", 5, 1, C.byref(ms))
    s.dev.call("pycs_time_step_kernel", 40, 1, C.byref(ms))
    a_, b_, c_ = C.c_int32(), C.c_int32(), C.c_int32()
    s.dev.call("pycs_step_kernel_info", C.byref(a_), C.byref(b_), C.byref(c_))
    print("MINB=%s threads=%d rows=%d ctas=%d : %.4f ms" % (mb, a_.value, b_.value, c_.value, ms.value / 40), flush=True)
    s.dev.close()
