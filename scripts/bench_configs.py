"""Throughput of the named BASELINE.json configurations on one GPU (device-resident state,
`pycs_run_timed`, CUDA events): one JSON line per config.  bench.py measures config 4 (the
headline); this script records the others for the profiles.
Usage: python scripts/bench_configs.py [steps]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pycs_b200  # noqa: E402,F401
from pycs_b200 import cs_datastruct, advection_ic, advection_vars  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
# name, N, vf, dt16, (recon, dp, opsplit, et, mt, mf)
CONFIGS = [
    ("config1: N=48 Gaussian hill, solid-body rotation, par defaults", 48, 1, 0.025, (3, 1, 1, 3, 1, 3)),
    ("config2: N=384 same test", 384, 1, 0.025, (3, 1, 1, 3, 1, 3)),
    ("config3: N=768 Nair-Lauritzen non-divergent flow, RK2 (AVLT-RK2-DG-PR)", 768, 2, 0.0125, (3, 2, 1, 3, 1, 3)),
    ("config4: N=1536 divergent flow, par defaults", 1536, 3, 0.00625, (3, 1, 1, 3, 1, 3)),
    ("N=1536 steady wind (vf=1), par defaults", 1536, 1, 0.025, (3, 1, 1, 3, 1, 3)),
    ("N=1536 PL07-RK1-DG-PR (SP-PL07 / MT-PL07)", 1536, 1, 0.025, (3, 1, 3, 3, 2, 3)),
]
for name, N, vf, dt16, tup in CONFIGS:
    g = cs_datastruct.cubed_sphere(N)
    sim = advection_ic.adv_simulation_par(g, dt16 * 16 / N, 5, 2, vf, 1, *tup)
    advection_vars.init_vars_adv(g, sim)
    dev = sim.dev
    dev.call("pycs_run", 0, 10, 1)
    ms = C.c_float()
    l0 = dev.launches()
    dev.call("pycs_run_timed", 10, steps, 1, C.byref(ms))
    print(json.dumps({"config": name, "N": N, "vf": vf, "tuple": tup, "steps": steps,
                      "ms_per_step": ms.value / steps, "cell_updates_per_s": 6.0 * N * N * steps / (ms.value * 1e-3),
                      "launches_per_step": (dev.launches() - l0) / steps, "kernel": dev.step_kernel_name()}), flush=True)
    dev.close()
