"""Config 5 of BASELINE.json: standalone Lagrange ghost-cell fill (interpolation.par path,
src/interpolation_test.py:104-176) for N = 96 ... 3072.  Prints one JSON line per size:
us per fill (1000 back-to-back fills on the handle's stream, launch overhead included -- the
fill is latency bound), effective GB/s on the algorithmic bytes
(1536*N B per fill: 24 strips x 4 lines x N cells x (8 B read + 8 B write)), and the max
error against the analytic field on the ghost cells (the reference's own check).
Usage: python scripts/bench_halo.py [N ...]"""
import json
import os
import sys
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pycs_b200  # noqa: E402,F401
from pycs_b200 import cs_datastruct, device, lagrange  # noqa: E402
from pycs_b200.sphgeo import sph2cart  # noqa: E402


def q_scalar_field(lon, lat):
    """Gaussian hill of the ghost-cell interpolation test (src/interpolation_test.py:55-62)."""
    X0, Y0, Z0 = sph2cart(np.pi / 4.0, np.pi / 6.0)
    X, Y, Z = sph2cart(lon, lat)
    return np.exp(-10.0 * ((X - X0) ** 2 + (Y - Y0) ** 2 + (Z - Z0) ** 2))


sizes = [int(a) for a in sys.argv[1:]] or [96, 192, 384, 768, 1536, 3072]
for N in sizes:
    g = cs_datastruct.cubed_sphere(N, centres_only=True)
    dev = device.Device(N, g.dx, g.dy, 0.01)
    sim = types.SimpleNamespace(degree=3, dev=dev)
    lagrange.lagrange_poly_ghostcell_pc(g, sim)
    Qe = q_scalar_field(g.pc.lon, g.pc.lat)
    Q = np.zeros_like(Qe)
    I = np.s_[4:N + 4, 4:N + 4, :]
    Q[I] = Qe[I]
    dev.upload(device.F["USER_A"], Q)
    reps = 1000
    for _ in range(20):
        dev.call("pycs_halo_fill_dg", device.F["USER_A"])
    dev.call("pycs_synchronize")
    t = time.perf_counter()
    for _ in range(reps):
        dev.call("pycs_halo_fill_dg", device.F["USER_A"])
    dev.call("pycs_synchronize")
    us = 1e6 * (time.perf_counter() - t) / reps
    out = dev.download(device.F["USER_A"])
    err = float(np.max(np.abs(out - Qe)))
    print(json.dumps({"config": "halo fill (config 5)", "N": N, "us_per_fill": us,
                      "algorithmic_bytes": 1536 * N, "effective_GBs": 1536 * N / us / 1e3,
                      "linf_vs_analytic_field": err, "fills": reps}), flush=True)
    dev.close()
