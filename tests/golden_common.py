"""Shared constants of the golden fixtures (mirrors tests/golden/make_golden.py)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# (recon, dp, split, et, mt, mf)
TUPLES = {
    "default": (3, 1, 1, 3, 1, 3),
    "PL07-RK1": (3, 1, 3, 2, 2, 1),
    "PL07-RK1-DG-PR": (3, 1, 3, 3, 2, 3),
    "AVLT-RK2-DG-AF": (3, 2, 1, 3, 1, 2),
    "AVLT-RK2-DG-PR": (3, 2, 1, 3, 1, 3),
    "PPM0-S72": (1, 1, 1, 1, 1, 1),
    "CW84-L04-S72-AF": (2, 2, 2, 1, 1, 2),
    "L04-L04-PL07-PR": (4, 2, 2, 2, 1, 3),
    "CW84-PL07-PL07": (2, 1, 3, 2, 2, 1),
    "L04-AVLT-DG-PR": (4, 1, 1, 3, 1, 3),
    "PPM0-RK2-PL07sp-S72-AF": (1, 2, 3, 1, 2, 2),
}
DT16 = {1: 0.025, 2: 0.0125, 3: 0.00625, 4: 0.0125}
INTER_KEYS = ("default", "PL07-RK1", "AVLT-RK2-DG-AF", "L04-L04-PL07-PR")


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def have(name):
    return os.path.exists(os.path.join(GOLDEN, name))


def relerr(a, b):
    """max|a-b| / max|b| (the 1e-12 parity metric of BASELINE.json)."""
    d = float(np.max(np.abs(np.asarray(a) - np.asarray(b))))
    s = float(np.max(np.abs(b)))
    return d / s if s > 0 else d
