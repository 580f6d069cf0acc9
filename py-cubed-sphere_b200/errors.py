"""Error norms (src/errors.py:99-113)."""
import numpy as np


def compute_errors(Q, Qref):
    E = abs(np.asarray(Qref) - np.asarray(Q))
    n = E.size
    return np.amax(abs(E)), np.sum(E) / n, np.sqrt(np.sum(E * E) / n)
