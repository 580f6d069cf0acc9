"""Error norms (src/errors.py:99-113)."""
import numpy as np


def compute_errors(Q, Qref):
    E = abs(np.asarray(Qref) - np.asarray(Q))
    n = E.size
    return np.amax(abs(E)), np.sum(E) / n, np.sqrt(np.sum(E * E) / n)


def print_errors_simul(error_linf, error_l1, error_l2, i):
    """Error of run i and its ratio to run i-1 (src/errors.py:87-94)."""
    fmt = "{:.2e}".format
    if i > 0:
        print('Norms          (Linf,    L1,       L2) ')
        print('Error E_' + str(i) + '    :', fmt(error_linf[i]), fmt(error_l1[i]), fmt(error_l2[i]))
        print('Ratio E_' + str(i) + '/E_' + str(i - 1) + ':', fmt(error_linf[i - 1] / error_linf[i]),
              fmt(error_l1[i - 1] / error_l1[i]), fmt(error_l2[i - 1] / error_l2[i]))
    else:
        print('Error(Linf, L1, L2) :', fmt(error_linf[i]), fmt(error_l1[i]), fmt(error_l2[i]))
