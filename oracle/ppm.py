"""Oracle: PPM reconstruction, upwind flux, flux divergence -- one axis at a time.

Every routine works along axis 0 of `[i][j][p]` arrays; the y direction is the
same code applied to `swapaxes(0, 1)` views (the reference spells both out:
src/reconstruction_1d.py:26-192 / :204-372, src/flux.py:20-70 / :75-128,
src/discrete_operators.py:109-120 / :128-138).
"""
import numpy as np


class Parabola:
    """q_L, q_R, dq, q6, f_L, f_R, f_upw, dF of one direction.

    src/cs_datastruct.py:635-684; arrays are stored x-like (the direction of
    the parabola is axis 0); `view()` gives them in grid orientation.
    """

    def __init__(self, P, direction):
        self.direction = direction
        for n in ("q_L", "q_R", "dq", "q6", "dF"):
            setattr(self, "_" + n, np.zeros((P, P, 6)))
        for n in ("f_L", "f_R", "f_upw"):
            setattr(self, "_" + n, np.zeros((P + 1, P, 6)))

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        a = self.__dict__["_" + name]
        return a if self.direction == "x" else np.swapaxes(a, 0, 1)


def reconstruct(Q, par, recon, i0, iend):
    """Edge values q_L, q_R on cells i0-1..iend (axis 0), all other indices.

    Q is x-like (axis 0 = sweep direction).  src/reconstruction_1d.py:26-192.
    """
    qL, qR = par._q_L, par._q_R
    s = lambda a, b: slice(i0 + a, iend + b)
    if recon == "PPM-0":                                   # :36-44
        e = (7.0 / 12.0) * (Q[s(-1, 2)] + Q[s(-2, 1)]) - (Q[s(0, 3)] + Q[s(-3, 0)]) / 12.0
        qR[s(-1, 1)] = e[1:]
        qL[s(-1, 1)] = e[:-1]
    elif recon == "PPM-PL07":                              # :46-62
        a1, a2, a3, a4, a5 = 2.0 / 60.0, -13.0 / 60.0, 47.0 / 60.0, 27.0 / 60.0, -3.0 / 60.0
        q1, q2, q3, q4, q5 = Q[s(-3, -1)], Q[s(-2, 0)], Q[s(-1, 1)], Q[s(0, 2)], Q[s(1, 3)]
        qR[s(-1, 1)] = a1 * q1 + a2 * q2 + a3 * q3 + a4 * q4 + a5 * q5
        qL[s(-1, 1)] = a5 * q1 + a4 * q2 + a3 * q3 + a2 * q4 + a1 * q5
    elif recon == "PPM-CW84":                              # :64-153
        Qm, Q0, Qp = Q[s(-3, 1)], Q[s(-2, 2)], Q[s(-1, 3)]   # cells i0-2..iend+1 centred on Q0
        d0 = 0.5 * (Qp - Qm)
        d1 = 2.0 * (Qp - Q0)
        d2 = 2.0 * (Q0 - Qm)
        dQ = np.minimum(np.minimum(abs(d0), abs(d1)), abs(d2)) * np.sign(d0)
        dQ[~((Qp - Q0) * (Q0 - Qm) > 0.0)] = 0.0
        # edges i0-1..iend+1 (:103-107): dQ index k <-> cell i0-2+k
        e = 0.5 * (Q[s(-1, 2)] + Q[s(-2, 1)]) - (dQ[1:] - dQ[:-1]) / 6.0
        r = e[1:].copy()
        l = e[:-1].copy()
        q = Q[s(-1, 1)]
        dq = r - l                                          # :118-119 (pre-limiter)
        q6 = 6 * q - 3 * (r + l)
        flat = (r - q) * (q - l) <= 0                       # :124-128
        r[flat] = q[flat]
        l[flat] = q[flat]
        over = abs(dq) < abs(q6)                            # :133-134 uses the OLD dq, q6
        left = (r - l) * (q - 0.5 * (r + l)) > ((r - l) ** 2) / 6.0
        right = -((r - l) ** 2) / 6.0 > (r - l) * (q - 0.5 * (r + l))
        ml, mr = over & left, over & right
        l[ml] = 3.0 * q[ml] - 2.0 * r[ml]                   # :149-150
        r[mr] = 3.0 * q[mr] - 2.0 * l[mr]                   # :151-153 (sees updated l)
        qR[s(-1, 1)] = r
        qL[s(-1, 1)] = l
    elif recon == "PPM-L04":                               # :155-192
        Qm, Q0, Qp = Q[s(-4, 2)], Q[s(-3, 3)], Q[s(-2, 4)]   # cells i0-3..iend+2
        dQ = 0.25 * (Qp - Qm)
        dmin = np.maximum(np.maximum(Qm, Q0), Qp) - Q0       # named dQ_min in the reference
        dmax = Q0 - np.minimum(np.minimum(Qm, Q0), Qp)
        mono = np.minimum(np.minimum(abs(dQ), dmin), dmax) * np.sign(dQ)
        # edges i0-1..iend+1: mono index k <-> cell i0-3+k
        e = 0.5 * (Q[s(-1, 2)] + Q[s(-2, 1)]) - (mono[2:-1] - mono[1:-2]) / 3.0
        r, l = e[1:], e[:-1]
        q0 = Q[s(-1, 1)]
        m = mono[2:-2]
        qmin = np.minimum(2.0 * abs(m), abs(l - q0)) * np.sign(2.0 * m)
        qL[s(-1, 1)] = q0 - qmin
        qmin = np.minimum(2.0 * abs(m), abs(r - q0)) * np.sign(2.0 * m)
        qR[s(-1, 1)] = q0 + qmin
    else:
        raise ValueError(recon)


def upwind_flux(Qa, par, c, u_avg, upos, mt, sg_c, sg_e, i0, iend):
    """PPM flux at edges i0..iend (axis 0).  src/flux.py:20-70.

    c = CFL at edges, u_avg = time-averaged contravariant wind, upos = boolean
    mask over edges i0..iend (shape (N+1, P, 6)), sg_c / sg_e = sqrt(g) at
    centres / edges along this axis.
    """
    qL, qR = par._q_L, par._q_R
    cells = slice(i0 - 1, iend + 1)
    if mt == "MT-0":                                       # :27-31 (in place!)
        qL[cells] = qL[cells] * sg_e[i0 - 1:iend + 1]
        qR[cells] = qR[cells] * sg_e[i0:iend + 2]
        q = Qa[cells] * sg_c[cells]
    elif mt == "MT-PL07":
        q = Qa[cells]
    else:
        raise ValueError(mt)
    par._dq[cells] = qR[cells] - qL[cells]                 # :43
    par._q6[cells] = 6 * q - 3 * (qR[cells] + qL[cells])    # :44
    dq, q6 = par._dq, par._q6
    edges = slice(i0, iend + 1)
    left = slice(i0 - 1, iend)                             # upwind cell for u >= 0
    uneg = ~upos
    cc = c[edges]
    fL = qR[left] + cc * 0.5 * (q6[left] - dq[left]) - q6[left] * cc * cc / 3.0     # :48-53
    fR = qL[edges] - cc * 0.5 * (q6[edges] + dq[edges]) - q6[edges] * cc * cc / 3.0  # :56-61
    par._f_L[edges][upos] = fL[upos]
    par._f_R[edges][uneg] = fR[uneg]
    f = par._f_upw[edges]
    f[upos] = fL[upos]                                     # :64-65
    f[uneg] = fR[uneg]
    par._f_upw[edges] = f * u_avg[edges]                   # :66
    if mt == "MT-PL07":                                    # :69-70
        par._f_upw[edges] = par._f_upw[edges] * sg_e[edges]


def flux_difference(par, dt, dx, i0, iend):
    """dF = -(f[i+1]-f[i]) * dt / dx on cells i0..iend-1 (src/discrete_operators.py:109-120)."""
    f = par._f_upw
    d = -(f[i0 + 1:iend + 1] - f[i0:iend])
    par._dF[i0:iend] = d * dt / dx
