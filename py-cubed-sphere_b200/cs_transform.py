"""Sphere -> cube-face coordinates (src/cs_transform.py:101-126, :208-240): the inverse gnomonic maps
used to find the cell a lat-lon point falls in (plot / output regridding, host numpy).

A panel's local frame is the panel-0 frame under the signed axis permutation of
cs_datastruct._ROT, so one table gives numerator / denominator of the two face coordinates."""
import numpy as np

# panel -> ((component, sign) of the x numerator, of the y numerator, of the common denominator)
_FACE = (
    ((1, 1), (2, 1), (0, 1)),      # x ~ Y/X,   y ~ Z/X
    ((0, -1), (2, 1), (1, 1)),     # x ~ -X/Y,  y ~ Z/Y
    ((1, 1), (2, -1), (0, 1)),     # x ~ Y/X,   y ~ -Z/X
    ((0, -1), (2, -1), (1, 1)),    # x ~ -X/Y,  y ~ -Z/Y
    ((1, 1), (0, -1), (2, 1)),     # x ~ Y/Z,   y ~ -X/Z
    ((1, -1), (0, -1), (2, 1)),    # x ~ -Y/Z,  y ~ -X/Z
)


def _ratios(X, Y, Z, panel):
    if not 0 <= panel < 6:
        print("ERROR: invalid panel.")
        raise SystemExit(1)
    v = (X, Y, Z)
    (kx, sx), (ky, sy), (kd, _) = _FACE[panel]
    den = v[kd]
    tx = v[kx] / den
    ty = v[ky] / den
    return (tx if sx > 0 else -tx), (ty if sy > 0 else -ty)


def inverse_equiangular_gnomonic_map(X, Y, Z, panel):
    """Angles (x, y) in [-pi/4, pi/4] of the points (X, Y, Z) of `panel` (src/cs_transform.py:101-126)."""
    tx, ty = _ratios(X, Y, Z, panel)
    return np.arctan(tx), np.arctan(ty)


def inverse_equidistant_gnomonic_map(X, Y, Z, panel):
    """Cube-face coordinates in [-a, a], a = 1/sqrt(3) (src/cs_transform.py:208-240).  The
    reference forms r = +-a/den first and multiplies; the same order is kept (bit-exact indices)."""
    if not 0 <= panel < 6:
        print("ERROR: invalid panel.")
        raise SystemExit(1)
    a = 1.0 / np.sqrt(3.0)
    v = (X, Y, Z)
    # (sign of r, x component and sign, y component and sign, denominator component)
    tab = ((1, (1, 1), (2, 1), 0), (1, (0, -1), (2, 1), 1), (-1, (1, -1), (2, 1), 0),
           (-1, (0, 1), (2, 1), 1), (1, (1, 1), (0, -1), 2), (-1, (1, 1), (0, 1), 2))
    sr, (kx, sx), (ky, sy), kd = tab[panel]
    r = (a if sr > 0 else -a) / v[kd]
    x = (v[kx] if sx > 0 else -v[kx]) * r
    y = (v[ky] if sy > 0 else -v[ky]) * r
    return x, y
