"""Host helpers of src/sphgeo.py used outside the kernels (set-up and tests)."""
import numpy as np


def sph2cart(lon, lat):
    """src/sphgeo.py:18-22."""
    return np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)


def cart2sph(X, Y, Z):
    """src/sphgeo.py:29-33."""
    return np.arctan2(Y, X), np.arctan2(Z, np.hypot(X, Y))


class point:
    """src/sphgeo.py:138-147 (arrays are attached by the grid builder)."""
    __slots__ = ("X", "Y", "Z", "lon", "lat")
