"""Host helpers of src/sphgeo.py used outside the kernels (set-up and tests)."""
import numpy as np


def sph2cart(lon, lat):
    """src/sphgeo.py:18-22."""
    return np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)


def cart2sph(X, Y, Z):
    """src/sphgeo.py:29-33."""
    return np.arctan2(Y, X), np.arctan2(Z, np.hypot(X, Y))


def latlon_to_contravariant(u_lon, v_lat, ex_lon, ex_lat, ey_lon, ey_lat, det):
    """src/sphgeo.py:120-125 (host; the device does the same in csrc/wind.cu)."""
    return (ey_lat * u_lon - ey_lon * v_lat) / det, (-ex_lat * u_lon + ex_lon * v_lat) / det


def contravariant_to_latlon(ucontra, vcontra, ex_lon, ex_lat, ey_lon, ey_lat):
    """src/sphgeo.py:130-133."""
    return ex_lon * ucontra + ey_lon * vcontra, ex_lat * ucontra + ey_lat * vcontra


class point:
    """src/sphgeo.py:138-147 (arrays are attached by the grid builder)."""
    __slots__ = ("X", "Y", "Z", "lon", "lat")
