"""Constants of the reference (src/constants.py:9-37) used on the hot path."""
import math

pi = math.pi
pio2 = pi * 0.5
pio4 = pio2 * 0.5
rad2deg = 180.0 / pi
deg2rad = 1.0 / rad2deg
nbfaces = 6
griddir, datadir, graphdir, pardir = "grid/", "data/", "graphs/", "par/"
Nlat = 720
Nlon = 2 * Nlat
