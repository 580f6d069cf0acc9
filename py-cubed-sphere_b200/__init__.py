"""pycs_b200 -- B200-native drop-in for the advection hot path of py-cubed-sphere.

Host-side mirror of the reference's Python operator surface (same module and
function names, same argument order, in-place semantics) over the C ABI of
`lib/libpycs_b200.so` (include/pycs_b200.h).  There is no CPU fallback: the
grid / parameter / table set-up is numpy host code, every operator runs as a
CUDA kernel and raises if the library or a GPU is missing.
"""
__all__ = [
    "constants", "configuration", "cs_datastruct", "sphgeo", "lagrange", "halo_data", "interpolation",
    "edges_treatment", "reconstruction_1d", "flux", "cfl", "averaged_velocity", "discrete_operators",
    "advection_ic", "advection_vars", "advection_timestep", "advection_sphere", "errors", "diagnostics",
    "output", "device", "parallel",
]
