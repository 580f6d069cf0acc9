"""Import shim: makes the directory `py-cubed-sphere_b200/` importable as the
package `pycs_b200` (a hyphen cannot appear in a Python package name).

    import pycs_b200
    from pycs_b200 import advection_timestep, cs_datastruct

The package is the host-side mirror of the reference's operator surface; all
arithmetic happens in `py-cubed-sphere_b200/lib/libpycs_b200.so` (CUDA, sm_100a).
"""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "py-cubed-sphere_b200")
_spec = importlib.util.spec_from_file_location(
    "pycs_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["pycs_b200"] = _mod
_spec.loader.exec_module(_mod)
