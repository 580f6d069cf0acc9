"""Import harness for the UNMODIFIED reference (/root/reference/src) in this container.

Only used by tests/golden/make_golden.py (and ad-hoc checks here); nothing that
runs on the GPU box imports this file, because /root/reference does not exist
there.

The reference imports numexpr, matplotlib, cartopy and netCDF4, none of which
are installed.  We inject:
  * a `numexpr` shim whose `evaluate(expr, local_dict=None)` evaluates the
    expression string with numpy in the caller's frame (single thread), and
  * inert stub modules for the plotting / netCDF imports.
No reference source is modified or copied.
"""
import sys
import types
import importlib

import numpy as np

REF_SRC = "/root/reference/src"

_NE_FUNCS = {
    "sin": np.sin, "cos": np.cos, "tan": np.tan, "arctan": np.arctan,
    "sqrt": np.sqrt, "exp": np.exp, "abs": np.abs, "where": np.where,
    "arcsin": np.arcsin, "arccos": np.arccos, "log": np.log,
}


def _evaluate(expr, local_dict=None, global_dict=None, **_kw):
    frame = sys._getframe(1)
    ns = dict(_NE_FUNCS)
    ns.update(frame.f_globals if global_dict is None else global_dict)
    ns.update(frame.f_locals if local_dict is None else local_dict)
    for k, f in _NE_FUNCS.items():      # numexpr function names win over locals
        ns.setdefault(k, f)
    return eval(expr, {"__builtins__": {}}, ns)  # noqa: S307 (trusted strings)


class _Anything(types.ModuleType):
    """Module stub: any attribute is another stub; calling it returns a stub."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        child = _Anything(self.__name__ + "." + name)
        setattr(self, name, child)
        return child

    def __call__(self, *a, **k):
        return _Anything(self.__name__ + "()")


def install():
    """Install stubs and put the reference's src/ on sys.path (idempotent)."""
    if "numexpr" not in sys.modules or not hasattr(sys.modules["numexpr"], "_pycs_shim"):
        ne = types.ModuleType("numexpr")
        ne.evaluate = _evaluate
        ne._pycs_shim = True
        sys.modules["numexpr"] = ne
    for name in (
        "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.cm",
        "matplotlib.ticker", "mpl_toolkits", "mpl_toolkits.axes_grid1",
        "cartopy", "cartopy.crs", "cartopy.feature", "cartopy.mpl",
        "cartopy.mpl.gridliner", "netCDF4",
    ):
        if name not in sys.modules:
            sys.modules[name] = _Anything(name)
    import scipy.special as sp
    if not hasattr(sp, "sph_harm"):
        sp.sph_harm = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)


def ref(module):
    """Import a reference module (e.g. ref('advection_timestep'))."""
    install()
    return importlib.import_module(module)
