"""Readers of the positional `.par` files (src/configuration.py:13-202).

The format is the reference's: a title line, then (comment line, value line) pairs; what
follows the last value is free description.  The reference reads by line position; here the
k-th value is the k-th line that does not start with '#', which is the same for every file
the reference accepts.  Integers are mapped to the same names, the same tuples are returned
in the same order, and bad input prints an error and exits like the reference does.
"""
import os

from . import constants


def _values(path, n, what):
    if not os.path.exists(path):
        print("ERROR in %s: file %s not found in /par." % (what, os.path.basename(path)))
        raise SystemExit(1)
    vals = []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if line and not line.startswith("#"):
                vals.append(line)
                if len(vals) == n:
                    break
    if len(vals) < n:
        print("ERROR in %s: %s holds %d values, %d expected." % (what, path, len(vals), n))
        raise SystemExit(1)
    return vals


def _pardir(pardir):
    """The reference reads ./par/ (src/constants.py:33); without one in the working directory the
    package's own par/ (same format, BASELINE config 1 defaults) is used."""
    if pardir is not None:
        return pardir
    if os.path.isdir(constants.pardir):
        return constants.pardir
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "par")


def get_parameters(pardir=None):
    """par/configuration.par -> N, transformation name, showonscreen, gridload, test_case,
    map projection name (src/configuration.py:13-85)."""
    path = os.path.join(_pardir(pardir), "configuration.par")
    N, transformation, gridload, showonscreen, test_case, map_projection = \
        [int(v) for v in _values(path, 6, "get_parameters")]
    showonscreen, gridload = bool(showonscreen), bool(gridload)
    maps = {1: "mercator", 2: "sphere"}
    transfs = {1: "gnomonic_equidistant", 2: "gnomonic_equiangular", 3: "overlapped"}
    if map_projection not in maps:
        print("ERROR: invalid map projection")
        raise SystemExit(1)
    if transformation not in transfs:
        print("ERROR: invalid transformation")
        raise SystemExit(1)
    if transformation == 3:
        gridload = True                                  # src/configuration.py:62-64
    transf, mp = transfs[transformation], maps[map_projection]
    print("\n--------------------------------------------------------")
    print("Parameters from file", path, "\n")
    print("Number of cells along a coordinate axis: ", N)
    print("Show process on the screen: ", showonscreen)
    print("Loadable grid: ", gridload)
    print("Transformation:", transf)
    print("Test case to be done: ", test_case)
    print("Map projection: ", mp)
    print("--------------------------------------------------------\n")
    return N, transf, showonscreen, gridload, test_case, mp


def get_advection_parameters(pardir=None):
    """par/advection.par -> dt, Tf, tc, ic, vf, recon, dp, opsplit, et, mt, mf
    (src/configuration.py:90-160)."""
    path = os.path.join(_pardir(pardir), "advection.par")
    v = _values(path, 11, "get_advection_parameters")
    Tf, dt = float(v[0]), float(v[1])
    ic, vf, tc, recon, dp, opsplit, et, mt, mf = [int(x) for x in v[2:]]
    print("\n--------------------------------------------------------")
    print("Parameters from file", path, "\n")
    for label, val in (("Total time of integration: ", Tf), ("Time step: ", dt), ("Initial condition: ", ic),
                       ("Vector field: ", vf), ("Adv test case: ", tc), ("Reconstruction scheme: ", recon),
                       ("Departure point scheme: ", dp), ("Splitting scheme: ", opsplit), ("Edge treatment: ", et),
                       ("Metric tensor treatment: ", mt), ("Mass fixer: ", mf)):
        print(label, val)
    print("--------------------------------------------------------\n")
    return dt, Tf, tc, ic, vf, recon, dp, opsplit, et, mt, mf


def get_interpolation_parameters(pardir=None):
    """par/interpolation.par -> tc, ic, vf (src/configuration.py:164-202; the interpolation degree
    is not a file parameter in the reference, the tests sweep it)."""
    path = os.path.join(_pardir(pardir), "interpolation.par")
    tc, ic, vf = [int(x) for x in _values(path, 3, "get_interpolation_parameters")]
    print("\n--------------------------------------------------------")
    print("Parameters from file", path, "\n")
    print("Test case: ", tc)
    print("Scalar field: ", ic)
    print("Vector field: ", vf)
    print("--------------------------------------------------------\n")
    return tc, ic, vf
