"""CPU: host-side logic of the product and the C-ABI library surface (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

from golden_common import load, have, GOLDEN

import pycs_b200  # noqa: F401  (import shim)
from pycs_b200 import cs_datastruct, lagrange, device

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "pycs_b200.h")).read()
    names = sorted(set(re.findall(r"\b(pycs_[a-zA-Z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 35
    assert os.path.exists(device.LIB_PATH), "library not built: run __graft_entry__.build()"
    lib = ctypes.CDLL(device.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # the python binding lists the same entry points
    bound = device.load_library()
    for n in names:
        assert hasattr(bound, n)


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(device.PycsError) as e:
        device.Device(16, 0.1, 0.1, 0.01)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_grid_matches_reference_bits():
    ref = load("grid_N16.npz")
    g = cs_datastruct.cubed_sphere(16, "gnomonic_equiangular", False, False)
    for pos in ("pc", "pu", "pv"):
        assert np.array_equal(ref["sqrtg_" + pos], getattr(g, "metric_tensor_" + pos)[:, :, 0]), pos
        for nm in ("prod_ex_elon_", "prod_ex_elat_", "prod_ey_elon_", "prod_ey_elat_", "determinant_ll2contra_"):
            assert np.array_equal(ref[nm + pos], getattr(g, nm + pos)), nm + pos
        assert np.array_equal(ref["lon_" + pos], getattr(g, pos).lon)
        assert np.array_equal(ref["lat_" + pos], getattr(g, pos).lat)
    for c in "XYZ":
        assert np.array_equal(ref[c + "_pc"], getattr(g.pc, c))


@pytest.mark.skipif(not have("lagrange_tables.npz"), reason="fixture not generated")
@pytest.mark.parametrize("N", [16, 48, 96, 384, 768, 1536, 3072])
def test_product_stencil_tables_bit_exact(N):
    """Halo index maps: Kmin / Kmax of the host precompute are bit-exact."""
    ref = load("lagrange_tables.npz")
    import types
    from oracle.grid import LeanGrid          # only as a cheap source of pc.X/Y/Z at large N
    g = LeanGrid.centres_only(N)
    kmin, kmax, w = lagrange.ghost_tables(g, 3)
    assert np.array_equal(kmin, ref["kmin_E_N%d" % N])
    assert np.array_equal(kmax, ref["kmax_E_N%d" % N])
    assert np.array_equal(w, ref["poly_E_N%d" % N])
    sim = types.SimpleNamespace(degree=3)
    lagrange.lagrange_poly_ghostcell_pc(g, sim)
    for s, side in enumerate("EWNS"):
        assert np.array_equal(sim.stencil_ghost_pc[0][s], ref["kmin_%s_N%d" % (side, N)])
        assert np.array_equal(sim.stencil_ghost_pc[1][s], ref["kmax_%s_N%d" % (side, N)])


@pytest.mark.parametrize("degree", [0, 1, 2, 3, 4])
def test_product_tables_all_degrees(degree):
    ref = load("halofill_N16.npz")
    g = cs_datastruct.cubed_sphere(16)
    kmin, kmax, w = lagrange.ghost_tables(g, degree)
    assert np.array_equal(kmin, ref["kmin_E_deg%d" % degree])
    assert np.array_equal(w, ref["poly_E_deg%d" % degree])


def test_invalid_scheme_combination_refused_before_touching_gpu():
    from pycs_b200 import advection_ic
    g = cs_datastruct.cubed_sphere(16)
    with pytest.raises(device.PycsError):
        advection_ic.adv_simulation_par(g, 0.01, 5, 2, 1, 1, 3, 1, 3, 3, 1, 3)   # SP-PL07 with MT-0
    with pytest.raises(SystemExit):
        advection_ic.adv_simulation_par(g, 0.01, 5, 9, 1, 1, 3, 1, 1, 3, 1, 3)   # bad ic


def test_par_readers_follow_the_reference_format(tmp_path, capsys):
    """configuration.get_parameters / get_advection_parameters / get_interpolation_parameters read the
    reference's positional .par format (value on every second line) and return its tuples."""
    import pycs_b200  # noqa: F401
    from pycs_b200 import configuration
    # the package defaults = BASELINE.json config 1
    N, transf, show, load, test_case, mp = configuration.get_parameters()
    assert (N, transf, show, load, test_case, mp) == (48, "gnomonic_equiangular", True, False, 5, "sphere")
    dt, Tf, tc, ic, vf, recon, dp, opsplit, et, mt, mf = configuration.get_advection_parameters()
    assert (Tf, tc, ic, vf, recon, dp, opsplit, et, mt, mf) == (5.0, 1, 2, 1, 3, 1, 1, 3, 1, 3)
    assert abs(dt - 0.025 * 16 / 48) < 1e-15 and int(Tf / dt) == 600
    assert configuration.get_interpolation_parameters() == (1, 1, 2)
    # a file written the way the reference ships it (title, then comment/value pairs, then free text)
    (tmp_path / "advection.par").write_text(
        "#Advection test case parameters \n#Total period definition (seconds)\n5\n#Time step (seconds)\n0.0025\n"
        "#Initial condition\n2\n#Vector field \n3\n#Test case\n2\n#Reconstruction (1, 2, 3, 4)\n4\n"
        "#Departure point scheme (1, 2)\n2\n#Operator splitting (1, 2, 3)\n2\n#Edge treatment (1, 2, 3)\n1\n"
        "#Metric tensor treatment (1, 2)\n1\n#Mass fixer (1, 2, 3)\n2\n#---- Description ----\n#  case(1) - x\n")
    assert configuration.get_advection_parameters(str(tmp_path)) == (0.0025, 5.0, 2, 2, 3, 4, 2, 2, 1, 1, 2)
    (tmp_path / "configuration.par").write_text(
        "#Parameters\n#N\n128\n#Kind of grid\n3\n#Loadable\n0\n#Show\n1\n#Test case\n5\n#Map\n1\n#---\n")
    assert configuration.get_parameters(str(tmp_path)) == (128, "overlapped", True, True, 5, "mercator")
    with pytest.raises(SystemExit):
        configuration.get_interpolation_parameters(str(tmp_path))      # missing file: print + exit
    capsys.readouterr()


# ---- output / regridding row: host numpy, pinned by the reference's own results
@pytest.mark.skipif(not have("regrid_N16.npz"), reason="fixture not generated")
@pytest.mark.parametrize("proj", ["gnomonic_equiangular", "gnomonic_equidistant"])
@pytest.mark.parametrize("N", [16, 21])
def test_ll2cs_index_maps_and_regridding_bit_exact(proj, N):
    from pycs_b200 import interpolation
    ref = load("regrid_N16.npz")
    g = cs_datastruct.cubed_sphere(N, proj, centres_only=True)
    ll = cs_datastruct.latlon_grid(36, 72)
    assert np.array_equal(ll.lon, ref["ll_lon"]) and np.array_equal(ll.lat, ref["ll_lat"])
    ll.ix, ll.jy, ll.mask = interpolation.ll2cs(g, ll)
    key = "%s_N%d" % (proj.split("_")[1], N)
    for mine, name in ((ll.ix, "ix_"), (ll.jy, "jy_"), (ll.mask, "mask_")):
        assert mine.dtype == ref[name + key].dtype and np.array_equal(mine, ref[name + key]), name
    assert set(np.unique(ll.mask)) == set(range(6)) and ll.ix.max() < N and ll.jy.max() < N
    q = cs_datastruct.scalar_field(g, "q", "center")
    q.f[...] = ref["field_" + key]
    assert np.array_equal(interpolation.nearest_neighbour(q, g, ll), ref["regrid_" + key])


def test_error_files_of_the_final_step(tmp_path, monkeypatch):
    """data/<grid>_adv_Q_error_...{.npy, text}: names and contents of src/output.py:136-150."""
    import types
    from pycs_b200 import interpolation, output
    monkeypatch.chdir(tmp_path)
    g = cs_datastruct.cubed_sphere(16, centres_only=True)
    ll = cs_datastruct.latlon_grid(12, 24)
    ll.ix, ll.jy, ll.mask = interpolation.ll2cs(g, ll)
    rng = np.random.default_rng(3)
    q, qe = rng.standard_normal((16, 16, 6)), rng.standard_normal((16, 16, 6))
    sim = types.SimpleNamespace(ic=2, vf=3, degree=3, opsplit_name="SP-AVLT", recon_name="PPM-PL07", dp_name="RK1",
                                et_name="ET-DG", mt_name="MT-0", mf_name="MF-PR", CFL=0.31, mass_change=-2e-16,
                                dt=0.00625, Tf=5.0, error_linf=[0, 1e-2], error_l1=[0, 2e-3], error_l2=[0, 3e-3])
    base = output.save_error_files(g, ll, sim, q, qe, 1)
    assert base == "data/gnomonic_equiangular_cs_16_adv_Q_error_ic2_vf3_SP-AVLT_PPM-PL07_RK1_ET-DG_MT-0_MF-PR_interp3"
    err = np.load(base + ".npy")
    assert err.shape == (24, 12)
    assert np.array_equal(err, (qe - q)[ll.ix, ll.jy, ll.mask])
    assert np.allclose(np.loadtxt(base), [1e-2, 2e-3, 3e-3, 0.31, -2e-16, 0.00625, 5.0], rtol=0, atol=0)


def test_print_errors_simul_lines(capsys):
    from pycs_b200.errors import print_errors_simul
    print_errors_simul([1e-2, 2.5e-3], [1e-3, 2.5e-4], [2e-3, 5e-4], 0)
    print_errors_simul([1e-2, 2.5e-3], [1e-3, 2.5e-4], [2e-3, 5e-4], 1)
    out = capsys.readouterr().out.splitlines()
    assert out[0] == "Error(Linf, L1, L2) : 1.00e-02 1.00e-03 2.00e-03"
    assert out[2] == "Error E_1    : 2.50e-03 2.50e-04 5.00e-04"
    assert out[3] == "Ratio E_1/E_0: 4.00e+00 4.00e+00 4.00e+00"


def test_interpolation_experiment_fields_and_dispatch(tmp_path, capsys):
    """the analytic fields of the ghost-cell experiment equal the reference's bits; test cases 2-4 are refused."""
    from pycs_b200.interpolation_test import q_scalar_field, interpolation_simulation_par, interpolation_test
    ref = load("halofill_N16.npz")
    g = cs_datastruct.cubed_sphere(16, centres_only=True)
    for ic in (1, 2):
        q = q_scalar_field(g.pc.lon, g.pc.lat, interpolation_simulation_par(ic, 3))
        assert np.array_equal(q[4:20, 4:20, :], ref["ic%d_deg3" % ic][4:20, 4:20, :])
    (tmp_path / "interpolation.par").write_text("#tc\n2\n#ic\n1\n#vf\n2\n")
    with pytest.raises(SystemExit):
        interpolation_test("mercator", "gnomonic_equiangular", False, True, pardir=str(tmp_path))
    assert "not provided" in capsys.readouterr().out


def test_divergence_sweep_follows_the_reference_table(monkeypatch, capsys):
    """error_analysis_div: N = 16, 32, ... with dt halved each time, Q = 1 (ic = 1), the four scheme tuples of
    src/operator_accuracy.py:57-62 in the reference's order; device work replaced by fakes."""
    import types
    from pycs_b200 import operator_accuracy as oa
    calls = []

    def fake_sim(grid, dt, Tf, ic, vf, tc, recon, dp, opsplit, et, mt, mf):
        calls.append((grid.N, dt, ic, vf, (recon, dp, opsplit, et, mt, mf)))
        return types.SimpleNamespace(recon_name="r", opsplit_name="s", dp_name="d", et_name="e",
                                     dev=types.SimpleNamespace(close=lambda: None))

    monkeypatch.setattr(oa, "cubed_sphere", lambda N, *a, **k: types.SimpleNamespace(N=N))
    monkeypatch.setattr(oa, "adv_simulation_par", fake_sim)
    monkeypatch.setattr(oa, "adv_sphere", lambda g, ll, sim, mp, plot, divtest: (1.0 / g.N, 2.0 / g.N, 3.0 / g.N)
                        if divtest and not plot else None)
    Nc, err = oa.error_analysis_div(3, "mercator", False, "gnomonic_equiangular", False, False, Ntest=3)
    assert list(Nc) == [16, 32, 64] and err.shape == (3, 3, 4)
    assert [c[0] for c in calls] == [16, 32, 64] * 4
    assert [c[1] for c in calls[:3]] == [0.00625, 0.003125, 0.0015625]
    assert all(c[2] == 1 and c[3] == 3 for c in calls)
    assert [c[4] for c in calls[::3]] == [(3, 1, 3, 2, 2, 1), (3, 1, 3, 3, 2, 3), (3, 2, 1, 3, 1, 2), (3, 2, 1, 3, 1, 3)]
    assert np.allclose(err[0, :, 0], [1 / 16, 1 / 32, 1 / 64])
    assert "Ratio E_1/E_0: 2.00e+00 2.00e+00 2.00e+00" in capsys.readouterr().out


def _split_plan(row_lo, row_hi, ns, band, edge_rows, rows):
    lib = device.load_library()
    ip = ctypes.POINTER(ctypes.c_int32)
    lib.pycs_split_plan.argtypes = [ctypes.c_int32] * 6 + [ip, ctypes.c_int32, ip, ip]
    lib.pycs_split_plan.restype = ctypes.c_int
    cap = 1 << 16
    tab, n, nb = (ctypes.c_int32 * (4 * cap))(), ctypes.c_int32(), ctypes.c_int32()
    assert lib.pycs_split_plan(row_lo, row_hi, ns, band, edge_rows, rows, tab, cap, ctypes.byref(n), ctypes.byref(nb)) == 0
    assert n.value <= cap
    return [tuple(tab[4 * k:4 * k + 4]) for k in range(n.value)], nb.value


@pytest.mark.parametrize("N,world,rank,band,edge_rows,rows", [(1536, 1, 0, 12, 16, 81), (1536, 8, 0, 12, 16, 28),
                                                              (1536, 8, 3, 12, 12, 40), (1536, 2, 1, 8, 24, 90),
                                                              (320, 1, 0, 12, 16, 60), (48, 1, 0, 12, 16, 48),
                                                              (130, 8, 7, 4, 8, 8), (384, 4, 2, 12, 16, 30),
                                                              # uniform tables of the one-kernel step (band = edge = rows)
                                                              (1536, 2, 0, 86, 86, 86), (1536, 2, 1, 86, 86, 86),
                                                              (1536, 4, 1, 43, 43, 43), (1536, 8, 7, 22, 22, 22),
                                                              (48, 2, 1, 24, 24, 24), (130, 2, 0, 33, 33, 33)])
def test_split_step_cta_table(N, world, rank, band, edge_rows, rows):
    """The CTA table of the split step: (1) the CTAs tile the slab exactly once per panel; (2) an interior
    CTA stages no ghost cell and no row outside the slab (rows r0-3 .. r1+2, columns of its strip -3 .. +2);
    (3) everything the exchange plan sends to a peer, and every source of a ghost cell this rank owns, is
    written by a boundary CTA -- those finish first and the exchange / next ghost fill follow them."""
    from pycs_b200.device import mgpu_plan
    lo, hi = 4, N + 4
    base, rem = divmod(N, world)
    a = lo + rank * base + min(rank, rem)
    b = a + base + (1 if rank < rem else 0)
    wmax = 154
    ns = (N + wmax - 1) // wmax
    wcols = (N + ns - 1) // ns
    wcols += wcols & 1
    tab, nb = _split_plan(a, b, ns, band, edge_rows, rows)
    cover = np.zeros((6, N + 8, N + 8), np.int32)
    bnd = np.zeros((6, N + 8, N + 8), bool)
    for k, (r0, r1, strip, panel) in enumerate(tab):
        j0, j1 = lo + strip * wcols, min(lo + (strip + 1) * wcols, hi)
        assert a <= r0 < r1 <= b and 0 <= strip < ns and 0 <= panel < 6
        cover[panel, r0:r1, j0:j1] += 1
        if k < nb:
            bnd[panel, r0:r1, j0:j1] = True
        else:
            assert r0 - 3 >= a and r1 + 2 < b                  # no row of another slab, no W / E ghost row
            assert j0 - 3 >= lo and j1 + 2 < hi                # no S / N ghost column
    assert np.all(cover[:, a:b, lo:hi] == 1) and cover.sum() == 6 * (b - a) * N
    # the 4-wide strips along all four panel edges and the 3 rows next to the neighbouring slabs
    need = np.zeros((N + 8, N + 8), bool)
    need[a:b, lo:lo + 4] = need[a:b, hi - 4:hi] = True
    need[a:min(a + 4, b)] |= True
    need[max(b - 4, a):b] |= True
    need[:, :lo] = need[:, hi:] = False
    assert np.all(bnd[:, need])
    if world > 1 and N <= 400:
        from oracle.grid import LeanGrid
        from oracle import halo as ohalo
        km = ohalo.lagrange_tables(LeanGrid.centres_only(N), 3)[0][0][0]
        for peer, panel, i0, i1, j0, j1 in mgpu_plan(N, world, rank, km, 3)[2]:
            assert np.all(bnd[panel, i0:i1, j0:j1]), (peer, panel, i0, i1, j0, j1)


def _ghost_cells_of_cta(N, r0, r1, j0, j1):
    """Python mirror of the cell enumeration of ghost_prologue (csrc/fused2b.cu, GH = 2): the ghost cells a CTA
    that updates rows [r0, r1) x columns [j0, j1) fills before its first row copy."""
    lo, hi, P = 4, N + 4, N + 8
    Ra, Rb = (0 if r0 - 3 < lo else r0 - 3), (P if r1 + 3 > hi else r1 + 3)
    Ca, Cb = (0 if j0 - 3 < lo else j0 - 3), (P if j1 + 3 > hi else j1 + 3)
    nL, nR, nT, nB = max(0, lo - Ca), max(0, Cb - hi), max(0, lo - Ra), max(0, Rb - hi)
    ncg, ncol = nL + nR, (nL + nR) * (Rb - Ra)
    ja, w, nrg = max(Ca, lo), min(Cb, hi) - max(Ca, lo), nT + nB
    cells = []
    for t in range(ncol + nrg * w):
        if t < ncol:
            c, i = t % ncg, Ra + t // ncg
            j = Ca + c if c < nL else hi + (c - nL)
        else:
            u = t - ncol
            rr, j = u // w, ja + u % w
            i = Ra + rr if rr < nT else hi + (rr - nT)
        cells.append((i, j))
    return cells


@pytest.mark.parametrize("N,world,rank,rows", [(1536, 1, 0, 81), (1536, 2, 1, 86), (1536, 8, 0, 22), (1536, 8, 3, 22),
                                               (320, 1, 0, 60), (48, 1, 0, 48), (16, 1, 0, 16), (130, 2, 0, 33)])
def test_one_kernel_step_ghost_rectangles(N, world, rank, rows):
    """One-kernel step: every CTA fills the ghost cells of the rectangle it stages before it reads them.  Per CTA:
    no cell twice, no interior cell, every ghost cell of the staged rows r0-3 .. r1+2 x columns j0-3 .. j1+2 is in
    the list.  Over a rank's CTAs: all four ghost layers of the rows it reads (slab +- 3) on the S / N sides, the
    whole W / E rows where the slab touches them, and the 4 x 4 corners there -- what the stand-alone ghost fill
    covers for that rank (stepper.cu: launch_ghost_fill), so the ring restore after a run finds a complete ring."""
    lo, hi, P = 4, N + 4, N + 8
    base, rem = divmod(N, world)
    a = lo + rank * base + min(rank, rem)
    b = a + base + (1 if rank < rem else 0)
    wmax = 154
    ns = (N + wmax - 1) // wmax
    wcols = (N + ns - 1) // ns
    wcols += wcols & 1
    tab, _ = _split_plan(a, b, ns, rows, rows, rows)
    filled = np.zeros((P, P), bool)
    for r0, r1, strip, panel in tab:
        if panel:
            continue
        j0, j1 = lo + strip * wcols, min(lo + (strip + 1) * wcols, hi)
        cells = _ghost_cells_of_cta(N, r0, r1, j0, j1)
        assert len(set(cells)) == len(cells)
        got = np.zeros((P, P), bool)
        for i, j in cells:
            assert 0 <= i < P and 0 <= j < P and not (lo <= i < hi and lo <= j < hi), (i, j)
            got[i, j] = True
        staged = np.zeros((P, P), bool)
        staged[max(r0 - 3, 0):min(r1 + 3, P), max(j0 - 3, 0):min(j1 + 3, P)] = True
        staged[lo:hi, lo:hi] = False
        assert np.all(got[staged]), (r0, r1, strip)
        filled |= got
    want = np.zeros((P, P), bool)
    ra, rb = max(a - 3, lo), min(b + 3, hi)
    want[ra:rb, :lo] = want[ra:rb, hi:] = True            # S / N ghost columns of the rows the rank reads
    if a == lo:
        want[:lo, :] = True                               # W ghost rows incl. corners
    if b == hi:
        want[hi:, :] = True                               # E ghost rows incl. corners
    assert np.array_equal(filled, want)


@pytest.mark.parametrize("vf", [1, 2, 3, 4])
def test_host_exact_contravariant_wind_matches_oracle_bits(vf):
    """the analytic wind of the ghost-edge experiment (host numpy of the product) against the oracle's."""
    import types
    from oracle.grid import LeanGrid
    from oracle import wind as owind
    from pycs_b200.advection_ic import velocity_adv
    from pycs_b200.sphgeo import latlon_to_contravariant
    g, og = cs_datastruct.cubed_sphere(16), LeanGrid(16)
    sim = types.SimpleNamespace(vf=vf, ic=1)
    for pos in ("pu", "pv"):
        pts = getattr(g, pos)
        ulon, vlat = velocity_adv(pts.lon, pts.lat, 0.0, sim)
        u, v = latlon_to_contravariant(ulon, vlat, *[getattr(g, n + pos) for n in (
            "prod_ex_elon_", "prod_ex_elat_", "prod_ey_elon_", "prod_ey_elat_", "determinant_ll2contra_")])
        opts = getattr(og, pos)
        ou, ov = owind.ll2contra(*owind.velocity_adv(opts.lon, opts.lat, 0.0, vf), og, pos)
        assert np.array_equal(u, ou) and np.array_equal(v, ov), pos


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference runs without a GPU: one JSON line with the keys the driver reads, the reference
    arm's own cpu_baseline and an e2e that repeats the line's value with zero copied bytes."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-n", "32"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "impl"):
        assert key in d, key
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["unit"] == "cell-updates/s" and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    assert d["config"]["N"] == 1536 and "workload" in d["config"]          # the arm's config is the GPU arm's
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
