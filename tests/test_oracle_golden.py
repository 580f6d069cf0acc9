"""CPU: pin the numpy oracle against fixtures produced by the unmodified reference."""
import numpy as np
import pytest

from golden_common import TUPLES, DT16, INTER_KEYS, load, have, relerr
from oracle.grid import LeanGrid
from oracle import halo, step as ost


@pytest.fixture(scope="module")
def g16():
    return LeanGrid(16)


def test_grid_bit_exact(g16):
    ref = load("grid_N16.npz")
    for pos in ("pc", "pu", "pv"):
        assert np.array_equal(ref["sqrtg_" + pos], getattr(g16, "metric_tensor_" + pos)[:, :, 0])
        for nm in ("prod_ex_elon_", "prod_ex_elat_", "prod_ey_elon_", "prod_ey_elat_",
                   "determinant_ll2contra_"):
            assert np.array_equal(ref[nm + pos], getattr(g16, nm + pos)), nm + pos
        assert np.array_equal(ref["lon_" + pos], getattr(g16, pos).lon)
        assert np.array_equal(ref["lat_" + pos], getattr(g16, pos).lat)


def test_halo_index_maps_bit_exact(g16):
    ref = load("halo_index_N16.npz")
    P = 24
    maps = halo.index_maps(g16)
    for s, side in enumerate(("east", "west", "north", "south")):
        nb, I, J = maps[s]
        assert np.array_equal((nb * P + I) * P + J, ref[side]), side


def test_copy_fill(g16):
    ref = load("copyfill_N16.npz")
    Qx, Qy = ref["Qx_in"].copy(), ref["Qy_in"].copy()
    halo.copy_fill(Qx, Qy, g16)
    assert np.array_equal(Qx, ref["Qx_out"]) and np.array_equal(Qy, ref["Qy_out"])


@pytest.mark.parametrize("degree", [0, 1, 2, 3, 4])
def test_lagrange_fill(g16, degree):
    ref = load("halofill_N16.npz")
    tables = halo.lagrange_tables(g16, degree)
    assert np.array_equal(tables[0][0][0], ref["kmin_E_deg%d" % degree])
    assert np.array_equal(tables[1][0], ref["poly_E_deg%d" % degree])
    for ic in (1, 2):
        Qe = ost.q_scalar_field(g16.pc.lon, g16.pc.lat, ic)
        Qn = np.zeros_like(Qe)
        I = np.s_[4:20, 4:20, :]
        Qn[I] = Qe[I]
        halo.dg_fill(Qn, g16, tables)
        assert np.array_equal(Qn, ref["ic%d_deg%d" % (ic, degree)])
        assert ost.compute_errors(Qn, Qe)[0] == float(ref["linf_ic%d_deg%d" % (ic, degree)])


@pytest.mark.skipif(not have("lagrange_tables.npz"), reason="fixture not generated")
@pytest.mark.parametrize("N", [16, 32, 48, 96, 192, 384, 768, 1536, 3072])
def test_stencil_tables_bit_exact(N):
    ref = load("lagrange_tables.npz")
    g = LeanGrid.centres_only(N)
    (kmin, kmax), poly = halo.lagrange_tables(g, 3)
    for s, side in enumerate("EWNS"):
        assert np.array_equal(kmin[s], ref["kmin_%s_N%d" % (side, N)]), (side, N)
        assert np.array_equal(kmax[s], ref["kmax_%s_N%d" % (side, N)]), (side, N)
    assert np.array_equal(poly[0], ref["poly_E_N%d" % N])
    assert np.max(np.abs(poly[0].sum(axis=2) - 1.0)) < 1e-14


@pytest.mark.parametrize("vf", [1, 2, 3, 4])
def test_steps_all_tuples(g16, vf):
    ref = load("steps_N16.npz")
    inter = load("intermediates_N16.npz")
    for name, tup in TUPLES.items():
        recon, dp, split, et, mt, mf = tup
        sim = ost.Simulation(g16, DT16[vf], 5, 2, vf, 1, recon, dp, split, et, mt, mf)
        ost.init_vars_adv(g16, sim)
        if vf == 1 and name == "default":
            assert np.array_equal(sim.Q, ref["Q0"])
        key = "vf%d_%s" % (vf, name)
        ost.run(g16, sim, 1)
        assert np.array_equal(sim.Q, ref[key + "_k1"]), key
        if name in INTER_KEYS and vf in (1, 2):
            for nm in ("q_L", "q_R", "dq", "q6", "f_upw", "dF"):
                assert np.array_equal(getattr(sim.px, nm), inter[key + "_px_" + nm]), (key, nm)
                assert np.array_equal(getattr(sim.py, nm), inter[key + "_py_" + nm]), (key, nm)
            assert np.array_equal(sim.div, inter[key + "_div"])
            assert np.array_equal(sim.U_pu.ucontra_averaged, inter[key + "_uavg"])
            assert np.array_equal(sim.U_pu.upos, inter[key + "_upos"])
        ost.run(g16, sim, 19, 1)
        assert np.array_equal(sim.Q, ref[key + "_k20"]), key


@pytest.mark.skipif(not have("norms.json"), reason="fixture not generated")
def test_full_period_norms_n16():
    import json, os
    from golden_common import GOLDEN
    rows = [r for r in json.load(open(os.path.join(GOLDEN, "norms.json")))
            if r["N"] == 16 and r["vf"] == 1]
    g = LeanGrid(16)
    for r in rows:
        recon, dp, split, et, mt, mf = r["tuple"]
        sim = ost.Simulation(g, r["dt"], 5, 2, 1, 1, recon, dp, split, et, mt, mf)
        ost.init_vars_adv(g, sim)
        ost.run(g, sim, r["steps"])
        e = ost.final_errors(g, sim, r["steps"])
        for a, b in zip(e, (r["linf"], r["l1"], r["l2"])):
            assert abs(a - b) <= 1e-13 * abs(b), r["scheme"]


@pytest.mark.skipif(not have("regrid_N16.npz"), reason="fixture not generated")
@pytest.mark.parametrize("vf", [1, 2, 3, 4])
def test_divergence_test_errors(g16, vf):
    """src/operator_accuracy.py: Q = 1, one step, norms of div - div_exact, against the reference's numbers
    (the oracle's div is the reference's bit for bit; div_exact is the product's host function)."""
    import pycs_b200  # noqa: F401
    from pycs_b200.advection_ic import div_exact
    from pycs_b200.errors import compute_errors
    ref = load("regrid_N16.npz")
    I = np.s_[4:20, 4:20, :]
    for name in ("PL07-RK1", "PL07-RK1-DG-PR", "AVLT-RK2-DG-AF", "AVLT-RK2-DG-PR"):
        recon, dp, split, et, mt, mf = TUPLES[name]
        sim = ost.Simulation(g16, DT16[vf], 5, 1, vf, 1, recon, dp, split, et, mt, mf)
        ost.init_vars_adv(g16, sim)
        ost.adv_time_step(g16, sim, 1, sim.dt)
        got = np.array(compute_errors(sim.div[I], div_exact(g16.pc.lon[I], g16.pc.lat[I], sim)))
        assert np.array_equal(got, ref["diverr_vf%d_%s" % (vf, name)]), (vf, name)


@pytest.mark.skipif(not have("recon_experiment.npz"), reason="fixture not generated")
@pytest.mark.parametrize("N", [16, 32])
def test_reconstruction_experiment_errors(N):
    """src/interpolation_test.py:506-617 with the oracle's ghost fills and PPM edge values: the reference's error
    norms per edge treatment x reconstruction, bit for bit."""
    from oracle import ppm as oppm
    ref = load("recon_experiment.npz")
    g = LeanGrid(N)
    i0, iend = g.i0, g.iend
    I = np.s_[i0:iend, i0:iend, :]
    for ic in (1, 2):
        Qe = ost.q_scalar_field(g.pc.lon, g.pc.lat, ic)
        q_pu = ost.q_scalar_field(g.pu.lon, g.pu.lat, ic)
        q_pv = ost.q_scalar_field(g.pv.lon, g.pv.lat, ic)
        for et in (1, 2, 3):
            for recon in (3, 4):
                sim = ost.Simulation(g, 0.01, 5, 1, 1, 1, recon, 1, 1, et, 1, 1)
                ost.init_vars_adv(g, sim)
                Q = np.zeros_like(Qe)
                Q[I] = Qe[I]
                ost.ghost_fill_scalar(Q, Q, g, sim)
                oppm.reconstruct(Q, sim.px, sim.recon_name, i0, iend)
                oppm.reconstruct(np.swapaxes(Q, 0, 1), sim.py, sim.recon_name, i0, iend)
                if sim.et_name == "ET-PL07":
                    ost.edges_extrapolation(Q, Q, sim.px, sim.py, g, sim)
                e = abs(q_pu[i0:iend, i0:iend, :] - sim.px.q_L[I])
                e = np.maximum(e, abs(q_pu[i0 + 1:iend + 1, i0:iend, :] - sim.px.q_R[I]))
                e = np.maximum(e, abs(q_pv[i0:iend, i0:iend, :] - sim.py.q_L[I]))
                e = np.maximum(e, abs(q_pv[i0:iend, i0 + 1:iend + 1, :] - sim.py.q_R[I]))
                got = np.array(ost.compute_errors(e, 0 * e))
                assert np.array_equal(got, ref["err_N%d_ic%d_et%d_recon%d" % (N, ic, et, recon)]), (ic, et, recon)


@pytest.mark.skipif(not have("vfinterp_experiment.npz"), reason="fixture not generated")
@pytest.mark.parametrize("N", [16, 32])
def test_ghost_edge_wind_experiment_errors(N):
    """src/interpolation_test.py:354-470 with the oracle's wind ghost fill: the eight relative errors per
    wind field and interpolation degree, bit for bit."""
    from oracle import wind as owind
    ref = load("vfinterp_experiment.npz")
    g = LeanGrid(N)
    i0, iend = g.i0, g.iend
    regions = (("pv", np.s_[iend:, i0 - 1:iend + 2, :]), ("pv", np.s_[:i0, i0 - 1:iend + 2, :]),
               ("pu", np.s_[i0 - 1:iend + 2, iend:, :]), ("pu", np.s_[i0 - 1:iend + 2, :i0, :]))
    for vf in (1, 2, 3):
        exact = {}
        for pos in ("pu", "pv"):
            pts = getattr(g, pos)
            ulon, vlat = owind.velocity_adv(pts.lon, pts.lat, 0.0, vf)
            exact[pos] = owind.ll2contra(ulon, vlat, g, pos)
        for degree in (0, 1, 2, 3, 4):
            sim = ost.Simulation(g, 0.01, 5, 1, vf, 1, 3, 2, 1, 3, 1, 1)
            sim.degree = degree
            ost.init_vars_adv(g, sim)
            got = {"pu": (sim.U_pu.ucontra, sim.U_pu.vcontra), "pv": (sim.U_pv.ucontra, sim.U_pv.vcontra)}
            errs = [np.amax(abs(got[pos][c][R] - exact[pos][c][R])) / np.amax(abs(exact[pos][c][R]))
                    for pos, R in regions for c in (0, 1)]
            assert np.array_equal(np.array(errs), ref["err_N%d_vf%d_deg%d" % (N, vf, degree)]), (vf, degree)
