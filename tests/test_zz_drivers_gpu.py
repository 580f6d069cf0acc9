"""GPU: the reference's experiment drivers (main.py test cases 3 and 4) run through the product modules, and the
split step (interior CTAs beside the ghost fill)."""
import numpy as np
import pytest

from golden_common import TUPLES, DT16, load, have

pytestmark = pytest.mark.gpu

# relative tolerance on the experiment drivers' error norms (north_star: norms within 1e-12 relative)
DRV_TOL = 1e-12


@pytest.fixture(scope="module")
def mods():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pycs_b200  # noqa: F401
    from pycs_b200 import cs_datastruct, advection_ic, advection_sphere
    import types
    return types.SimpleNamespace(**locals())


@pytest.fixture(scope="module")
def g16(mods):
    return mods.cs_datastruct.cubed_sphere(16)


@pytest.mark.skipif(not have("regrid_N16.npz"), reason="fixture not generated")
@pytest.mark.parametrize("vf", [1, 3])
def test_divergence_test_driver_vs_reference(mods, g16, vf):
    """adv_sphere(divtest_flag=True): one step with Q = 1, norms of div - div_exact
    (src/operator_accuracy.py:100, src/output.py:152-169) against the reference's numbers."""
    ref = load("regrid_N16.npz")
    for name in ("PL07-RK1", "PL07-RK1-DG-PR", "AVLT-RK2-DG-AF", "AVLT-RK2-DG-PR"):
        recon, dp, split, et, mt, mf = TUPLES[name]
        sim = mods.advection_ic.adv_simulation_par(g16, DT16[vf], 5, 1, vf, 1, recon, dp, split, et, mt, mf)
        got = np.array(mods.advection_sphere.adv_sphere(g16, None, sim, "mercator", False, True))
        want = ref["diverr_vf%d_%s" % (vf, name)]
        assert np.max(np.abs(got - want) / want) <= DRV_TOL, (vf, name, got, want)
        sim.dev.close()


def test_interpolation_experiment_driver(mods):
    """par/interpolation.par path (src/interpolation_test.py:104-176): Linf of the ghost-cell fill per degree
    at N = 16 against the reference's numbers, and 4th-order decay for degree 3 at N = 32."""
    from pycs_b200.interpolation_test import error_analysis_sf_interpolation
    ref = load("halofill_N16.npz")
    for ic in (1, 2):
        Nc, err = error_analysis_sf_interpolation(ic, "mercator", "gnomonic_equiangular", False, False, Ntest=2)
        assert list(Nc) == [16, 32]
        for d in range(5):
            want = float(ref["linf_ic%d_deg%d" % (ic, d)])
            assert abs(err[0, d] - want) <= 1e-12 * max(want, 1.0), (ic, d, err[0, d], want)
        if ic == 1:                 # 2.605e-3 -> 2.799e-4
            assert err[1, 3] < err[0, 3] / 8.0


@pytest.mark.parametrize("N,vf", [(320, 3), (320, 1), (1536, 3)])
def test_split_step_matches_operator_path(mods, N, vf, monkeypatch):
    """PYCS_SPLIT=1 (DESIGN s7.1): interior CTAs beside the ghost fill, boundary CTAs after it, two streams;
    several run calls (separable wind, pending projection) against the operator path."""
    from pycs_b200 import advection_vars, advection_timestep
    monkeypatch.setenv("PYCS_SPLIT", "1")
    monkeypatch.setenv("PYCS_ONEKERNEL", "0")
    g = mods.cs_datastruct.cubed_sphere(N)

    def sim_of():
        s = mods.advection_ic.adv_simulation_par(g, DT16[vf] * 16 / N, 5, 2, vf, 1, *TUPLES["default"])
        advection_vars.init_vars_adv(g, s)
        return s

    a, b = sim_of(), sim_of()
    k = 0
    for n in ((1, 4, 7) if N < 1000 else (3,)):
        advection_timestep.run_steps(g, a, k, n, fused=True)
        k += n
    advection_timestep.run_steps(g, b, 0, k, fused=False)
    qa, qb = np.asarray(a.Q), np.asarray(b.Q)
    assert np.max(np.abs(qa - qb)) / np.max(np.abs(qb)) <= 1e-12
    a.dev.close()
    b.dev.close()


@pytest.mark.skipif(not have("recon_experiment.npz"), reason="fixture not generated")
def test_reconstruction_experiment_driver(mods):
    """interpolation_test tc 4 (src/interpolation_test.py:506-617): ghost fill + PPM edge values on the GPU
    operators, error norms per edge treatment x reconstruction against the reference's numbers."""
    from pycs_b200.interpolation_test import error_analysis_recon
    ref = load("recon_experiment.npz")
    for ic in (1, 2):
        Nc, err = error_analysis_recon(ic, "mercator", "gnomonic_equiangular", False, False, Ntest=2)
        assert list(Nc) == [16, 32] and err.shape == (2, 3, 2, 3)
        for i, N in enumerate((16, 32)):
            for e, et in enumerate((1, 2, 3)):
                for r, recon in enumerate((3, 4)):
                    want = ref["err_N%d_ic%d_et%d_recon%d" % (N, ic, et, recon)]
                    assert np.max(np.abs(err[i, e, r] - want) / want) <= DRV_TOL, (N, ic, et, recon, err[i, e, r], want)


@pytest.mark.skipif(not have("vfinterp_experiment.npz"), reason="fixture not generated")
@pytest.mark.parametrize("vf", [1, 2, 3])
def test_ghost_edge_wind_experiment_driver(mods, vf):
    """interpolation_test tc 3 (src/interpolation_test.py:354-470) on the device wind ghost fill: the eight
    relative errors per degree at N = 16 against the reference's numbers."""
    from pycs_b200 import advection_vars
    from pycs_b200.interpolation_test import ghost_edge_wind_errors, error_analysis_vf_interpolation_ghost_cells
    ref = load("vfinterp_experiment.npz")
    g = mods.cs_datastruct.cubed_sphere(16)
    for degree in (0, 1, 2, 3, 4):
        sim = mods.advection_ic.adv_simulation_par(g, 0.01, 5.0, 1, vf, 1, 3, 2, 1, 3, 1, 1)
        sim.degree = degree
        advection_vars.init_vars_adv(g, sim)
        got = ghost_edge_wind_errors(g, sim)
        want = ref["err_N16_vf%d_deg%d" % (vf, degree)]
        assert np.max(np.abs(got - want) / want) <= DRV_TOL, (vf, degree, got, want)
        sim.dev.close()
    Nc, err = error_analysis_vf_interpolation_ghost_cells(vf, "mercator", "gnomonic_equiangular", False, False,
                                                          Ntest=2, degrees=(3,))
    assert abs(err[1, 0] - np.max(ref["err_N32_vf%d_deg3" % vf])) <= DRV_TOL * err[1, 0]
