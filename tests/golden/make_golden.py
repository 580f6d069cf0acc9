#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference; see _refshim.py).
Usage:  python tests/golden/make_golden.py small|kmin|norms|big384|big768|big1536|all

  small    N=16 step / operator / halo-fill / grid fixtures (seconds)
  kmin     Lagrange stencil tables (index maps) for N = 16 ... 3072 (bit-exact pin)
  norms    full-period error norms at N=16, 32, 48
  big384   config 2: N=384 default scheme, full period, sub-sampled Q + norms (~50 min)
  big768   config 3: N=768 vf=2 RK2, first 50 steps, sub-sampled Q
  big1536  config 4: N=1536 vf=3, first 20 steps, sub-sampled Q (needs ~15 GB RAM)
  vfinterp the vector-field ghost-cell experiment of interpolation_test (tc 3), per degree at N=16, 32
  recon    the reconstruction experiment of interpolation_test (tc 4): error norms per ET x recon at N=16, 32
  regrid   lat-lon -> cubed-sphere index maps (ll2cs), nearest-neighbour regridding and the
           one-step divergence-test errors (drivers / output row of SURVEY s8 f4)
"""
import json
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refshim as R  # noqa: E402

R.install()
cs = R.ref("cs_datastruct")
rtr = R.ref("cs_transform")
ric = R.ref("advection_ic")
rav = R.ref("advection_vars")
rts = R.ref("advection_timestep")
rlag = R.ref("lagrange")
rint = R.ref("interpolation")
rhalo = R.ref("halo_data")
rerr = R.ref("errors")
rdiag = R.ref("diagnostics")
rtest = R.ref("interpolation_test")

# (recon, dp, split, et, mt, mf); the first five are the reference's own
# tuples (par default + src/advection_error.py:61-66), the rest widen coverage.
TUPLES = {
    "default": (3, 1, 1, 3, 1, 3),
    "PL07-RK1": (3, 1, 3, 2, 2, 1),
    "PL07-RK1-DG-PR": (3, 1, 3, 3, 2, 3),
    "AVLT-RK2-DG-AF": (3, 2, 1, 3, 1, 2),
    "AVLT-RK2-DG-PR": (3, 2, 1, 3, 1, 3),
    "PPM0-S72": (1, 1, 1, 1, 1, 1),
    "CW84-L04-S72-AF": (2, 2, 2, 1, 1, 2),
    "L04-L04-PL07-PR": (4, 2, 2, 2, 1, 3),
    "CW84-PL07-PL07": (2, 1, 3, 2, 2, 1),
    "L04-AVLT-DG-PR": (4, 1, 1, 3, 1, 3),
    "PPM0-RK2-PL07sp-S72-AF": (1, 2, 3, 1, 2, 2),
}
DT16 = {1: 0.025, 2: 0.0125, 3: 0.00625, 4: 0.0125}   # src/advection_error.py:43-50


def grid(N):
    return cs.cubed_sphere(N, "gnomonic_equiangular", False, False)


def new_sim(g, vf, tup, ic=2, dt=None):
    recon, dp, split, et, mt, mf = tup
    dt = DT16[vf] * 16 / g.N if dt is None else dt
    sim = ric.adv_simulation_par(g, dt, 5, ic, vf, 1, recon, dp, split, et, mt, mf)
    rav.init_vars_adv(g, sim)
    return sim


def advance(g, sim, k0, k1):
    for k in range(k0 + 1, k1 + 1):
        t = k * sim.dt
        rts.adv_time_step(g, sim, k, t)
        rts.update_adv(g, sim, t)


def errors(g, sim, k):
    I = np.s_[g.i0:g.iend, g.j0:g.jend, :]
    qe = ric.qexact_adv(g.pc.lon[I], g.pc.lat[I], k * sim.dt, sim)
    return [float(x) for x in rerr.compute_errors(sim.Q[I], qe)]


def sample_index(N, S):
    return np.array(sorted(set(range(0, N, S)) | {1, N - 2, N - 1}))


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print("wrote", name, os.path.getsize(path) // 1024, "KiB", flush=True)


# ---------------------------------------------------------------------------
def make_small():
    N = 16
    g = grid(N)
    # grid inputs (to pin the lean grid builders)
    garr = {}
    for pos in ("pc", "pu", "pv"):
        garr["sqrtg_" + pos] = getattr(g, "metric_tensor_" + pos)[:, :, 0]
        for nm in ("prod_ex_elon_", "prod_ex_elat_", "prod_ey_elon_", "prod_ey_elat_",
                   "determinant_ll2contra_"):
            garr[nm + pos] = getattr(g, nm + pos)
        garr["lon_" + pos] = getattr(g, pos).lon
        garr["lat_" + pos] = getattr(g, pos).lat
    for c in "XYZ":
        garr[c + "_pc"] = getattr(g.pc, c)
    save("grid_N16.npz", **garr)

    # halo index maps: push index-valued fields through the reference gather
    P = N + g.ng
    ii, jj = np.meshgrid(np.arange(P), np.arange(P), indexing="ij")
    code = (np.arange(6)[None, None, :] * P + ii[:, :, None]) * P + jj[:, :, None]
    h = rhalo.get_halo_data_interpolation(code.astype(float), g)
    save("halo_index_N16.npz", east=h[0].astype(np.int64), west=h[1].astype(np.int64),
         north=h[2].astype(np.int64), south=h[3].astype(np.int64))
    # copy fill with distinct Qx / Qy (x/y swap on rotated edges)
    rng = np.random.default_rng(7)
    Qx, Qy = rng.standard_normal((P, P, 6)), rng.standard_normal((P, P, 6))
    Qx0, Qy0 = Qx.copy(), Qy.copy()
    rint.ghost_cells_adjacent_panels(Qx, Qy, g, None)
    save("copyfill_N16.npz", Qx_in=Qx0, Qy_in=Qy0, Qx_out=Qx, Qy_out=Qy)

    # Lagrange fill of the two analytic test fields, degrees 0..4
    out = {}
    for ic in (1, 2):
        for degree in (0, 1, 2, 3, 4):
            sim = types.SimpleNamespace(ic=ic, degree=degree)
            Qe = rtest.q_scalar_field(g.pc.lon, g.pc.lat, sim)
            Qn = np.zeros_like(Qe)
            Qn[g.i0:g.iend, g.j0:g.jend, :] = Qe[g.i0:g.iend, g.j0:g.jend, :]
            rlag.lagrange_poly_ghostcell_pc(g, sim)
            rint.ghost_cell_pc_lagrange_interpolation(Qn, g, sim)
            out["ic%d_deg%d" % (ic, degree)] = Qn
            out["linf_ic%d_deg%d" % (ic, degree)] = np.array(rerr.compute_errors(Qn, Qe)[0])
            if ic == 1:
                out["kmin_E_deg%d" % degree] = np.asarray(sim.stencil_ghost_pc[0][0])
                out["poly_E_deg%d" % degree] = np.asarray(sim.lagrange_poly_ghost_pc[0])
    save("halofill_N16.npz", **out)

    # steps: every tuple x every wind field; Q after 1 and 20 steps
    steps = {}
    inter = {}
    for vf in (1, 2, 3, 4):
        for name, tup in TUPLES.items():
            sim = new_sim(g, vf, tup)
            key = "vf%d_%s" % (vf, name)
            if vf == 1 and name == "default":
                steps["Q0"] = sim.Q.copy()
            advance(g, sim, 0, 1)
            steps[key + "_k1"] = sim.Q.copy()
            if name in ("default", "PL07-RK1", "AVLT-RK2-DG-AF", "L04-L04-PL07-PR") and vf in (1, 2):
                for nm in ("q_L", "q_R", "dq", "q6", "f_upw", "dF"):
                    inter[key + "_px_" + nm] = getattr(sim.px, nm).copy()
                    inter[key + "_py_" + nm] = getattr(sim.py, nm).copy()
                inter[key + "_div"] = sim.div.copy()
                inter[key + "_cx"] = sim.cx.copy()
                inter[key + "_cy"] = sim.cy.copy()
                inter[key + "_uavg"] = sim.U_pu.ucontra_averaged.copy()
                inter[key + "_vavg"] = sim.U_pv.vcontra_averaged.copy()
                inter[key + "_ucontra"] = sim.U_pu.ucontra.copy()
                inter[key + "_vcontra"] = sim.U_pv.vcontra.copy()
                inter[key + "_upos"] = sim.U_pu.upos.copy()
                inter[key + "_vpos"] = sim.U_pv.vpos.copy()
            advance(g, sim, 1, 20)
            steps[key + "_k20"] = sim.Q.copy()
            if name in ("default", "AVLT-RK2-DG-PR") and vf >= 2:
                # wind state after update_adv of step 20
                inter[key + "_k20_ucontra"] = sim.U_pu.ucontra.copy()
                inter[key + "_k20_vcontra"] = sim.U_pv.vcontra.copy()
                inter[key + "_k20_ucontra_old"] = sim.U_pu.ucontra_old.copy()
                inter[key + "_k20_pc_ulon"] = sim.U_pc.ulon.copy()
                inter[key + "_k20_pc_vlat"] = sim.U_pc.vlat.copy()
    save("steps_N16.npz", **steps)
    save("intermediates_N16.npz", **inter)


def make_kmin():
    out = {}
    for N in (16, 32, 48, 96, 192, 384, 768, 1536, 3072):
        # duck-typed grid: only what lagrange_poly_ghostcell_pc reads (src/lagrange.py:29-69)
        a = np.pi / 4
        dx = 2 * a / N
        P = N + 8
        xc = np.linspace(-a + dx / 2.0 - 4 * dx, a - dx / 2.0 + 4 * dx, P)
        X, Y, Z, lon, lat = rtr.equiangular_gnomonic_map(xc, xc, P, P, 1.0)
        g = types.SimpleNamespace(N=N, ng=8, ngl=4, ngr=4, i0=4, iend=N + 4, j0=4, jend=N + 4,
                                  dx=dx, dy=dx, projection="gnomonic_equiangular",
                                  pc=types.SimpleNamespace(X=X, Y=Y, Z=Z))
        sim = types.SimpleNamespace(degree=3)
        rlag.lagrange_poly_ghostcell_pc(g, sim)
        for s, side in enumerate("EWNS"):
            out["kmin_%s_N%d" % (side, N)] = np.asarray(sim.stencil_ghost_pc[0][s]).astype(np.int32)
            out["kmax_%s_N%d" % (side, N)] = np.asarray(sim.stencil_ghost_pc[1][s]).astype(np.int32)
        out["poly_E_N%d" % N] = np.asarray(sim.lagrange_poly_ghost_pc[0])
        print("kmin", N, flush=True)
    save("lagrange_tables.npz", **out)


def make_norms():
    table = []
    for N in (16, 32, 48):
        g = grid(N)
        for vf in (1, 2, 3):
            for name in ("default", "PL07-RK1", "PL07-RK1-DG-PR", "AVLT-RK2-DG-AF", "AVLT-RK2-DG-PR"):
                if N == 48 and not (vf == 1 and name == "default"):
                    continue
                sim = new_sim(g, vf, TUPLES[name])
                nsteps = int(sim.Tf / sim.dt)
                m0, _ = rdiag.mass_computation(sim.Q, g, 1.0)
                t = time.time()
                advance(g, sim, 0, nsteps)
                e = errors(g, sim, nsteps)
                _, dm = rdiag.mass_computation(sim.Q, g, m0)
                row = dict(N=N, vf=vf, scheme=name, tuple=TUPLES[name], steps=nsteps, dt=sim.dt,
                           cfl=float(sim.CFL), linf=e[0], l1=e[1], l2=e[2], mass_change=float(dm),
                           sumQ=float(np.sum(sim.Q[g.i0:g.iend, g.j0:g.jend, :])),
                           cpu_s=time.time() - t)
                table.append(row)
                print(row, flush=True)
                if N == 48:
                    save("config1_N48_final.npz", Q=sim.Q[g.i0:g.iend, g.j0:g.jend, :])
    with open(os.path.join(HERE, "norms.json"), "w") as f:
        json.dump(table, f, indent=1)


def make_big(N, vf, tup_name, checkpoints, S):
    t = time.time()
    g = grid(N)
    print("grid", N, time.time() - t, flush=True)
    sim = new_sim(g, vf, TUPLES[tup_name])
    idx = sample_index(N, S) + g.i0
    out = {"sample_index": idx, "dt": np.array(sim.dt), "tuple": np.array(TUPLES[tup_name]),
           "vf": np.array(vf), "N": np.array(N), "cfl": np.array(sim.CFL)}
    k = 0
    for kk in checkpoints:
        t = time.time()
        advance(g, sim, k, kk)
        k = kk
        out["Q_k%d" % k] = sim.Q[np.ix_(idx, idx, np.arange(6))]
        out["sumQ_k%d" % k] = np.array(np.sum(sim.Q[g.i0:g.iend, g.j0:g.jend, :]))
        out["err_k%d" % k] = np.array(errors(g, sim, k))
        out["mass_k%d" % k] = np.array(rdiag.mass_computation(sim.Q, g, 1.0)[0])
        print("N", N, "step", k, "errs", out["err_k%d" % k], "cpu", time.time() - t, flush=True)
        save("config_N%d_vf%d.npz" % (N, vf), **out)


def make_regrid():
    """src/interpolation.py:38-149 (ll2cs, nearest_neighbour) and the divergence test of
    src/operator_accuracy.py / src/output.py:152-169, for small grids."""
    rint.ll2cs_netcdf = lambda *a, **k: None          # the netCDF cache writer (netCDF4 is a stub here)
    out = {}
    rng = np.random.default_rng(11)
    for proj in ("gnomonic_equiangular", "gnomonic_equidistant"):
        for N in (16, 21):
            g = cs.cubed_sphere(N, proj, False, False)
            ll = cs.latlon_grid(36, 72)
            ix, jy, mask = rint.ll2cs(g, ll)
            ll.ix, ll.jy, ll.mask = ix, jy, mask
            key = "%s_N%d" % (proj.split("_")[1], N)
            out["ix_" + key], out["jy_" + key], out["mask_" + key] = ix, jy, mask
            q = cs.scalar_field(g, "q", "center")
            q.f[:, :, :] = rng.standard_normal(q.f.shape)
            out["field_" + key] = q.f.copy()
            out["regrid_" + key] = rint.nearest_neighbour(q, g, ll)
    out["ll_lon"], out["ll_lat"] = ll.lon, ll.lat
    # divergence test: Q = 1, one step, div against div_exact (src/operator_accuracy.py:29-110)
    g = grid(16)
    I = np.s_[g.i0:g.iend, g.j0:g.jend, :]
    for vf in (1, 2, 3, 4):
        for name in ("PL07-RK1", "PL07-RK1-DG-PR", "AVLT-RK2-DG-AF", "AVLT-RK2-DG-PR"):
            sim = new_sim(g, vf, TUPLES[name], ic=1)
            rts.adv_time_step(g, sim, 1, sim.dt)
            dex = ric.div_exact(g.pc.lon[I], g.pc.lat[I], sim)
            out["diverr_vf%d_%s" % (vf, name)] = np.array(rerr.compute_errors(sim.div[I], dex))
    save("regrid_N16.npz", **out)


def make_recon():
    """Reconstruction experiment (src/interpolation_test.py:506-617): ghost fill + PPM edge values of an
    analytic field against the field at the edges, per edge treatment x reconstruction."""
    redge = R.ref("edges_treatment")
    rrec = R.ref("reconstruction_1d")
    out = {}
    for N in (16, 32):
        g = grid(N)
        i0, iend, j0, jend = g.i0, g.iend, g.j0, g.jend
        for ic in (1, 2):
            for et in (1, 2, 3):
                for recon in (3, 4):
                    sim = rtest.recon_simulation_par(ic, recon, et)
                    Q = np.zeros((N + g.ng, N + g.ng, 6))
                    Qe = rtest.q_scalar_field(g.pc.lon, g.pc.lat, sim)
                    q_pu = rtest.q_scalar_field(g.pu.lon, g.pu.lat, sim)
                    q_pv = rtest.q_scalar_field(g.pv.lon, g.pv.lat, sim)
                    Q[i0:iend, j0:jend, :] = Qe[i0:iend, j0:jend, :]
                    rlag.lagrange_poly_ghostcell_pc(g, sim)
                    redge.edges_ghost_cell_treatment_scalar(Q, Q, g, sim)
                    px, py = cs.ppm_parabola(g, sim, 'x'), cs.ppm_parabola(g, sim, 'y')
                    rrec.ppm_reconstruction(Q, Q, px, py, g, sim)
                    e = abs(q_pu[i0:iend, j0:jend, :] - px.q_L[i0:iend, j0:jend, :])
                    e = np.maximum(e, abs(q_pu[i0 + 1:iend + 1, j0:jend, :] - px.q_R[i0:iend, j0:jend, :]))
                    e = np.maximum(e, abs(q_pv[i0:iend, j0:jend, :] - py.q_L[i0:iend:, j0:jend, :]))
                    e = np.maximum(e, abs(q_pv[i0:iend, j0 + 1:jend + 1, :] - py.q_R[i0:iend:, j0:jend, :]))
                    out["err_N%d_ic%d_et%d_recon%d" % (N, ic, et, recon)] = np.array(rerr.compute_errors(e, 0 * e))
    save("recon_experiment.npz", **out)


def make_vfinterp():
    """Vector-field ghost-cell experiment (src/interpolation_test.py:354-470): wind from the cell edges to the
    centres (cubic), Lagrange ghost fill of the lat-lon wind, centres to ghost edges; relative Linf error of the
    contravariant wind on the ghost edges, per interpolation degree."""
    sg = R.ref("sphgeo")
    out = {}
    for N in (16, 32):
        g = grid(N)
        i0, iend, j0, jend = g.i0, g.iend, g.j0, g.jend
        for vf in (1, 2, 3):
            ex = {}
            for pos in ("pu", "pv"):
                pts = getattr(g, pos)
                sim0 = rtest.interpolation_simulation_par(vf, 3)
                ulon, vlat = ric.velocity_adv(pts.lon, pts.lat, 0.0, sim0)
                ex[pos] = sg.latlon_to_contravariant(ulon, vlat, getattr(g, "prod_ex_elon_" + pos),
                                                     getattr(g, "prod_ex_elat_" + pos), getattr(g, "prod_ey_elon_" + pos),
                                                     getattr(g, "prod_ey_elat_" + pos),
                                                     getattr(g, "determinant_ll2contra_" + pos))
            for degree in (0, 1, 2, 3, 4):
                sim = rtest.interpolation_simulation_par(vf, degree)
                U_pu, U_pv, U_pc = cs.velocity(g, 'pu'), cs.velocity(g, 'pv'), cs.velocity(g, 'pc')
                U_pu.ucontra[i0:iend + 1, j0:jend, :] = ex["pu"][0][i0:iend + 1, j0:jend, :]
                U_pv.vcontra[i0:iend, j0:jend + 1, :] = ex["pv"][1][i0:iend, j0:jend + 1, :]
                rlag.lagrange_poly_ghostcell_pc(g, sim)
                rint.wind_edges2center_cubic_interpolation(U_pc, U_pu, U_pv, g, sim)
                rint.wind_center2ghostedge_cubic_interpolation(U_pc, U_pu, U_pv, g, sim)
                rel = lambda a, b: np.amax(abs(a - b)) / np.amax(abs(b))
                E = np.s_[iend:, j0 - 1:jend + 2, :]
                W = np.s_[:i0, j0 - 1:jend + 2, :]
                Nn = np.s_[i0 - 1:iend + 2, jend:, :]
                S = np.s_[i0 - 1:iend + 2, :j0, :]
                errs = [rel(U_pv.ucontra[E], ex["pv"][0][E]), rel(U_pv.vcontra[E], ex["pv"][1][E]),
                        rel(U_pv.ucontra[W], ex["pv"][0][W]), rel(U_pv.vcontra[W], ex["pv"][1][W]),
                        rel(U_pu.ucontra[Nn], ex["pu"][0][Nn]), rel(U_pu.vcontra[Nn], ex["pu"][1][Nn]),
                        rel(U_pu.ucontra[S], ex["pu"][0][S]), rel(U_pu.vcontra[S], ex["pu"][1][S])]
                out["err_N%d_vf%d_deg%d" % (N, vf, degree)] = np.array(errs)
    save("vfinterp_experiment.npz", **out)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "small"
    if what in ("small", "all"):
        make_small()
    if what in ("kmin", "all"):
        make_kmin()
    if what in ("norms", "all"):
        make_norms()
    if what in ("regrid", "all"):
        make_regrid()
    if what in ("recon", "all"):
        make_recon()
    if what in ("vfinterp", "all"):
        make_vfinterp()
    if what in ("big384", "all"):
        make_big(384, 1, "default", [1, 10, 100, 1000, 4800], 8)
    if what in ("big768", "all"):
        make_big(768, 2, "AVLT-RK2-DG-PR", [1, 10, 50], 16)
    if what in ("big1536", "all"):
        make_big(1536, 3, "default", [1, 5, 20], 32)
