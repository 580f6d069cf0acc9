"""Oracle: simulation state and the advection step.

  src/advection_ic.py:21-203     adv_simulation_par      -> Simulation
  src/advection_ic.py:215-281    qexact_adv / q0_adv     -> qexact_adv()
  src/advection_vars.py:19-107   init_vars_adv           -> init_vars_adv()
  src/advection_timestep.py:19-75  adv_time_step, update_adv
  src/discrete_operators.py:18-138 divergence, F/G operators
  src/flux.py:9-15               compute_fluxes
  src/reconstruction_1d.py:387-394 ppm_reconstruction
  src/edges_treatment.py:31-278  edge extrapolation / parabola + flux averaging
  src/errors.py:99-113, src/diagnostics.py:14-26  error norms, mass
"""
import numpy as np

from . import halo, ppm, wind
from .wind import pi, deg2rad

RECON = {1: "PPM-0", 2: "PPM-CW84", 3: "PPM-PL07", 4: "PPM-L04"}
DP = {1: "RK1", 2: "RK2"}
SPLIT = {1: "SP-AVLT", 2: "SP-L04", 3: "SP-PL07"}
ET = {1: "ET-S72", 2: "ET-PL07", 3: "ET-DG"}
MT = {1: "MT-0", 2: "MT-PL07"}
MF = {1: "MF-0", 2: "MF-AF", 3: "MF-PR"}

# The 12 cube edges as (dir, side, panel) pairs + flip of the along-edge index.
# dir 'x': the edge is an i = const line of that panel, 'y': j = const;
# side 'lo' / 'hi' = at i0 (j0) / iend (jend).  From the pairings spelled out
# in src/edges_treatment.py:38-76 (and reused at :135-190, :242-278).
CUBE_EDGES = (
    (("x", "hi", 0), ("x", "lo", 1), False), (("x", "hi", 1), ("x", "lo", 2), False),
    (("x", "hi", 2), ("x", "lo", 3), False), (("x", "hi", 3), ("x", "lo", 0), False),
    (("y", "lo", 4), ("y", "hi", 0), False), (("x", "hi", 4), ("y", "hi", 1), False),
    (("y", "hi", 4), ("y", "hi", 2), True), (("x", "lo", 4), ("y", "hi", 3), True),
    (("y", "hi", 5), ("y", "lo", 0), False), (("y", "lo", 1), ("x", "hi", 5), True),
    (("y", "lo", 2), ("y", "lo", 5), True), (("y", "lo", 3), ("x", "lo", 5), False),
)


class Simulation:
    """src/advection_ic.py:21-203 (names, scalars and state arrays)."""

    def __init__(self, g, dt, Tf, ic, vf, tc, recon, dp, opsplit, et, mt, mf):
        self.ic, self.vf, self.tc = ic, vf, tc
        self.dt, self.dto2, self.twodt = dt, dt * 0.5, dt * 2.0
        self.Tf = Tf
        self.degree = 3                                     # :50
        if ic not in (1, 2, 3, 4) or vf not in (1, 2, 3, 4):
            raise ValueError("invalid ic / vf")
        self.recon_name, self.dp_name = RECON[recon], DP[dp]
        self.opsplit_name, self.et_name = SPLIT[opsplit], ET[et]
        self.mt_name, self.mf_name = MT[mt], MF[mf]
        P = g.N + g.ng
        self.Q = np.zeros((P, P, 6))
        self.gQ = np.zeros((P, P, 6))
        self.div = np.zeros((P, P, 6))
        self.cx = np.zeros((P + 1, P, 6))
        self.cy = np.zeros((P, P + 1, 6))
        self.CFL = 0.0
        self.tables = None


def sph2cart(lon, lat):
    return np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)


def qexact_adv(lon, lat, t, sim):
    """src/advection_ic.py:215-281."""
    if sim.ic == 1:
        return np.ones(np.shape(lon))
    if sim.ic == 2:
        X, Y, Z = sph2cart(lon, lat)
        if sim.vf == 1:
            alpha = -45.0 * deg2rad
            wt = -(2.0 * pi / 5.0) * t
            cosa, sina = np.cos(alpha), np.sin(alpha)
            cos2a, sin2a = cosa * cosa, sina * sina
            coswt, sinwt = np.cos(wt), np.sin(wt)
            rX = (coswt * cos2a + sin2a) * X - sinwt * cosa * Y + (coswt * cosa * sina - cosa * sina) * Z
            rY = sinwt * cosa * X + coswt * Y + sina * sinwt * Z
            rZ = (coswt * sina * cosa - sina * cosa) * X - sinwt * sina * Y + (coswt * sin2a + cos2a) * Z
            X0, Y0, Z0 = sph2cart(np.pi / 4.0, np.pi / 6.0)
            return np.exp(-10.0 * ((rX - X0) ** 2 + (rY - Y0) ** 2 + (rZ - Z0) ** 2))
        X0, Y0, Z0 = sph2cart(0.0, 0.0)
        return np.exp(-10.0 * ((X - X0) ** 2 + (Y - Y0) ** 2 + (Z - Z0) ** 2))
    if sim.ic == 3:
        X, Y, Z = sph2cart(lon, lat)
        if sim.vf == 1:
            c1, c2 = (0, pi / 3.0), (0, -pi / 3.0)
        else:
            c1, c2 = (-pi / 6.0, 0), (pi / 6.0, 0)
        X1, Y1, Z1 = sph2cart(*c1)
        X2, Y2, Z2 = sph2cart(*c2)
        b0 = 5.0
        return np.exp(-b0 * ((X - X1) ** 2 + (Y - Y1) ** 2 + (Z - Z1) ** 2)) \
            + np.exp(-b0 * ((X - X2) ** 2 + (Y - Y2) ** 2 + (Z - Z2) ** 2))
    alpha = -45.0 * deg2rad                                 # ic == 4
    f = (-np.cos(lon) * np.cos(lat) * np.sin(alpha) + np.sin(lat) * np.cos(alpha))
    return 1.0 - f * f


def q_scalar_field(lon, lat, ic):
    """Analytic fields of the ghost-cell interpolation test (src/interpolation_test.py:55-71)."""
    if ic == 1:
        X0, Y0, Z0 = sph2cart(np.pi / 4.0, np.pi / 6.0)
        X, Y, Z = sph2cart(lon, lat)
        return np.exp(-10.0 * ((X - X0) ** 2 + (Y - Y0) ** 2 + (Z - Z0) ** 2))
    m = n = 1
    return (-np.cos(lon) * np.sin(m * lon) * m * np.cos(n * lat) ** 4 / np.cos(lat)
            - np.sin(lon) * np.cos(m * lon) * m ** 2 * np.cos(n * lat) ** 4 / np.cos(lat)
            + 12.0 * np.sin(lon) * np.cos(m * lon) * np.cos(n * lat) ** 2 * np.sin(n * lat) ** 2 * n ** 2 * np.cos(lat)
            - 4.0 * np.sin(lon) * np.cos(m * lon) * np.cos(n * lat) ** 4 * n ** 2 * np.cos(lat)
            + 4.0 * np.sin(lon) * np.cos(m * lon) * np.cos(n * lat) ** 3 * np.sin(n * lat) * n * np.sin(lat)) / np.cos(lat)


def init_vars_adv(g, sim):
    """src/advection_vars.py:19-107."""
    i0, iend, j0, jend = g.i0, g.iend, g.j0, g.jend
    P = g.N + g.ng
    sim.U_pu, sim.U_pv, sim.U_pc = wind.Velocity(P, "pu"), wind.Velocity(P, "pv"), wind.Velocity(P, "pc")
    iu = np.s_[i0:iend + 1, j0:jend, :]
    iv = np.s_[i0:iend, j0:jend + 1, :]
    for U, pos, idx in ((sim.U_pu, "pu", iu), (sim.U_pv, "pv", iv)):
        pts = getattr(g, pos)
        U.ulon[idx], U.vlat[idx] = wind.velocity_adv(pts.lon[idx], pts.lat[idx], 0.0, sim.vf)
        U.ucontra[idx], U.vcontra[idx] = wind.ll2contra(U.ulon[idx], U.vlat[idx], g, pos, idx)
    if g.projection == "gnomonic_equiangular":
        sim.tables = halo.lagrange_tables(g, sim.degree)
        sim.stencil_ghost_pc, sim.lagrange_poly_ghost_pc = sim.tables
    wind.ghost_fill_vector(sim.U_pu, sim.U_pv, sim.U_pc, g, sim)
    sim.U_pu.ucontra_old[...] = sim.U_pu.ucontra
    sim.U_pv.vcontra_old[...] = sim.U_pv.vcontra
    wind.time_averaged_velocity(g, sim)
    sim.cx = sim.U_pu.ucontra * sim.dt / g.dx               # cfl_x, src/cfl.py:10
    sim.cy = sim.U_pv.vcontra * sim.dt / g.dy
    sim.CFL = max(abs(np.amax(sim.cx[i0:iend + 1])), abs(np.amax(sim.cy[:, j0:jend + 1])))
    sim.px, sim.py = ppm.Parabola(P, "x"), ppm.Parabola(P, "y")
    ic = np.s_[i0:iend, j0:jend, :]
    sim.Q[ic] = qexact_adv(g.pc.lon[ic], g.pc.lat[ic], 0, sim)


def ghost_fill_scalar(Qx, Qy, g, sim):
    """src/edges_treatment.py:284-290."""
    if sim.et_name in ("ET-S72", "ET-PL07"):
        halo.copy_fill(Qx, Qy, g)
    else:
        halo.dg_fill(Qx, g, sim.tables)


def _line(par_x, par_y, name_lo, name_hi, key, g, ghost=False):
    """View of the boundary line of a parabola array for edge key=(dir,side,panel).

    ghost=False: the interior cell touching the edge; True: the ghost cell
    beyond it.  For 'lo' sides the facing edge value is name_lo, for 'hi'
    name_hi (pass the same name twice to address one array).
    """
    d, side, p = key
    par = par_x if d == "x" else par_y
    lo, hi = (g.i0, g.iend) if d == "x" else (g.j0, g.jend)
    if side == "lo":
        idx, name = (lo - 1 if ghost else lo), name_lo
    else:
        idx, name = (hi if ghost else hi - 1), name_hi
    arr = getattr(par, "_" + name)          # x-like storage: axis 0 is the sweep axis
    return arr[idx, lo:hi, p]


def edges_extrapolation(Qx, Qy, px, py, g, sim):
    """ET-PL07 one-sided edge values, averaging and ghost parabolas.

    src/edges_treatment.py:82-190 (non-overlapped projections).
    """
    for Qa, par, (lo, hi) in ((Qx, px, (g.i0, g.iend)),
                              (np.swapaxes(Qy, 0, 1), py, (g.j0, g.jend))):
        J = slice(lo, hi)                    # i0==j0, iend==jend
        qL, qR = par._q_L, par._q_R
        qL[lo, J] = 1.5 * Qa[lo, J] - 0.5 * Qa[lo + 1, J]                  # eq. 47 (:91-96)
        qR[hi - 1, J] = 1.5 * Qa[hi - 1, J] - 0.5 * Qa[hi - 2, J]
        qR[lo, J] = (3.0 * Qa[lo, J] + 11.0 * Qa[lo + 1, J]
                     - 2.0 * (Qa[lo + 2, J] - Qa[lo, J])) / 14.0           # eq. 49 (:100-109)
        qL[lo + 1, J] = qR[lo, J]
        qL[hi - 1, J] = (3.0 * Qa[hi - 1, J] + 11.0 * Qa[hi - 2, J]
                         - 2.0 * (Qa[hi - 3, J] - Qa[hi - 1, J])) / 14.0
        qR[hi - 2, J] = qL[hi - 1, J]
        if sim.recon_name == "PPM-PL07":                                    # :111-115
            qR[lo + 1, J] = qL[lo + 2, J]
            qL[hi - 2, J] = qR[hi - 3, J]
    # average the two panels' values on each cube edge (:31-76)
    for A, B, flip in CUBE_EDGES:
        a = _line(px, py, "q_L", "q_R", A, g)
        b = _line(px, py, "q_L", "q_R", B, g)
        a[...] = (a + (np.flip(b) if flip else b)) * 0.5
        b[...] = np.flip(a) if flip else a
    # ghost-cell parabolas next to every edge (:120-190); L/R swap when the two
    # panels' axes run against each other (same side on both panels)
    for A, B, flip in CUBE_EDGES:
        swap = A[1] == B[1]
        for dst, src in ((A, B), (B, A)):
            for name in ("q_L", "q_R"):
                other = {"q_L": "q_R", "q_R": "q_L"}[name] if swap else name
                s = _line(px, py, other, other, src, g)
                _line(px, py, name, name, dst, g, ghost=True)[...] = np.flip(s) if flip else s


def average_flux_cube_edges(px, py, g):
    """MF-AF (src/edges_treatment.py:231-278)."""
    def fl(key):
        d, side, p = key
        par = px if d == "x" else py
        lo, hi = (g.i0, g.iend) if d == "x" else (g.j0, g.jend)
        return par._f_upw[lo if side == "lo" else hi, lo:hi, p]
    a_, b_ = 0.5, 1.0 - 0.5
    for A, B, flip in CUBE_EDGES:
        sgn = -1.0 if A[1] == B[1] else 1.0
        a, b = fl(A), fl(B)
        bb = np.flip(b) if flip else b
        a[...] = a_ * a + sgn * (b_ * bb)
        b[...] = sgn * (np.flip(a) if flip else a)


def compute_fluxes(Qx, Qy, g, sim):
    """src/flux.py:9-15 + src/reconstruction_1d.py:387-394."""
    sw = lambda a: np.swapaxes(a, 0, 1)
    px, py = sim.px, sim.py
    ppm.reconstruct(Qx, px, sim.recon_name, g.i0, g.iend)
    ppm.reconstruct(sw(Qy), py, sim.recon_name, g.j0, g.jend)
    if sim.et_name == "ET-PL07":
        edges_extrapolation(Qx, Qy, px, py, g, sim)
    ppm.upwind_flux(Qx, px, sim.cx, sim.U_pu.ucontra_averaged, sim.U_pu.upos, sim.mt_name,
                    g.metric_tensor_pc, g.metric_tensor_pu, g.i0, g.iend)
    ppm.upwind_flux(sw(Qy), py, sw(sim.cy), sw(sim.U_pv.vcontra_averaged), sw(sim.U_pv.vpos),
                    sim.mt_name, sw(g.metric_tensor_pc), sw(g.metric_tensor_pv), g.j0, g.jend)


def divergence(g, sim):
    """src/discrete_operators.py:18-101."""
    i0, iend, j0, jend = g.i0, g.iend, g.j0, g.jend
    N, ng, dt = g.N, g.ng, sim.dt
    mt = g.metric_tensor_pc
    sim.gQ[...] = sim.Q * mt                                             # :31
    sim.cx[...] = sim.U_pu.ucontra_averaged * dt / g.dx                  # :34-35
    sim.cy[...] = sim.U_pv.vcontra_averaged * dt / g.dy
    compute_fluxes(sim.Q, sim.Q, g, sim)                                 # :38
    ppm.flux_difference(sim.px, dt, g.dx, i0, iend)                      # :42-43
    ppm.flux_difference(sim.py, dt, g.dy, j0, jend)
    pxdF, pydF = sim.px.dF, sim.py.dF
    gQ, Q = sim.gQ, sim.Q
    if sim.opsplit_name == "SP-AVLT":                                    # :49-51
        gQx = gQ + 0.5 * pxdF
        gQy = gQ + 0.5 * pydF
    else:
        c1x = g.metric_tensor_pu[1:] * sim.cx[1:]                        # :57-58 / :66-67
        c2x = g.metric_tensor_pu[:N + ng] * sim.cx[:N + ng]
        c1y = g.metric_tensor_pv[:, 1:] * sim.cy[:, 1:]
        c2y = g.metric_tensor_pv[:, :N + ng] * sim.cy[:, :N + ng]
        if sim.opsplit_name == "SP-L04":                                 # :59-60
            gQx = gQ + 0.5 * pxdF + 0.5 * (c1x - c2x) * Q
            gQy = gQ + 0.5 * pydF + 0.5 * (c1y - c2y) * Q
        else:                                                            # SP-PL07 :68-69
            Qx = 0.5 * (Q + (Q + pxdF) / (1.0 - (c1x - c2x)))
            Qy = 0.5 * (Q + (Q + pydF) / (1.0 - (c1y - c2y)))
    if sim.mt_name == "MT-0":                                            # :72-73
        Qx, Qy = gQx / mt, gQy / mt
    if sim.et_name in ("ET-S72", "ET-PL07"):                             # :76-78
        ghost_fill_scalar(Qx, Qy, g, sim)
    compute_fluxes(Qy, Qx, g, sim)                                       # :81 (swapped)
    if sim.mf_name == "MF-AF":                                           # :85-86
        average_flux_cube_edges(sim.px, sim.py, g)
    ppm.flux_difference(sim.px, dt, g.dx, i0, iend)                      # :89-90
    ppm.flux_difference(sim.py, dt, g.dy, j0, jend)
    sim.div[...] = -(sim.px.dF + sim.py.dF) / (dt * mt)                  # :95
    if sim.mf_name == "MF-PR":                                           # :98-101
        I = np.s_[i0:iend, j0:jend, :]
        m0 = np.sum(sim.div[I] * mt[I])
        a2 = np.sum(mt[I] * mt[I])
        sim.div[I] = sim.div[I] - mt[I] * m0 / a2


def adv_time_step(g, sim, k, t):
    """src/advection_timestep.py:19-43."""
    ghost_fill_scalar(sim.Q, sim.Q, g, sim)
    if sim.vf >= 2:
        wind.ghost_fill_vector(sim.U_pu, sim.U_pv, sim.U_pc, g, sim)
        wind.time_averaged_velocity(g, sim)
    divergence(g, sim)
    I = np.s_[g.i0:g.iend, g.j0:g.jend, :]
    sim.Q[I] = sim.Q[I] - sim.dt * sim.div[I]


def update_adv(g, sim, t):
    """src/advection_timestep.py:48-75."""
    if sim.vf < 2:
        return
    i0, iend, j0, jend = g.i0, g.iend, g.j0, g.jend
    iu = np.s_[i0:iend + 1, j0:jend, :]
    iv = np.s_[i0:iend, j0:jend + 1, :]
    U, V = sim.U_pu, sim.U_pv
    U.ulon[iu], U.vlat[iu] = wind.velocity_adv(g.pu.lon[iu], g.pu.lat[iu], t, sim.vf)
    V.ulon[iv], V.vlat[iv] = wind.velocity_adv(g.pv.lon[iv], g.pv.lat[iv], t, sim.vf)
    U.ucontra_old[...] = U.ucontra
    V.vcontra_old[...] = V.vcontra
    U.ucontra[...], U.vcontra[...] = wind.ll2contra(U.ulon, U.vlat, g, "pu")
    V.ucontra[...], V.vcontra[...] = wind.ll2contra(V.ulon, V.vlat, g, "pv")


def compute_errors(Q, Qref):
    """src/errors.py:99-113 (absolute Linf, mean-abs, RMS)."""
    E = abs(Qref - Q)
    n = E.size
    return np.amax(abs(E)), np.sum(E) / n, np.sqrt(np.sum(E * E) / n)


def mass_computation(Q, g, total_mass0):
    """src/diagnostics.py:14-26."""
    I = np.s_[g.i0:g.iend, g.j0:g.jend, :]
    total = np.sum(Q[I] * g.metric_tensor_pc[I] * g.dx * g.dy)
    if abs(total_mass0) > 10 ** (-10):
        return total, abs(total_mass0 - total) / abs(total_mass0)
    return total, abs(total_mass0 - total)


def run(g, sim, nsteps, t0_step=0):
    """adv_sphere's time loop without output (src/advection_sphere.py:45-57)."""
    for k in range(t0_step + 1, t0_step + nsteps + 1):
        t = k * sim.dt
        adv_time_step(g, sim, k, t)
        update_adv(g, sim, t)


def final_errors(g, sim, k):
    """The error block of output_adv (src/output.py:33-42) at step k."""
    I = np.s_[g.i0:g.iend, g.j0:g.jend, :]
    qe = qexact_adv(g.pc.lon[I], g.pc.lat[I], k * sim.dt, sim)
    return compute_errors(sim.Q[I], qe)
