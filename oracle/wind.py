"""Oracle: winds -- analytic fields, basis conversion, departure velocity, ghost fill.

  src/advection_ic.py:287-313   velocity_adv                 -> velocity_adv()
  src/sphgeo.py:120-133         latlon <-> contravariant      -> ll2contra(), contra2ll()
  src/averaged_velocity.py:14-62  time_averaged_velocity      -> time_averaged_velocity()
  src/interpolation.py:347-430  wind_edges2center_cubic_...   -> edges_to_centres()
  src/interpolation.py:436-532  wind_center2ghostedge_cubic_..-> centres_to_ghost_edges()
  src/edges_treatment.py:296-347 edges_ghost_cell_treatment_vector -> ghost_fill_vector()
"""
import numpy as np

from . import halo

pi = np.pi
deg2rad = 1.0 / (180.0 / pi)           # src/constants.py:23-24


class Velocity:
    """src/cs_datastruct.py:701-735."""

    def __init__(self, P, pos):
        shape = {"pu": (P + 1, P, 6), "pv": (P, P + 1, 6), "pc": (P, P, 6)}[pos]
        self.pos = pos
        for n in ("ulon", "vlat", "ucontra", "vcontra"):
            setattr(self, n, np.zeros(shape))
        if pos == "pu":
            self.ucontra_averaged = np.zeros(shape)
            self.ucontra_old = np.zeros(shape)
            self.upos = self.uneg = None
        elif pos == "pv":
            self.vcontra_averaged = np.zeros(shape)
            self.vcontra_old = np.zeros(shape)
            self.vpos = self.vneg = None


def velocity_adv(lon, lat, t, vf):
    """Analytic wind in lat-lon components (src/advection_ic.py:287-313)."""
    cos, sin = np.cos, np.sin
    if vf == 1:
        alpha = -45.0 * deg2rad
        u0 = 2.0 * pi / 5.0
        ulon = u0 * (cos(lat) * cos(alpha) + sin(lat) * cos(lon) * sin(alpha))
        vlat = -u0 * sin(lon) * sin(alpha)
    elif vf == 2:
        T, k = 5.0, 2.0
        lonp = lon - 2 * pi * t / T
        ulon = k * (sin((lonp + pi)) ** 2) * (sin(2. * lat)) * (cos(pi * t / T)) + 2. * pi * cos(lat) / T
        vlat = k * (sin(2 * (lonp + pi))) * (cos(lat)) * (cos(pi * t / T))
    elif vf == 3:
        T, k = 5.0, 1.0
        ulon = -k * (sin((lon + pi) / 2.0) ** 2) * (sin(2.0 * lat)) * (cos(lat) ** 2) * (cos(pi * t / T))
        vlat = (k / 2.0) * (sin((lon + pi))) * (cos(lat) ** 3) * (cos(pi * t / T))
    elif vf == 4:
        m = n = 1
        ulon = -m * (sin(lon) * sin(m * lon) * cos(n * lat) ** 3)
        vlat = -4 * n * (cos(n * lat) ** 3) * sin(n * lat) * cos(m * lon) * sin(lon)
    else:
        raise ValueError(vf)
    return ulon, vlat


def ll2contra(ulon, vlat, g, pos, idx=np.s_[:, :, :]):
    """src/sphgeo.py:120-125 with the coefficient arrays of `pos` restricted to idx."""
    exlon = getattr(g, "prod_ex_elon_" + pos)[idx]
    exlat = getattr(g, "prod_ex_elat_" + pos)[idx]
    eylon = getattr(g, "prod_ey_elon_" + pos)[idx]
    eylat = getattr(g, "prod_ey_elat_" + pos)[idx]
    det = getattr(g, "determinant_ll2contra_" + pos)[idx]
    u = eylat * ulon - eylon * vlat
    v = -exlat * ulon + exlon * vlat
    return u / det, v / det


def contra2ll(u, v, g, pos, idx=np.s_[:, :, :]):
    """src/sphgeo.py:130-133."""
    exlon = getattr(g, "prod_ex_elon_" + pos)[idx]
    exlat = getattr(g, "prod_ex_elat_" + pos)[idx]
    eylon = getattr(g, "prod_ey_elon_" + pos)[idx]
    eylat = getattr(g, "prod_ey_elat_" + pos)[idx]
    return exlon * u + eylon * v, exlat * u + eylat * v


def _averaged_1d(u, u_old, u_avg, i0, iend, dto2, dx, rk2):
    """x-like departure velocity along axis 0; returns the upwind mask.

    src/averaged_velocity.py:21-49.
    """
    upos = u[i0:iend + 1] >= 0
    if not rk2:
        u_avg[...] = u
        return upos
    uneg = ~upos
    ui = 1.5 * u - 0.5 * u_old                                 # :42
    a = u[i0:iend + 1] * dto2 / dx                             # :44
    u1, u2 = ui[i0 - 1:iend], ui[i0:iend + 1]
    u3, u4 = ui[i0:iend + 1], ui[i0 + 1:iend + 2]
    edges = u_avg[i0:iend + 1]
    edges[upos] = ((1.0 - a) * u2 + a * u1)[upos]             # :48
    edges[uneg] = (-a * u4 + (1.0 + a) * u3)[uneg]            # :49
    return upos


def time_averaged_velocity(g, sim):
    rk2 = sim.dp_name == "RK2"
    U, V = sim.U_pu, sim.U_pv
    U.upos = _averaged_1d(U.ucontra, U.ucontra_old, U.ucontra_averaged,
                          g.i0, g.iend, sim.dto2, g.dx, rk2)
    U.uneg = ~U.upos
    sw = lambda a: np.swapaxes(a, 0, 1)
    V.vpos = sw(_averaged_1d(sw(V.vcontra), sw(V.vcontra_old), sw(V.vcontra_averaged),
                             g.j0, g.jend, sim.dto2, g.dy, rk2))
    V.vneg = ~V.vpos


def edges_to_centres(U_pc, U_pu, U_pv, g, sim):
    """C-grid normal components -> both components at the centres of the
    boundary ring, then lat-lon, then duo-grid fill of ulon, vlat.

    src/interpolation.py:347-430.  The index expressions (including the ones
    that mix i- and j-names, valid because i0==j0 and iend==jend) are kept.
    """
    i0, iend, j0, jend, ngl = g.i0, g.iend, g.j0, g.jend, g.ngl
    a1, a2, a3, a4 = 5.0 / 16.0, 15.0 / 16.0, -5.0 / 16.0, 1.0 / 16.0
    b1, b2, b3, b4 = -1.0 / 16.0, 9.0 / 16.0, 9.0 / 16.0, -1.0 / 16.0
    u, v = U_pu.ucontra, U_pv.vcontra
    uc, vc = U_pc.ucontra, U_pc.vcontra
    J = slice(j0, jend)
    I = slice(i0, iend)
    # west (:359-372)
    uc[i0, J] = a1 * u[i0, J] + a2 * u[i0 + 1, J] + a3 * u[i0 + 2, J] + a4 * u[i0 + 3, J]
    uc[i0 + 1:i0 + ngl, J] = b1 * u[i0:i0 - 1 + ngl, J] + b2 * u[i0 + 1:i0 + ngl, J] \
        + b3 * u[i0 + 2:i0 + 1 + ngl, J] + b4 * u[i0 + 3:i0 + 2 + ngl, J]
    Jm = slice(j0 + ngl, jend - ngl)
    vc[i0:i0 + ngl, Jm] = b1 * v[i0:i0 + ngl, j0 + ngl - 1:jend - ngl - 1] \
        + b2 * v[i0:i0 + ngl, j0 + ngl:jend - ngl] \
        + b3 * v[i0:i0 + ngl, j0 + ngl + 1:jend - ngl + 1] \
        + b4 * v[i0:i0 + ngl, j0 + ngl + 2:jend - ngl + 2]
    # east (:374-388)
    uc[iend - 1, J] = a4 * u[iend - 3, J] + a3 * u[iend - 2, J] + a2 * u[iend - 1, J] + a1 * u[iend, J]
    uc[iend - ngl:iend - 1, J] = b4 * u[iend - ngl - 1:iend - 2, J] + b3 * u[iend - ngl:iend - 1, J] \
        + b2 * u[iend - ngl + 1:iend, J] + b1 * u[iend - ngl + 2:iend + 1, J]
    vc[iend - ngl:iend, Jm] = b1 * v[iend - ngl:iend, j0 + ngl - 1:jend - ngl - 1] \
        + b2 * v[iend - ngl:iend, j0 + ngl:jend - ngl] \
        + b3 * v[iend - ngl:iend, j0 + ngl + 1:jend - ngl + 1] \
        + b4 * v[iend - ngl:iend, j0 + ngl + 2:jend - ngl + 2]
    # south (:390-404)
    vc[I, j0] = a1 * v[I, j0] + a2 * v[I, j0 + 1] + a3 * v[I, j0 + 2] + a4 * v[I, j0 + 3]
    vc[I, j0 + 1:j0 + ngl] = b1 * v[I, j0:j0 - 1 + ngl] + b2 * v[I, j0 + 1:j0 + ngl] \
        + b3 * v[I, j0 + 2:j0 + 1 + ngl] + b4 * v[I, j0 + 3:j0 + 2 + ngl]
    Im = slice(i0 + ngl, iend - ngl)
    uc[Im, j0:j0 + ngl] = b1 * u[i0 + ngl - 1:iend - ngl - 1, j0:j0 + ngl] \
        + b2 * u[i0 + ngl:iend - ngl, j0:j0 + ngl] \
        + b3 * u[i0 + ngl + 1:iend - ngl + 1, j0:j0 + ngl] \
        + b4 * u[i0 + ngl + 2:iend - ngl + 2, j0:j0 + ngl]
    # north (:406-418)
    vc[I, jend - 1] = a4 * v[I, jend - 3] + a3 * v[I, jend - 2] + a2 * v[I, jend - 1] + a1 * v[I, jend]
    vc[I, jend - ngl:jend - 1] = b4 * v[I, jend - ngl - 1:jend - 2] + b3 * v[I, jend - ngl:jend - 1] \
        + b2 * v[I, jend - ngl + 1:jend] + b1 * v[I, jend - ngl + 2:jend + 1]
    uc[Im, jend - ngl:jend] = b1 * u[i0 + ngl - 1:iend - ngl - 1, jend - ngl:jend] \
        + b2 * u[i0 + ngl:iend - ngl, jend - ngl:jend] \
        + b3 * u[i0 + ngl + 1:iend - ngl + 1, jend - ngl:jend] \
        + b4 * u[i0 + ngl + 2:iend - ngl + 2, jend - ngl:jend]
    # to lat-lon on the interior (:422-425), then ghost centres (:427-430)
    idx = np.s_[i0:iend, j0:jend, :]
    U_pc.ulon[idx], U_pc.vlat[idx] = contra2ll(uc[idx], vc[idx], g, "pc", idx)
    halo.dg_fill(U_pc.ulon, g, sim.tables)
    halo.dg_fill(U_pc.vlat, g, sim.tables)


def centres_to_ghost_edges(U_pc, U_pu, U_pv, g):
    """src/interpolation.py:436-532."""
    i0, iend, j0, jend = g.i0, g.iend, g.j0, g.jend
    a1, a2 = 9.0 / 16.0, -1.0 / 16.0

    def conv(U, pos, idx):
        U.ucontra[idx], U.vcontra[idx] = ll2contra(U.ulon[idx], U.vlat[idx], g, pos, idx)

    # pu ghost rows south / north (:443-463): edges i0..iend
    for js in (np.s_[:j0], np.s_[jend:]):
        for n in ("ulon", "vlat"):
            c = getattr(U_pc, n)
            getattr(U_pu, n)[i0:iend + 1, js] = a1 * (c[i0:iend + 1, js] + c[i0 - 1:iend, js]) \
                + a2 * (c[i0 + 1:iend + 2, js] + c[i0 - 2:iend - 1, js])
        conv(U_pu, "pu", np.s_[i0:iend + 1, js, :])
    # pv ghost columns west / east (:469-487): edges j0..jend
    for is_ in (np.s_[:i0], np.s_[iend:]):
        for n in ("ulon", "vlat"):
            c = getattr(U_pc, n)
            getattr(U_pv, n)[is_, j0:jend + 1] = a1 * (c[is_, j0:jend + 1] + c[is_, j0 - 1:jend]) \
                + a2 * (c[is_, j0 + 1:jend + 2] + c[is_, j0 - 2:jend - 1])
        conv(U_pv, "pv", np.s_[is_, j0:jend + 1, :])
    # the extra lines the RK2 departure point reads (:493-532)
    for n in ("ulon", "vlat"):
        c = getattr(U_pc, n)
        getattr(U_pu, n)[i0 - 1] = a1 * (c[i0 - 2] + c[i0 - 1]) + a2 * (c[i0] + c[i0 - 3])
    conv(U_pu, "pu", np.s_[i0 - 1, :, :])
    for n in ("ulon", "vlat"):
        c = getattr(U_pc, n)
        getattr(U_pu, n)[iend + 1] = a1 * (c[iend] + c[iend + 1]) + a2 * (c[iend - 1] + c[iend + 2])
    conv(U_pu, "pu", np.s_[iend + 1, :, :])
    for n in ("ulon", "vlat"):
        c = getattr(U_pc, n)
        getattr(U_pv, n)[:, j0 - 1] = a1 * (c[:, j0 - 2] + c[:, j0 - 1]) + a2 * (c[:, j0] + c[:, j0 - 3])
    conv(U_pv, "pv", np.s_[:, j0 - 1, :])
    for n in ("ulon", "vlat"):
        c = getattr(U_pc, n)
        # the reference writes iend+2 in the vlat line (:526); iend == jend
        getattr(U_pv, n)[:, jend + 1] = a1 * (c[:, jend] + c[:, jend + 1]) + a2 * (c[:, jend - 1] + c[:, jend + 2])
    conv(U_pv, "pv", np.s_[:, jend + 1, :])


# (dst field, dst panel, line, src field, src panel, line, sign, flip) of
# src/edges_treatment.py:311-347; 'u' lines are ucontra[idx, j0:jend], 'v'
# lines are vcontra[i0:iend, idx].  Offsets are relative to i0/iend resp. j0/jend.
_RK2_COPIES = (
    ("u", 0, ("hi", 1), "u", 1, ("lo", 1), 1, 0), ("u", 1, ("hi", 1), "u", 2, ("lo", 1), 1, 0),
    ("u", 2, ("hi", 1), "u", 3, ("lo", 1), 1, 0), ("u", 3, ("hi", 1), "u", 0, ("lo", 1), 1, 0),
    ("u", 1, ("lo", -1), "u", 0, ("hi", -1), 1, 0), ("u", 2, ("lo", -1), "u", 1, ("hi", -1), 1, 0),
    ("u", 3, ("lo", -1), "u", 2, ("hi", -1), 1, 0), ("u", 0, ("lo", -1), "u", 3, ("hi", -1), 1, 0),
    ("v", 0, ("hi", 1), "v", 4, ("lo", 1), 1, 0), ("v", 4, ("lo", -1), "v", 0, ("hi", -1), 1, 0),
    ("u", 4, ("hi", 1), "v", 1, ("hi", -1), -1, 0), ("v", 1, ("hi", 1), "u", 4, ("hi", -1), -1, 0),
    ("v", 4, ("hi", 1), "v", 2, ("hi", -1), -1, 1), ("v", 2, ("hi", 1), "v", 4, ("hi", -1), -1, 1),
    ("u", 4, ("lo", -1), "v", 3, ("hi", -1), 1, 1), ("v", 3, ("hi", 1), "u", 4, ("lo", 1), 1, 1),
    ("v", 5, ("hi", 1), "v", 0, ("lo", 1), 1, 0), ("v", 0, ("lo", -1), "v", 5, ("hi", -1), 1, 0),
    ("v", 1, ("lo", -1), "u", 5, ("hi", -1), 1, 1), ("u", 5, ("hi", 1), "v", 1, ("lo", 1), 1, 1),
    ("v", 2, ("lo", -1), "v", 5, ("lo", 1), -1, 1), ("v", 5, ("lo", -1), "v", 2, ("lo", 1), -1, 1),
    ("v", 3, ("lo", -1), "u", 5, ("lo", 1), -1, 0), ("u", 5, ("lo", -1), "v", 3, ("lo", 1), -1, 0),
)


def ghost_fill_vector(U_pu, U_pv, U_pc, g, sim):
    """src/edges_treatment.py:296-347."""
    if sim.et_name == "ET-DG":
        edges_to_centres(U_pc, U_pu, U_pv, g, sim)
        centres_to_ghost_edges(U_pc, U_pu, U_pv, g)
    elif sim.dp_name == "RK2":
        i0, iend, j0, jend = g.i0, g.iend, g.j0, g.jend

        def line(f, p, where):
            base = (i0 if where[0] == "lo" else iend) + where[1]
            if f == "u":
                return U_pu.ucontra[base, j0:jend, p]
            return U_pv.vcontra[i0:iend, base, p]

        # the reference applies the copies sequentially, in this order
        for df, dp, dw, sf, sp, sw, sign, flip in _RK2_COPIES:
            src = line(sf, sp, sw)
            if flip:
                src = np.flip(src)
            line(df, dp, dw)[...] = src if sign > 0 else -src
