"""Diagnostics and data files of output_adv (src/output.py:22-184); the plots are out of scope.

What a run leaves behind in the reference besides figures -- and what this module writes when the
caller passes a lat-lon grid (`ll_grid`, with `ix, jy, mask` from interpolation.ll2cs):
  data/<grid name>_adv_Q_error_ic*_vf*_<scheme names>_interp<d>.npy   error field on the lat-lon grid
  data/<grid name>_adv_Q_error_ic*_vf*_<scheme names>_interp<d>        seven numbers, one per line:
        Linf, L1, L2 error, CFL, mass change, dt, Tf                    (src/output.py:136-150)
"""
import os

import numpy as np

from .constants import datadir
from .cs_datastruct import scalar_field
from .errors import compute_errors
from .advection_ic import qexact_adv, div_exact
from .diagnostics import mass_computation
from .interpolation import nearest_neighbour


def print_diagnostics_adv(error_linf, error_l1, error_l2, mass_change, t, Nsteps):
    print('\nStep', t, 'from', Nsteps)
    print('Error (Linf, L1, L2) :', "{:.2e}".format(error_linf), "{:.2e}".format(error_l1), "{:.2e}".format(error_l2))
    print('Total mass variation:', "{:.2e}".format(mass_change))


def _scheme_tag(simulation, with_mf):
    names = [simulation.opsplit_name, simulation.recon_name, simulation.dp_name, simulation.et_name, simulation.mt_name]
    if with_mf:
        names.append(simulation.mf_name)
    return "_".join(names)


def save_error_files(cs_grid, ll_grid, simulation, q, q_exact, k):
    """The .npy / text pair of src/output.py:136-150 for the final step."""
    fq, fe = scalar_field(cs_grid, 'q', 'center'), scalar_field(cs_grid, 'q_exact', 'center')
    fq.f[:, :, :], fe.f[:, :, :] = q, q_exact
    err_ll = nearest_neighbour(fe, cs_grid, ll_grid) - nearest_neighbour(fq, cs_grid, ll_grid)
    base = datadir + cs_grid.name + '_adv_Q_error_ic' + str(simulation.ic) + '_vf' + str(simulation.vf) + '_' + \
        _scheme_tag(simulation, True) + '_interp' + str(simulation.degree)
    os.makedirs(os.path.dirname(base) or ".", exist_ok=True)
    np.save(base, err_ll, allow_pickle=False)
    np.savetxt(base, np.array([simulation.error_linf[k], simulation.error_l1[k], simulation.error_l2[k],
                               simulation.CFL, simulation.mass_change, simulation.dt, simulation.Tf]))
    return base


def errors_device(simulation, q_exact):
    """compute_errors (src/errors.py:99-113) of the device-resident Q against a host exact field (N,N,6)."""
    import ctypes as C
    qe = np.ascontiguousarray(q_exact, dtype=np.float64)
    out = (C.c_double * 3)()
    dp = C.POINTER(C.c_double)
    simulation.dev.call("pycs_errors", qe.ctypes.data_as(dp), C.cast(out, dp))
    return out[0], out[1], out[2]


def errors_exact_device(simulation, t):
    """The same against qexact_adv(t) evaluated on the device (csrc/grid.cu)."""
    import ctypes as C
    out = (C.c_double * 3)()
    simulation.dev.call("pycs_errors_exact", float(t), C.cast(out, C.POINTER(C.c_double)))
    return out[0], out[1], out[2]


def output_adv(cs_grid, ll_grid, simulation, plot, k, t, Nsteps, plotstep, map_projection, divtest_flag):
    if plot or k == Nsteps:
        i0, iend, j0, jend = cs_grid.i0, cs_grid.iend, cs_grid.j0, cs_grid.jend
        I = np.s_[i0:iend, j0:jend, :]
        lean = getattr(cs_grid, "lean", False)
        sharded = simulation.dev.row_range() != (i0, iend)
        if lean and not sharded:
            # exact solution and the three norms on the device: nothing is transferred
            simulation.error_linf[k], simulation.error_l1[k], simulation.error_l2[k] = errors_exact_device(simulation, t)
            q = q_exact = None
        else:
            q_exact = qexact_adv(cs_grid.pc.lon[I], cs_grid.pc.lat[I], t, simulation)
            if sharded:
                q = simulation.Q[I]
                simulation.error_linf[k], simulation.error_l1[k], simulation.error_l2[k] = compute_errors(q, q_exact)
            else:
                # |Q - Qexact| reduced on the device (csrc/ppm.cu k_errors); only the exact field goes up
                simulation.error_linf[k], simulation.error_l1[k], simulation.error_l2[k] = \
                    errors_device(simulation, q_exact)
                q = None
        if simulation.error_linf[k] > 100.0:
            print('Stopping due to large errors.')
            print('The CFL number is:', simulation.CFL)
            raise SystemExit(1)
        simulation.total_mass, simulation.mass_change = \
            mass_computation(simulation.Q, cs_grid, simulation.total_mass0)
        if k > 0 and (not divtest_flag):
            print_diagnostics_adv(simulation.error_linf[k], simulation.error_l1[k], simulation.error_l2[k],
                                  simulation.mass_change, k, Nsteps)
        if k == 0 or k == Nsteps or (plotstep > 0 and k % plotstep == 0):
            if divtest_flag:
                # divergence test: the norms are those of div against the exact divergence
                # (src/output.py:152-169)
                d = np.asarray(simulation.div)[I]
                dex = div_exact(cs_grid.pc.lon[I], cs_grid.pc.lat[I], simulation)
                simulation.error_linf[k], simulation.error_l1[k], simulation.error_l2[k] = compute_errors(d, dex)
            elif t > 0 and k == Nsteps and ll_grid is not None and len(np.shape(ll_grid.mask)) == 2 and q_exact is not None:
                save_error_files(cs_grid, ll_grid, simulation, simulation.Q[I] if q is None else q, q_exact, k)
