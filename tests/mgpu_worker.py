"""Worker of the multi-GPU parity test: launched with torchrun, one rank per GPU.
Each rank runs the sharded fused step and, on the same device, an unsharded replica, and
compares its own slab (<= 1e-13 of max|Q|; the MF-PR sum is reduced in a different order)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pycs_b200  # noqa: E402,F401
from pycs_b200 import cs_datastruct, advection_ic, advection_vars, advection_timestep, parallel  # noqa: E402
from golden_common import TUPLES, DT16  # noqa: E402


def make(g, vf, tup, dev):
    dt = DT16[vf] * 16 / g.N
    sim = advection_ic.adv_simulation_par(g, dt, 5, 2, vf, 1, *tup, device=dev)
    advection_vars.init_vars_adv(g, sim)
    return sim


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    cases = [(48, 1, "default", (1, 4, 7)), (130, 3, "default", (2, 9)), (130, 2, "AVLT-RK2-DG-PR", (3, 5)),
             (64, 1, "PL07-RK1-DG-PR", (6,)), (384, 3, "default", (10,)), (1536, 3, "default", (3, 4))]
    if len(sys.argv) > 1:
        cases = [(int(sys.argv[1]), 3, "default", (10,))]
    worst = 0.0
    # the multi-GPU step is the one-kernel step over the CTA table (boundary CTAs first, in-kernel ghost fill and
    # exchange), replayed from CUDA graphs (default) or launched directly; "split": the two-stream split step
    # (PYCS_ONEKERNEL=0) it replaced
    for impl in ("graph", "nograph", "split"):
        os.environ["PYCS_GRAPH"] = "1" if impl == "graph" else "0"
        os.environ["PYCS_ONEKERNEL"] = "0" if impl == "split" else "1"
        for N, vf, name, calls in (cases if impl != "split" else cases[:3]):
            g = cs_datastruct.cubed_sphere(N)
            a = make(g, vf, TUPLES[name], local)
            b = make(g, vf, TUPLES[name], local)
            lo, hi = parallel.shard(a)
            k = 0
            for n in calls:
                advection_timestep.run_steps(g, a, k, n, fused=True)
                advection_timestep.run_steps(g, b, k, n, fused=True)
                k += n
            qa, qb = np.asarray(a.Q), np.asarray(b.Q)
            err = float(np.max(np.abs(qa[lo:hi, 4:N + 4] - qb[lo:hi, 4:N + 4])) / np.max(np.abs(qb)))
            full = parallel.gather_field(a)
            errf = float(np.max(np.abs(full[4:N + 4, 4:N + 4] - qb[4:N + 4, 4:N + 4])) / np.max(np.abs(qb)))
            worst = max(worst, err, errf)
            print("rank %d/%d impl=%s N=%d vf=%d %s rows [%d,%d): slab err %.2e, gathered err %.2e"
                  % (rank, world, impl, N, vf, name, lo, hi, err, errf), flush=True)
            if N <= 384:
                # the host-buffer step on a sharded handle moves only the rank's slab (pycs_adv_time_step_host)
                import ctypes as C
                Qa = np.ascontiguousarray(qb.copy())
                Qb = np.ascontiguousarray(qb.copy())
                Qa[:lo] = np.nan                       # rows of other ranks are neither read nor written
                Qa[hi:] = np.nan
                dp = C.POINTER(C.c_double)
                for kk in range(k + 1, k + 4):
                    a.dev.call("pycs_adv_time_step_host", Qa.ctypes.data_as(dp), kk, kk * a.dt, 1)
                    b.dev.call("pycs_adv_time_step_host", Qb.ctypes.data_as(dp), kk, kk * b.dt, 1)
                errh = float(np.max(np.abs(Qa[lo:hi, 4:N + 4] - Qb[lo:hi, 4:N + 4])) / np.max(np.abs(Qb[4:N + 4, 4:N + 4])))
                ok_nan = bool(np.all(np.isnan(Qa[:lo])) and np.all(np.isnan(Qa[hi:])))
                worst = max(worst, errh, 0.0 if ok_nan else 1.0)
                print("rank %d/%d impl=%s N=%d host-buffer steps: slab err %.2e, other rows untouched: %s"
                      % (rank, world, impl, N, errh, ok_nan), flush=True)
            a.dev.call("pycs_synchronize")
            dist.barrier()
            a.dev.close()
            b.dev.close()
    t = torch.tensor([worst], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = float(t.item()) <= 1e-13
    if rank == 0:
        print("MGPU %s worst=%.3e" % ("OK" if ok else "FAIL", float(t.item())), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
