"""Multi-GPU tests.  CPU part: the decomposition plan (host-only C ABI call) and the
torch.distributed rendezvous helper under gloo, world_size 2.  GPU part: the sharded fused
step against an unsharded replica, torchrun with 2 ranks (skipped on a 1-GPU box)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("N,world", [(16, 2), (48, 4), (50, 3), (1536, 8), (130, 8)])
def test_plan_covers_what_every_rank_reads(N, world):
    """Own rows + received rectangles must contain: own rows, 3 halo rows on each side, and the
    four 4-wide boundary strips of the panel (the sources of the Lagrange ghost fill)."""
    import pycs_b200  # noqa: F401
    from pycs_b200.device import mgpu_plan
    lo, hi = 4, N + 4
    plans = [mgpu_plan(N, world, r) for r in range(world)]
    assert plans[0][0] == lo and plans[-1][1] == hi
    for r in range(world - 1):
        assert plans[r][1] == plans[r + 1][0]
    have = [np.zeros((N + 8, N + 8), bool) for _ in range(world)]
    for r, (a, b, jobs) in enumerate(plans):
        have[r][a:b, lo:hi] = True
    for r, (a, b, jobs) in enumerate(plans):
        for peer, i0, i1, j0, j1 in jobs:
            assert peer != r and a <= i0 < i1 <= b and lo <= j0 < j1 <= hi      # a rank only sends its own cells
            have[peer][i0:i1, j0:j1] = True
    for r, (a, b, jobs) in enumerate(plans):
        need = np.zeros((N + 8, N + 8), bool)
        need[max(a - 3, lo):min(b + 3, hi), lo:hi] = True
        need[lo:hi, lo:lo + 4] = True
        need[lo:hi, hi - 4:hi] = True
        need[lo:lo + 4, lo:hi] = True
        need[hi - 4:hi, lo:hi] = True
        assert not np.any(need & ~have[r]), r


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import pycs_b200  # noqa: F401
    from pycs_b200 import parallel
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    blocks = parallel.all_gather_bytes(bytes([rank]) * 192)
    q.put((rank, [b[0] for b in blocks], [len(b) for b in blocks]))
    dist.destroy_process_group()


def test_handle_rendezvous_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, firsts, lens in res:
        assert firsts == [0, 1] and lens == [192, 192]


@pytest.mark.gpu
def test_sharded_step_matches_single_gpu():
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", "29641", os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1200)
    sys.stdout.write(r.stdout[-4000:])
    sys.stderr.write(r.stderr[-4000:])
    assert r.returncode == 0 and "MGPU OK" in r.stdout
