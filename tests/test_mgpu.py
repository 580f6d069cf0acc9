"""Multi-GPU tests.  CPU part: the exchange plan (host-only C ABI call) against the oracle's ghost fill, and the
torch.distributed rendezvous helper under gloo, world_size 2.  GPU part: the sharded fused
step against an unsharded replica, torchrun with 2 ranks (skipped on a 1-GPU box)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tables(N):
    from oracle.grid import LeanGrid
    from oracle import halo as ohalo
    og = LeanGrid.centres_only(N)
    return og, ohalo, ohalo.lagrange_tables(og, 3)


@pytest.mark.parametrize("N,world", [(16, 2), (48, 4), (50, 3), (130, 8), (96, 8)])
def test_plan_delivers_every_cell_a_rank_reads(N, world):
    """The exchange plan against the oracle's ghost fill: a rank that holds ONLY its own rows plus the
    rectangles the plan delivers (everything else NaN) must end up, after the reference's two-phase
    Lagrange fill, with exactly the values of the whole-sphere fill on every cell its step reads: the rows
    of its slab +- 3 (with the W / E ghost rows at the ends), all columns, ghost cells and corners included."""
    import pycs_b200  # noqa: F401
    from pycs_b200.device import mgpu_plan
    og, ohalo, tables = _tables(N)
    kminE = tables[0][0][0]
    lo, hi, P = 4, N + 4, N + 8
    plans = [mgpu_plan(N, world, r, kminE, 3) for r in range(world)]
    assert plans[0][0] == lo and plans[-1][1] == hi
    for r in range(world - 1):
        assert plans[r][1] == plans[r + 1][0]
    rng = np.random.default_rng(N * 10 + world)
    full = np.zeros((P, P, 6))
    full[lo:hi, lo:hi, :] = rng.standard_normal((N, N, 6))
    want = full.copy()
    ohalo.dg_fill(want, og, tables)
    sent = 0
    for r, (a, b, _) in enumerate(plans):
        mine = np.full((P, P, 6), np.nan)
        mine[a:b, lo:hi, :] = full[a:b, lo:hi, :]
        for s, (sa, sb, rects) in enumerate(plans):
            for peer, panel, i0, i1, j0, j1 in rects:
                assert peer != s and sa <= i0 < i1 <= sb and lo <= j0 < j1 <= hi      # a rank only sends its own cells
                if peer == r:
                    mine[i0:i1, j0:j1, panel] = full[i0:i1, j0:j1, panel]
                    sent += (i1 - i0) * (j1 - j0) if r == 0 else 0
        ohalo.dg_fill(mine, og, tables)
        ia = 0 if a - 3 < lo else a - 3
        ib = P if b + 3 > hi else b + 3
        got, exp = mine[ia:ib], want[ia:ib]
        assert not np.any(np.isnan(got)), (r, np.argwhere(np.isnan(got))[:5])
        assert np.array_equal(got, exp), r
    # nothing is broadcast: what rank 0 receives is a small multiple of its halo rows + strip pieces
    assert sent <= 6 * (3 * N + 16 * (N // world + 16) + 8 * N + 64 * world)


def test_plan_is_targeted_at_bench_size():
    """N = 1536 on 8 ranks: a few hundred rectangles in total, each rank sends a few hundred KB."""
    import pycs_b200  # noqa: F401
    from pycs_b200.device import mgpu_plan
    og, ohalo, tables = _tables(1536)
    kminE = tables[0][0][0]
    for r in (0, 3, 7):
        a, b, rects = mgpu_plan(1536, 8, r, kminE, 3)
        cells = sum((i1 - i0) * (j1 - j0) for _, _, i0, i1, j0, j1 in rects)
        assert len(rects) <= 400, len(rects)
        assert cells * 8 <= 1.5e6, cells


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
