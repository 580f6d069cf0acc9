"""Multi-GPU plumbing (one process per GPU): rendezvous helpers around the C ABI's
pycs_mgpu_* entry points.  The reference is single process; the data path
between GPUs is peer-mapped stores inside the library (csrc/mgpu.cu) -- torch.distributed is
used only to exchange the 192-byte IPC handle blocks and for barriers / timing reductions.
"""
import numpy as np


def all_gather_bytes(b):
    """Every rank's byte string, in rank order (torch.distributed, any backend)."""
    import torch.distributed as dist
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, bytes(b))
    return out


def shard(simulation):
    """Shard the fused step of `simulation` over the ranks of the default process group.
    Returns this rank's rows (row_lo, row_hi) in padded-panel index."""
    import torch.distributed as dist
    return simulation.dev.mgpu_setup(dist.get_rank(), dist.get_world_size(), all_gather_bytes)


def own_rows(simulation, field="Q"):
    """This rank's valid slab of a centre field: (row_lo, row_hi, array[row_lo:row_hi])."""
    a, b = simulation.dev.row_range()
    return a, b, np.asarray(getattr(simulation, field))[a:b]


def gather_field(simulation, field="Q"):
    """Assemble the full field on every rank from the ranks' slabs (host side; for output
    and tests, not on the step path)."""
    import torch.distributed as dist
    a, b, mine = own_rows(simulation, field)
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, (a, b, mine))
    full = np.asarray(getattr(simulation, field)).copy()
    for lo, hi, rows in parts:
        full[lo:hi] = rows
    return full
