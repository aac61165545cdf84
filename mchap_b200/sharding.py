"""Partition of loci over ranks and host-side gather of per-locus results.

The reference's only parallelism is a process pool over ``np.array_split(loci, n_cores)``
(mchap/application/baseclass.py:360-388) with a writer draining results in arrival order.  Here a
rank owns one GPU and a contiguous block of loci; nothing crosses GPUs on the data path — results
are gathered on the host in locus order.
"""
import numpy as np


def locus_block(n_loci, rank, world_size):
    """Half-open [start, stop) of the contiguous block of loci owned by ``rank`` — identical to the
    sizes ``np.array_split(np.arange(n_loci), world_size)`` produces."""
    base, extra = divmod(int(n_loci), int(world_size))
    start = rank * base + min(rank, extra)
    stop = start + base + (1 if rank < extra else 0)
    return start, stop


def gather_by_locus(local_results, n_loci, group=None):
    """All ranks pass the list of results of their block (in local locus order); every rank gets
    the full list in global locus order.  Uses torch.distributed object gather (host side)."""
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        assert len(local_results) == n_loci
        return list(local_results)
    world = dist.get_world_size(group)
    parts = [None] * world
    dist.all_gather_object(parts, list(local_results), group=group)
    out = []
    for r, part in enumerate(parts):
        start, stop = locus_block(n_loci, r, world)
        assert len(part) == stop - start, "rank %d returned %d results for %d loci" % (r, len(part), stop - start)
        out.extend(part)
    return out
