"""Locus sharding across the GPUs of one box.

Loci are independent (reference: every per-locus object lives inside the region loop,
src/bam_processor.cpp:563-627), so the multi-GPU plan is: one process per GPU, a contiguous
chunk of the position-sorted locus list per rank (keeps the output ordered, as the reference's
VCFWriter heap expects in-order arrival: src/vcf_writer.cpp:7-36), NO data-path collective, and
a host-side gather of the per-locus results in locus order on rank 0.
"""
import numpy as np


def shard_range(n_items, rank, world, weights=None):
    """Contiguous [begin, end) of rank's shard.  With ``weights`` (per-item cost, e.g. DP cells)
    the cut points balance the summed weight instead of the item count."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    if weights is None:
        base, rem = divmod(n_items, world)
        begin = rank * base + min(rank, rem)
        return begin, begin + base + (1 if rank < rem else 0)
    w = np.asarray(weights, dtype=np.float64)
    if len(w) != n_items:
        raise ValueError("weights length")
    c = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [int(np.searchsorted(c, c[-1] * r / world, side="left")) for r in range(world + 1)]
    cuts[0], cuts[-1] = 0, n_items
    for r in range(1, world + 1):
        cuts[r] = max(cuts[r], cuts[r - 1])
    return cuts[rank], cuts[rank + 1]


def gather_in_locus_order(local, group=None, dst=0):
    """Host-side gather of per-rank result arrays (rank r holds the results of its contiguous locus
    chunk) into one array in locus order on ``dst``; other ranks get None.  Works on any backend
    (``gloo`` for the CPU tests; the data path never needs a collective)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return np.asarray(local)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    a = np.ascontiguousarray(local)
    parts = [None] * world if rank == dst else None
    dist.gather_object((a.dtype.str, a.shape, a.tobytes()), parts, dst=dst, group=group)
    if rank != dst:
        return None
    arrs = [np.frombuffer(b, dtype=np.dtype(d)).reshape(s) for d, s, b in parts]
    return np.concatenate(arrs, axis=0)
