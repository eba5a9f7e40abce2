"""CPU, world_size 2 over gloo: locus sharding + host-side ordered gather.
The per-rank 'hot path' here is the oracle (no GPU in this container); what is under test is the
partitioning / gather logic that bench.py and the multi-GPU driver use."""
import os
import socket
import sys

import numpy as np
import pytest

from longtr_b200 import shard


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 100):
        for w in (1, 2, 3, 8):
            cuts = [shard.shard_range(n, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
            assert max(e - b for b, e in cuts) - min(e - b for b, e in cuts) <= 1


def test_weighted_shards_balance_cost():
    rng = np.random.default_rng(1)
    w = rng.integers(1, 1000, size=500).astype(float)
    w[:20] *= 50
    tot = [w[slice(*shard.shard_range(500, r, 4, w))].sum() for r in range(4)]
    assert max(tot) < 1.35 * (w.sum() / 4)
    cuts = [shard.shard_range(500, r, 4, w) for r in range(4)]
    assert cuts[0][0] == 0 and cuts[-1][1] == 500 and all(cuts[i][1] == cuts[i + 1][0] for i in range(3))


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist
    import synth
    from oracle import pyoracle as po
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b = synth.make_pair_batch(4242, n_loci=21, n_lo=20, n_hi=80)
        lhb, lrb = b["locus_hap_begin"].astype(np.int64), b["locus_read_begin"].astype(np.int64)
        per_locus = (lhb[1:] - lhb[:-1]) * (lrb[1:] - lrb[:-1])
        l0, l1 = shard.shard_range(21, rank, world, weights=per_locus)
        sub = dict(locus_hap_begin=(b["locus_hap_begin"][l0:l1 + 1] - b["locus_hap_begin"][l0]).astype(np.uint32),
                   locus_read_begin=(b["locus_read_begin"][l0:l1 + 1] - b["locus_read_begin"][l0]).astype(np.uint32),
                   hap_off=(b["hap_off"][lhb[l0]:lhb[l1] + 1] - b["hap_off"][lhb[l0]]).astype(np.uint32),
                   read_off=(b["read_off"][lrb[l0]:lrb[l1] + 1] - b["read_off"][lrb[l0]]).astype(np.uint32),
                   hap_bytes=b["hap_bytes"][b["hap_off"][lhb[l0]]:b["hap_off"][lhb[l1]]],
                   read_bytes=b["read_bytes"][b["read_off"][lrb[l0]]:b["read_off"][lrb[l1]]])
        local, _ = po.viterbi_batch(sub)
        full = shard.gather_in_locus_order(local)
        if rank == 0:
            want, _ = po.viterbi_batch(b)
            q.put(bool(np.array_equal(full, want)))
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


def test_two_ranks_gather_in_locus_order():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
