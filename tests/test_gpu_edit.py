"""N2 on the GPU: ltr_edit_distances / ltr_cluster_greedy (edit_kernel.cu) through the C ABI against the oracle and the
golden file recorded from the reference's HaplotypeGenerator::needleman_wunsch / greedy_clustering.  Exact integers."""
import json
import os

import numpy as np
import pytest

import edit_cases as ec
from longtr_b200 import abi
from longtr_b200.engine import LongTRError
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "edit.json")


def run_pairs(engine, cases):
    seqs, index = [], {}
    for a, b, _ in cases:
        for s in (a, b):
            if s not in index:
                index[s] = len(seqs)
                seqs.append(s)
    data, off = abi.pack_seqs(seqs)
    pa = [index[a] for a, _, _ in cases]
    pb = [index[b] for _, b, _ in cases]
    return engine.edit_distances(data, off, pa, pb, [T for _, _, T in cases])


def test_pairs_golden(engine):
    g = json.load(open(GOLDEN))
    rows = g["pairs"] + g["pairs_at_threshold"]
    got, st = run_pairs(engine, [(a, b, T) for a, b, T, _ in rows])
    assert got.tolist() == [w for _, _, _, w in rows]
    assert st.n_fallback > 50  # pairs whose distance equals T went through the exact kernel


def test_pairs_vs_oracle(engine):
    cases = ec.pair_cases(seed=3, n_random=400)
    cases += ec.at_threshold_cases(cases, lambda a, b: po.edit_score(a, b, 999))
    got, _ = run_pairs(engine, cases)
    want = [po.edit_score(a, b, T) for a, b, T in cases]
    assert got.tolist() == want


def test_pairs_long_strings(engine):
    """Several strips of both kernels (strings of up to 5 kb), thresholds at the distance."""
    rng = np.random.default_rng(17)
    cases = []
    for k in range(12):
        a = ec.tr_allele(rng, ec.rand_seq(rng, 37), int(rng.integers(60, 130)), 4)
        b = ec.mutate(rng, a, 0.01, 0.01)
        d = po.edit_score(a, b, 999)
        cases += [(a, b, 700), (b, a, 700), (a, b, d), (b, a, d), (a, b, max(0, d - 1))]
    got, _ = run_pairs(engine, cases)
    assert got.tolist() == [po.edit_score(a, b, T) for a, b, T in cases]


def test_clusters_golden(engine):
    g = json.load(open(GOLDEN))
    seqs, begin, items, Ts = ec.pack_sets([(c["seqs"], c["T"]) for c in g["clusters"]])
    data, off = abi.pack_seqs(seqs)
    cent, ncent, ok, st = engine.cluster_greedy(data, off, begin, items, Ts)
    for k, c in enumerate(g["clusters"]):
        assert int(ok[k]) == c["ok"]
        if c["ok"]:
            assert cent[begin[k]:begin[k + 1]].tolist() == c["centroid_of"]
            assert int(ncent[k]) == c["n_centroids"]


def test_clusters_vs_oracle_all_thresholds(engine):
    """Every set at every threshold of HaplotypeGenerator.cpp:403 in ONE call: the sets share the strings."""
    base = ec.cluster_cases(seed=31, n_sets=20)
    seqs, begin, items, Ts = [], [0], [], []
    for s, _ in base:
        first = len(seqs)
        seqs += s
        for T in ec.THRESHOLDS:
            items += list(range(first, first + len(s)))
            begin.append(len(items))
            Ts.append(T)
    data, off = abi.pack_seqs(seqs)
    cent, ncent, ok, _ = engine.cluster_greedy(data, off, begin, items, Ts)
    k = 0
    for s, _ in base:
        sd, so = abi.pack_seqs(s)
        for T in ec.THRESHOLDS:
            ook, ocent, on = po.greedy_cluster(sd, so, np.arange(len(s), dtype=np.uint32), T)
            assert int(ok[k]) == ook, (k, T)
            if ook:
                assert cent[begin[k]:begin[k + 1]].tolist() == ocent.tolist()
                assert int(ncent[k]) == on
            k += 1


def test_clusters_many_sets(engine):
    """The bench workload (hundreds of sets x 11 thresholds, > 200 000 items: many blocks, many pairs per warp) against
    the oracle on a sample of its sets."""
    from longtr_b200.workloads import generate_cluster_sets
    data, off, sb = generate_cluster_sets(256)
    items, begin, Ts = [], [0], []
    for k in range(len(sb) - 1):
        ids = np.arange(sb[k], sb[k + 1], dtype=np.uint32)
        for T in ec.THRESHOLDS:
            items.append(ids)
            begin.append(begin[-1] + len(ids))
            Ts.append(T)
    cent, ncent, ok, _ = engine.cluster_greedy(data, off, begin, np.concatenate(items), Ts)
    nT = len(ec.THRESHOLDS)
    for k in range(0, len(sb) - 1, 16):
        ids = np.arange(sb[k], sb[k + 1], dtype=np.uint32)
        for ti in (0, 1, 2, 5):
            ook, ocent, on = po.greedy_cluster(data, off, ids, ec.THRESHOLDS[ti])
            q = k * nT + ti
            assert int(ok[q]) == ook
            if ook:
                assert cent[begin[q]:begin[q + 1]].tolist() == ocent.tolist() and int(ncent[q]) == on


def test_edit_errors(engine):
    data, off = abi.pack_seqs(["ACGT", "ACGA"])
    with pytest.raises(LongTRError):
        engine.edit_distances(data, off, [0], [2], [20])       # sequence index out of range
    with pytest.raises(LongTRError):
        engine.edit_distances(data, off, [0], [1], [1000])     # threshold beyond the reference's row-minimum start
    with pytest.raises(LongTRError):
        engine.cluster_greedy(data, off, [0, 2], [0, 5], [20])
    bad = off.copy()
    bad[1] = 9
    with pytest.raises(LongTRError):
        engine.edit_distances(data, bad, [0], [1], [20])
    out, _ = engine.edit_distances(data, off, [], [], [])
    assert len(out) == 0
    cent, ncent, ok, _ = engine.cluster_greedy(data, off, [0, 0, 2], [0, 1], [20, 20])
    assert ncent.tolist() == [0, 1] and ok.tolist() == [1, 1] and cent.tolist() == [0, 0]
