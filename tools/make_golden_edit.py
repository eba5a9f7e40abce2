"""Records tests/golden/edit.json from the REFERENCE's own HaplotypeGenerator::needleman_wunsch / greedy_clustering
(oracle/_ref/libltr_ref.so, compiled in place from /root/reference by oracle/build_ref.sh):

    python tools/make_golden_edit.py

Inputs come from tests/edit_cases.py (seeded); the file holds inputs and the reference's answers."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import edit_cases as ec  # noqa: E402
from longtr_b200 import abi  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def main():
    assert po.ref_available()
    pairs = ec.pair_cases(seed=101, n_random=330)
    out = {"source": "HaplotypeGenerator.cpp:201-271 via oracle/edit_driver.cpp", "pairs": [], "pairs_at_threshold": [],
           "clusters": []}
    for a, b, T in pairs:
        if len(a) * len(b) <= 1400 * 1400:
            out["pairs"].append([a, b, T, po.edit_score(a, b, T, "ref")])
    for a, b, T in ec.at_threshold_cases(pairs, lambda x, y: po.edit_score(x, y, 999, "ref")):
        if len(a) + len(b) < 900:
            out["pairs_at_threshold"].append([a, b, T, po.edit_score(a, b, T, "ref")])
    for seqs, T in ec.cluster_cases(seed=77, n_sets=30):
        if sum(len(s) for s in seqs) > 40000:
            continue
        data, off = abi.pack_seqs(seqs)
        ok, cent, n = po.greedy_cluster(data, off, np.arange(len(seqs), dtype=np.uint32), T, "ref")
        out["clusters"].append({"seqs": seqs, "T": T, "ok": ok, "centroid_of": cent.tolist(), "n_centroids": n})
    path = os.path.join(ROOT, "tests", "golden", "edit.json")
    json.dump(out, open(path, "w"))
    print(path, len(out["pairs"]), len(out["pairs_at_threshold"]), len(out["clusters"]), os.path.getsize(path))


if __name__ == "__main__":
    main()
