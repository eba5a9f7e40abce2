"""N4: the restatement of the stutter-model EM (oracle/pyem.py) against the reference's EMStutterGenotyper -- recorded
(tests/golden/em.json, tools/make_golden_em.py) and, where oracle/_ref is present, live on fresh loci.  Bit for bit: the
restatement runs on the same libm as the reference."""
import json
import os

import numpy as np
import pytest

import em_cases
from oracle import pyem
from oracle import pyoracle as po

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "em.json")


def _check(c, r):
    assert r["trained"] == c["trained"] and r["n_iter"] == c["n_iter"], c["seed"]
    assert [float(x).hex() for x in r["params"]] == c["params"], c["seed"]
    assert float(r["lls"][-1]).hex() == c["ll"], c["seed"]
    assert [float(x).hex() for x in r["log_gt_priors"]] == c["log_gt_priors"], c["seed"]


def test_restatement_reproduces_the_recorded_reference():
    g = json.load(open(GOLD))
    for c in g["cases"][:60]:
        L = em_cases.em_locus(c["seed"])
        _check(c, pyem.em_train(L["reads_per_sample"], L["bp_diff"], L["log_p1"], L["log_p2"], L["motif_len"], L["haploid"]))
    for c in g["short"]:
        L = em_cases.em_locus(c["seed"])
        _check(c, pyem.em_train(L["reads_per_sample"], L["bp_diff"], L["log_p1"], L["log_p2"], L["motif_len"], L["haploid"],
                                max_iter=c["max_iter"]))
    assert any(not c["trained"] for c in g["short"])


@pytest.mark.skipif(not po.ref_em_available(), reason="oracle/_ref/libltr_ref_em.so not built")
def test_restatement_matches_the_reference_on_fresh_loci():
    for seed in range(5000, 5030):
        L = em_cases.em_locus(seed)
        a = po.ref_em_train(L["reads_per_sample"], L["bp_diff"], L["log_p1"], L["log_p2"], L["motif_len"], L["haploid"])
        b = pyem.em_train(L["reads_per_sample"], L["bp_diff"], L["log_p1"], L["log_p2"], L["motif_len"], L["haploid"])
        assert a["trained"] == b["trained"] and a["n_iter"] == b["n_iter"]
        assert np.array_equal(a["params"], b["params"]) and np.array_equal(a["log_gt_priors"], b["log_gt_priors"])
        assert a["lls"][-1] == b["lls"][-1]


def test_single_precision_approximations_known_values():
    """fastexp / fastlog / fasterexp / fasterlog against values computed with the reference's header (fastonebigheader.h)."""
    assert abs(float(pyem.fastexp(-1.0)) - np.exp(-1.0)) < 1e-4 and abs(float(pyem.fastlog(2.0)) - np.log(2.0)) < 1e-4
    assert abs(float(pyem.fasterexp(-1.0)) - np.exp(-1.0)) < 0.03 and abs(float(pyem.fasterlog(2.0)) - np.log(2.0)) < 0.06
