"""CPU: the oracle restatement (oracle/longtr_oracle.c) against the golden vectors recorded from
the unmodified reference sources (tests/golden, tools/make_golden.py) and SURVEY Appendix A."""
import numpy as np
import pytest

import golden_util as gu
from oracle import pyoracle as po

# SURVEY.md Appendix A (captured from the reference during the survey, %.17g)
APPENDIX_A = {
    "A1": ([-13.913093791954452, -17.912461765226908, -0.0068942923244321719], 203),
    "A2": ([-10.911829738499364, -11.911671731817478, -0.0051562188236857764], 192),
    "A3": ([-8.0872831978190813, -11.1676949945636, -4.5732380299363298], 125),
}


def test_golden_file_agrees_with_survey_appendix_a():
    for c in gu.load("appendix_a"):
        want, seed = APPENDIX_A[c["name"]]
        assert list(gu.unhex(c["ll"])) == want
        assert c["seeds"] == [seed]


LONG_CASES = [c for c in gu.load("appendix_a") + gu.load("process_reads_long")
              if not (c["switch"] != 0 and c["period"] == 1)]  # short path: tests/test_stutter_*.py


@pytest.mark.parametrize("case", LONG_CASES, ids=lambda c: c["name"])
def test_process_reads_matches_reference(case):
    L, keep = gu.flat_locus(case)
    P, H = len(case["reads"]), len(case["alleles"])
    want = gu.unhex(case["ll"], (P, H))
    ll, seeds, _ = po.process_reads(L, P, H, fill=case.get("fill", 0.0))
    assert np.array_equal(ll, want), (ll, want)
    assert list(seeds) == case["seeds"] or case.get("realign_read") is not None
    if case.get("realign_read") is not None:
        for r, on in enumerate(case["realign_read"]):
            if on:
                assert seeds[r] == case["seeds"][r]


SHORT_CASES = [c for c in gu.load("appendix_a") if c["switch"] != 0] + gu.load("process_reads_short")


@pytest.mark.parametrize("case", SHORT_CASES, ids=lambda c: c["name"])
def test_process_reads_short_path_matches_reference(case):
    """Homopolymer / --stutter-align-len path: same doubles, same float bit tricks -> bit equality."""
    L, keep = gu.flat_locus(case)
    P, H = len(case["reads"]), len(case["alleles"])
    ll, seeds, _ = po.process_reads(L, P, H, fill=case.get("fill", 0.0))
    assert np.array_equal(ll, gu.unhex(case["ll"], (P, H)))
    for r in range(P):
        if case.get("realign_read") is None or case["realign_read"][r]:
            assert seeds[r] == case["seeds"][r]


@pytest.mark.parametrize("case", gu.load("pair_batches"), ids=lambda c: c["name"])
def test_pair_batches_match_reference(case):
    b = gu.pair_batch(case)
    ll, _cells = po.viterbi_batch(b, aln_params=case["aln_params"], n_threads=2)
    assert np.array_equal(ll, gu.unhex(case["ll"]))


@pytest.mark.parametrize("case", gu.load("posteriors"), ids=lambda c: c["name"])
def test_posteriors_match_reference(case):
    S, H = case["S"], case["H"]
    lab = np.array(case["label"], dtype=np.int32)
    R = len(lab)
    cl, post, tot, total, best = po.log_sample_posteriors(gu.unhex(case["ll"], (R, H)), gu.unhex(case["log_p1"]),
                                                          gu.unhex(case["log_p2"]), lab, S, haploid=case["haploid"])
    # same libm, same order of operations -> bit equality
    assert np.array_equal(cl, gu.unhex(case["ll_clamped"], (R, H)))
    assert np.array_equal(post, gu.unhex(case["post"], (S, H, H)))
    assert np.array_equal(tot, gu.unhex(case["totals"]))
    assert total == float.fromhex(case["total"])
    assert list(best.ravel()) == case["best"]


@pytest.mark.parametrize("case", gu.load("pruning"), ids=lambda c: c["name"])
def test_both_posterior_passes_of_genotype_match_reference(case):
    """tests/golden/pruning.json: the two calls of Genotyper::calc_log_sample_posteriors inside the reference's
    SeqStutterGenotyper::genotype (before / after the uncalled alleles are removed), recorded by oracle/_ref/ltr_ref_trace.
    The restatement reproduces both passes bit for bit, and the surviving alleles are exactly the reference allele plus
    those in some voting sample's optimal pair (src/seq_stutter_genotyper.cpp:250-311)."""
    S, H, R = case["S"], case["H"], case["R"]
    lab = np.repeat(np.arange(S), case["reads_per_sample"]).astype(np.int32)
    p1, p2 = gu.unhex(case["log_p1"]), gu.unhex(case["log_p2"])
    ll = gu.unhex(case["ll"], (R, H))
    _cl, post, _tot, _total, best = po.log_sample_posteriors(ll, p1, p2, lab, S, haploid=case["haploid"])
    assert np.array_equal(post.ravel(), gu.unhex(case["first_post"]))
    assert list(best.ravel()) == case["first_gts"]
    seeds = np.array(case["seeds"])
    voters = [s for s in range(S) if np.any(seeds[lab == s] >= 0)]
    kept = sorted({0} | {int(a) for s in voters for a in best[s]})
    assert kept == case["kept"]
    K = len(kept)
    out_ll = gu.unhex(case["out_ll"], (R, K))
    if case["n_calls"] > 1:
        # the kept columns are carried over after the first pass clamped them (genotyper.cpp:57-58)
        assert np.array_equal(out_ll, np.maximum(ll, -600.0)[:, kept])
        _cl2, post2, tot2, _t2, best2 = po.log_sample_posteriors(out_ll, p1, p2, lab, S, haploid=case["haploid"])
        assert np.array_equal(post2.ravel(), gu.unhex(case["out_post"]))
        assert np.array_equal(tot2, gu.unhex(case["out_totals"]))
        assert list(best2.ravel()) == case["out_gts"]
