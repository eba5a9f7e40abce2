"""GPU parity through the reference-facing entry points: ltr_process_reads_flat (HapAligner::process_reads)
and ltr_genotype_locus (calc_log_sample_posteriors + extract_genotypes_and_likelihoods) against the values
recorded from the unmodified reference (tests/golden)."""
import numpy as np
import pytest

import golden_util as gu
import synth
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

LONG_CASES = [c for c in gu.load("appendix_a") + gu.load("process_reads_long")
              if not (c["switch"] != 0 and c["period"] == 1)]


@pytest.mark.parametrize("case", LONG_CASES, ids=lambda c: c["name"])
def test_process_reads_flat_matches_reference(engine, case):
    L, keep = gu.flat_locus(case)
    P, H = len(case["reads"]), len(case["alleles"])
    ll, seeds = engine.process_reads_flat(L, P, H, fill=case.get("fill", 0.0))
    assert np.array_equal(ll, gu.unhex(case["ll"], (P, H)))  # bit-exact, untouched slots included
    for r in range(P):
        if case.get("realign_read") is None or case["realign_read"][r]:
            assert seeds[r] == case["seeds"][r]


@pytest.mark.parametrize("seed", range(12))
def test_process_reads_flat_matches_oracle_on_fresh_loci(engine, seed):
    loc = synth.make_locus(7000 + seed, n_reads=10, sub=0.01, indel=0.02, ref_len=40 + 25 * seed)
    L, keep = synth.to_flat(loc)
    P, H = len(loc["reads"]), len(loc["alleles"])
    want, wseeds, _ = po.process_reads(L, P, H)
    got, gseeds = engine.process_reads_flat(L, P, H)
    assert np.array_equal(got, want)
    assert np.array_equal(gseeds, wseeds)


def test_process_reads_flat_batch_matches_single_calls(engine):
    """ltr_process_reads_flat_batch: golden long-path cases (several parameter sets, untouched slots), a homopolymer
    locus on the stutter path and fresh loci in ONE call == the reference's values / the per-locus entry point."""
    cases = list(LONG_CASES) + [c for c in gu.load("process_reads_short")][:2]
    loci, keeps, shapes, fills = [], [], [], []
    for c in cases:
        L, keep = gu.flat_locus(c)
        loci.append(L)
        keeps.append(keep)
        shapes.append((len(c["reads"]), len(c["alleles"])))
        fills.append(c.get("fill", 0.0))
    fresh = []
    for seed in range(40):
        loc = synth.make_locus(9100 + seed, n_reads=12, sub=0.01, indel=0.02, ref_len=40 + 11 * seed)
        L, keep = synth.to_flat(loc)
        loci.append(L)
        keeps.append(keep)
        shapes.append((len(loc["reads"]), len(loc["alleles"])))
        fills.append(0.0)
        fresh.append(loc)
    lls, seeds = engine.process_reads_flat_batch(loci, shapes, fill=fills)
    for i, c in enumerate(cases):
        P, H = shapes[i]
        assert np.array_equal(lls[i], gu.unhex(c["ll"], (P, H))), c["name"]
        for r in range(P):
            if c.get("realign_read") is None or c["realign_read"][r]:
                assert seeds[i][r] == c["seeds"][r]
    for k in range(len(fresh)):
        i = len(cases) + k
        want, wseeds = engine.process_reads_flat(loci[i], *shapes[i])
        assert np.array_equal(lls[i], want) and np.array_equal(seeds[i], wseeds)
    empty_ll, empty_seeds = engine.process_reads_flat_batch([], [])
    assert empty_ll == [] and empty_seeds == []
    # one malformed locus: the call fails as a whole, loudly (the pipeline then retries the loci one by one)
    from longtr_b200 import LongTRError
    bad, keep_bad = synth.to_flat(fresh[0])
    bad.period = 0
    with pytest.raises(LongTRError):
        engine.process_reads_flat_batch([loci[-1], bad, loci[-2]], [shapes[-1], shapes[-1], shapes[-2]])


# posteriors: CUDA exp/log vs glibc (<= 1 ulp each) over sums of <= a few hundred terms
RTOL, ATOL = 1e-12, 1e-10


@pytest.mark.parametrize("case", gu.load("calls"), ids=lambda c: c["name"])
def test_genotype_locus_matches_reference(engine, case):
    S, H = case["S"], case["H"]
    R = sum(case["reads_per_sample"])
    got = engine.genotype_locus(gu.unhex(case["ll"], (R, H)), gu.unhex(case["log_p1"]), gu.unhex(case["log_p2"]),
                                case["reads_per_sample"], haploid=case["haploid"])
    assert np.array_equal(got["ll_clamped"], gu.unhex(case["out_ll_clamped"], (R, H)))
    assert list(got["best_gts"].ravel()) == case["out_best_gts"]  # GT identical
    for k in ("log_sample_posteriors", "sample_total_lls", "log_phased_posteriors", "log_unphased_posteriors",
              "hap_log_phased_posteriors", "phased_gls"):
        np.testing.assert_allclose(got[k].ravel(), gu.unhex(case["out_" + k]), rtol=RTOL, atol=ATOL, err_msg=k)
    # GL / GLDIFF / unphased haplotype posterior go through the reference's approximate two-argument
    # fast_log_sum_exp (single-precision fastlog(1+fastexp(d)), mathops.cpp:87-96): a 1e-13 change of its
    # argument can move the result by one float ulp (~6e-8), so the floor is 1e-6 for these fields
    for k in ("gls", "gl_diffs", "hap_log_unphased_posteriors"):
        np.testing.assert_allclose(got[k].ravel(), gu.unhex(case["out_" + k]), rtol=1e-9, atol=1e-6, err_msg=k)
    # PL = (int)(-10*dGL): an integer boundary can flip on a 1e-12 difference
    assert np.max(np.abs(got["pls"].ravel() - np.array(case["out_pls"]))) <= 1
    assert abs(got["total_ll"] - float.fromhex(case["total_ll"])) <= ATOL + RTOL * abs(got["total_ll"])


def test_genotype_locus_pruned_drops_uncalled_alleles(engine):
    """SeqStutterGenotyper::genotype's second stage: uncalled non-reference alleles are removed and the posteriors
    recomputed on the kept LL columns (checked against the oracle's posteriors on the pruned matrix)."""
    rng = np.random.default_rng(77)
    for t in range(20):
        S, H = int(rng.integers(1, 4)), int(rng.integers(2, 8))
        rps = [int(x) for x in rng.integers(3, 12, size=S)]
        R = sum(rps)
        lab = np.repeat(np.arange(S), rps).astype(np.int32)
        true = rng.integers(0, H, size=(S, 2))
        ll = -rng.exponential(30, size=(R, H)) - 8
        hp = rng.integers(0, 2, size=R)
        for r in range(R):
            ll[r, true[lab[r], hp[r]]] = -rng.exponential(0.2)
        p1 = np.where(hp == 0, -1e-6, -1000.0)
        p2 = np.where(hp == 0, -1000.0, -1e-6)
        seeds = np.full(R, 10, np.int32)
        if t % 5 == 0 and S > 1:
            seeds[lab == S - 1] = -1                    # a sample without aligned reads does not vote
        got = engine.genotype_locus_pruned(ll, p1, p2, rps, seeds=seeds)
        # expected kept set from the oracle's first pass
        _cl, _post, _tot, _total, best = po.log_sample_posteriors(ll, p1, p2, lab, S)
        voters = [s for s in range(S) if np.any(seeds[lab == s] >= 0)]
        called = {0} | {int(a) for s in voters for a in best[s]}
        kept = sorted(called)
        assert list(got["kept"]) == kept
        sub = np.maximum(ll, -600.0)[:, kept]
        _cl2, wpost, wtot, _t2, wbest = po.log_sample_posteriors(sub, p1, p2, lab, S)
        K = len(kept)
        np.testing.assert_allclose(got["log_sample_posteriors"].ravel()[:S * K * K], wpost.ravel(), rtol=RTOL, atol=ATOL)
        assert list(got["best_gts"].ravel()) == list(wbest.ravel())


@pytest.mark.parametrize("case", gu.load("pruning"), ids=lambda c: c["name"])
def test_genotype_locus_pruned_matches_reference_genotype(engine, case):
    """ltr_genotype_locus_pruned against what the reference's SeqStutterGenotyper::genotype did on the same LL matrix
    (tests/golden/pruning.json, recorded by oracle/_ref/ltr_ref_trace from src/seq_stutter_genotyper.cpp:634-645): same
    surviving alleles, same second-pass posteriors (1e-12: CUDA exp/log vs libm), same optimal pairs."""
    S, H, R = case["S"], case["H"], case["R"]
    got = engine.genotype_locus_pruned(gu.unhex(case["ll"], (R, H)), gu.unhex(case["log_p1"]), gu.unhex(case["log_p2"]),
                                       case["reads_per_sample"], haploid=case["haploid"],
                                       seeds=np.array(case["seeds"], np.int32))
    assert list(got["kept"]) == case["kept"]
    K = len(case["kept"])
    np.testing.assert_allclose(got["log_sample_posteriors"].ravel()[:S * K * K], gu.unhex(case["out_post"]),
                               rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(got["sample_total_lls"].ravel()[:S], gu.unhex(case["out_totals"]), rtol=RTOL, atol=ATOL)
    assert list(got["best_gts"].ravel()) == case["out_gts"]
