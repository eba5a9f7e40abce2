"""GPU: ltr_genotyper_run (many raw loci -> calls) against (a) the per-locus entry points of the same library fed by a
plain-Python pooling of the reads, and (b) what the reference's SeqStutterGenotyper::genotype did on the same loci
(tests/golden/pruning.json, recorded by oracle/_ref/ltr_ref_trace: alleles and flanks as its HaplotypeGenerator built
them, final allele set, optimal pairs and posteriors)."""
import numpy as np
import pytest

import golden_util as gu
import synth
from longtr_b200 import Genotyper, build_locus_batch
from longtr_b200.flat import make_flat_locus

pytestmark = pytest.mark.gpu

ONT = (-1.0, -0.458675, -1.0, -0.458675, -0.0202027, -4.60517, -4.60517)


@pytest.fixture(scope="module")
def genotyper():
    g = Genotyper(devices=(0,), host_threads=4, chunk_loci=7)   # small chunks: several jobs in flight even in tests
    yield g
    g.close()


def _sampled_locus(seed, n_samples):
    rng = np.random.default_rng(seed)
    loc = synth.make_locus(9000 + seed, n_reads=int(rng.integers(6, 40)), n_decoys=int(rng.integers(0, 4)),
                           indel=float(rng.choice([1e-3, 0.02])))
    reads = []
    order = sorted(range(len(loc["reads"])), key=lambda i: int(rng.integers(0, n_samples)))  # any read set, sample-major
    sample_of = np.sort(rng.integers(0, n_samples, size=len(order)))
    for k, i in enumerate(order):
        r = loc["reads"][i]
        hp = int(rng.integers(0, 3))
        reads.append(dict(start=r["start"], stop=r["stop"], seq=r["seq"], cigar=r["cigar"], sample=int(sample_of[k]),
                          log_p1=[-1e-6, -1000.0, -0.6931471805599453][hp], log_p2=[-1000.0, -1e-6, -0.6931471805599453][hp]))
    return dict(lflank=loc["lflank"], rflank=loc["rflank"], alleles=loc["alleles"], repeat_start=loc["repeat_start"],
                repeat_end=loc["repeat_end"], n_samples=n_samples, reads=reads, period=loc["period"], motif=loc["motif"])


def _expected_by_per_locus_calls(engine, L):
    """Pool in Python, then HapAligner::process_reads on the pools (ltr_process_reads_flat) and the removal of uncalled
    alleles + call extraction (ltr_genotype_locus_pruned) -- the per-locus route through the same library."""
    pools, pool_of = [], []
    for r in L["reads"]:
        for k, q in enumerate(pools):
            if q["seq"] == r["seq"]:
                pool_of.append(k)
                break
        else:
            pool_of.append(len(pools))
            pools.append(r)
    flat, keep = make_flat_locus(L["lflank"], L["alleles"], L["rflank"], L["repeat_start"], L["repeat_end"], L["period"],
                                 [(q["start"], q["stop"], q["seq"], "I" * len(q["seq"]), q["cigar"]) for q in pools],
                                 motif=L["motif"])
    ll, seeds = engine.process_reads_flat(flat, len(pools), len(L["alleles"]))
    rows = ll[pool_of]
    rps = np.bincount([r["sample"] for r in L["reads"]], minlength=L["n_samples"]).astype(np.int32)
    p1 = np.array([r["log_p1"] for r in L["reads"]])
    p2 = np.array([r["log_p2"] for r in L["reads"]])
    got = engine.genotype_locus_pruned(rows, p1, p2, rps, seeds=seeds[pool_of])
    return len(pools), got


def test_batch_calls_match_the_per_locus_route(engine, genotyper):
    loci = [_sampled_locus(s, 1 + s % 3) for s in range(40)]
    out = genotyper.run(build_locus_batch(loci))
    assert (out["status"] == 0).all()
    for l, L in enumerate(loci):
        n_pools, want = _expected_by_per_locus_calls(engine, L)
        assert out["n_pools"][l] == n_pools
        a0, a1 = out["locus_allele_begin"][l], out["locus_allele_begin"][l + 1]
        kept = list(np.nonzero(out["kept_mask"][a0:a1])[0])
        assert kept == list(want["kept"]), l
        s0, s1 = out["locus_sample_begin"][l], out["locus_sample_begin"][l + 1]
        S, K = s1 - s0, len(kept)
        assert [int(kept[g]) for g in want["best_gts"].ravel()] == list(out["gts"][s0:s1].ravel())
        np.testing.assert_allclose(out["log_phased_posteriors"][s0:s1], want["log_phased_posteriors"], rtol=1e-12, atol=1e-10)
        np.testing.assert_allclose(out["log_unphased_posteriors"][s0:s1], want["log_unphased_posteriors"], rtol=1e-12, atol=1e-10)
        np.testing.assert_allclose(out["sample_total_lls"][s0:s1], want["sample_total_lls"], rtol=1e-12, atol=1e-10)
        np.testing.assert_allclose(out["gl_diffs"][s0:s1], want["gl_diffs"], rtol=1e-9, atol=1e-6)
        n_gl = K * (K + 1) // 2   # ltr_genotype_locus_pruned packs its arrays for the K surviving alleles
        want_gls, want_pls = want["gls"].ravel(), want["pls"].ravel()
        for s in range(S):
            g0 = out["gl_begin"][s0 + s]
            np.testing.assert_allclose(out["gls"][g0:g0 + n_gl], want_gls[s * n_gl:(s + 1) * n_gl], rtol=1e-9, atol=1e-6)
            assert np.max(np.abs(out["pls"][g0:g0 + n_gl] - want_pls[s * n_gl:(s + 1) * n_gl])) <= 1
        assert list(out["n_reads"][s0:s1]) == list(np.bincount([r["sample"] for r in L["reads"]], minlength=S))


def test_a_malformed_locus_fails_alone(genotyper):
    loci = [_sampled_locus(100 + s, 2) for s in range(12)]
    good = genotyper.run(build_locus_batch(loci))
    bad = [dict(l) for l in loci]
    bad[5] = dict(bad[5], reads=[dict(r) for r in bad[5]["reads"]])
    bad[5]["reads"][0]["cigar"] = "7Q" + bad[5]["reads"][0]["cigar"]        # an operation LongTR dies on
    bad[9] = dict(bad[9], reads=[dict(r) for r in bad[9]["reads"]])
    bad[9]["reads"][1]["log_p1"] = 0.5                                      # the reference asserts log_p <= 0
    out = genotyper.run(build_locus_batch(bad))
    assert out["status"][5] == -3 and out["status"][9] == -3
    ok = [l for l in range(12) if l not in (5, 9)]
    assert (out["status"][ok] == 0).all()
    for l in ok:
        s0, s1 = out["locus_sample_begin"][l], out["locus_sample_begin"][l + 1]
        assert np.array_equal(out["gts"][s0:s1], good["gts"][s0:s1])
        assert np.array_equal(out["log_unphased_posteriors"][s0:s1], good["log_unphased_posteriors"][s0:s1])


def test_batch_calls_match_the_reference_genotyper(genotyper):
    """Loci as the reference saw them (reads with their CIGARs, sample by sample; flanks and candidate alleles as its
    HaplotypeGenerator built them): same surviving alleles, same optimal pairs, same posteriors as
    SeqStutterGenotyper::genotype (src/seq_stutter_genotyper.cpp:599-645)."""
    real = {c["name"]: c for c in gu.load_real_cases()}
    cases = gu.load("pruning")
    groups = {}
    for c in cases:
        groups.setdefault(tuple(c["aln_params"]) if c["aln_params"] else None, []).append(c)
    n_checked = 0
    for params, cs in groups.items():
        loci = []
        for c in cs:
            reads = c["reads"] if c["reads"] is not None else [
                dict(start=r["start"], stop=r["stop"], seq=r["seq"], cigar=r["cigar"], sample=r["sample"], log_p1=r["log_p1"],
                     log_p2=r["log_p2"]) for r in real[c["real_case"]]["reads"]]
            loci.append(dict(lflank=c["lflank"], rflank=c["rflank"], alleles=c["alleles"], repeat_start=c["repeat_start"],
                             repeat_end=c["repeat_end"], n_samples=c["S"], haploid=c["haploid"], reads=reads))
        out = genotyper.run(build_locus_batch(loci), aln_params=params)
        assert (out["status"] == 0).all()
        for l, c in enumerate(cs):
            a0, a1 = out["locus_allele_begin"][l], out["locus_allele_begin"][l + 1]
            kept = list(np.nonzero(out["kept_mask"][a0:a1])[0])
            assert kept == c["kept"], c["name"]
            s0, s1 = out["locus_sample_begin"][l], out["locus_sample_begin"][l + 1]
            S, K = c["S"], len(kept)
            want_gts = [kept[g] for g in c["out_gts"]]
            assert list(out["gts"][s0:s1].ravel()) == want_gts, c["name"]
            post = gu.unhex(c["out_post"], (S, K, K))
            want_lpp = [post[s, c["out_gts"][2 * s], c["out_gts"][2 * s + 1]] for s in range(S)]
            np.testing.assert_allclose(out["log_phased_posteriors"][s0:s1], want_lpp, rtol=1e-10, atol=1e-9, err_msg=c["name"])
            np.testing.assert_allclose(out["sample_total_lls"][s0:s1], gu.unhex(c["out_totals"]), rtol=1e-10, atol=1e-9)
            n_checked += 1
    assert n_checked == len(cases)


def _n_devices():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_devices() < 2, reason="needs two GPUs (one process driving several devices)")
def test_one_process_two_devices_same_calls_in_input_order(genotyper):
    """ltr_genotyper_create(devices = {0, 1}): chunks of the batch go round robin over the devices and every result is written
    at its input index -- identical to what one device returns."""
    loci = [_sampled_locus(1000 + s, 1 + s % 3) for s in range(96)]
    batch = build_locus_batch(loci)
    want = genotyper.run(batch)
    two = Genotyper(devices=(0, 1), host_threads=8, chunk_loci=8)   # 12 chunks, 6 per device
    try:
        got = two.run(batch)
    finally:
        two.close()
    for key in ("status", "kept_mask", "n_kept", "n_pools", "gts", "n_reads", "pls", "locus_sample_begin", "locus_allele_begin"):
        np.testing.assert_array_equal(got[key], want[key], err_msg=key)
    for key in ("log_phased_posteriors", "log_unphased_posteriors", "gl_diffs", "sample_total_lls", "gls"):
        np.testing.assert_array_equal(got[key], want[key], err_msg=key)   # same kernels on the same inputs: same doubles
