"""CPU: the host-side (integer / small-vector) half of the drop-in boundary -- no GPU involved.
trim_alignment, calc_seed_base and extract_genotypes_and_likelihoods of the C++ host mirror
(longtr_b200/csrc/host) against values recorded from the reference (tests/golden) and the oracle."""
import numpy as np
import pytest

import golden_util as gu
import synth
from longtr_b200 import abi
from oracle import pyoracle as po


@pytest.mark.parametrize("case", gu.load("appendix_a") + gu.load("process_reads_long"), ids=lambda c: c["name"])
def test_seed_base_matches_reference(case):
    L, keep = gu.flat_locus(case)
    got = [abi.seed_base(L, r) for r in range(len(case["reads"]))]
    assert got == case["seed_bases"]


@pytest.mark.parametrize("case", gu.load("appendix_a") + gu.load("process_reads_long"), ids=lambda c: c["name"])
def test_trim_matches_oracle(case):
    L, keep = gu.flat_locus(case)
    for r in range(len(case["reads"])):
        want = po.trim_read(L, r)
        got = abi.trim_read(L, r)
        if len(got) == 0:  # the oracle helper already substitutes the 10 bp pseudo read
            got = case["lflank"][-5:].encode() + case["rflank"][:5].encode()
        assert got == want


def test_trim_window_appendix_a():
    """trimmed read = bases aligned to [block.start-5, block.end+5)."""
    a1 = [c for c in gu.load("appendix_a") if c["name"] == "A1"][0]
    L, keep = gu.flat_locus(a1)
    assert abi.trim_read(L, 0).decode() == a1["lflank"][30:] + "CTGAA" + "AC" * 12 + "GGTCT" + a1["rflank"][:5]


@pytest.mark.parametrize("seed", range(30))
def test_trim_with_indels_near_the_pads(seed):
    loc = synth.make_locus(9000 + seed, n_reads=6, sub=0.02, indel=0.08, ref_len=30)
    L, keep = synth.to_flat(loc)
    for r in range(6):
        want = po.trim_read(L, r)
        got = abi.trim_read(L, r)
        if len(got) == 0:
            got = loc["lflank"][-5:].encode() + loc["rflank"][:5].encode()
        assert got == want


def test_bad_cigar_is_an_error_not_an_exit():
    loc = synth.make_locus(1, n_reads=1)
    loc["reads"][0]["cigar"] = "10Q"
    L, keep = synth.to_flat(loc)
    assert abi.load().ltr_seed_base_flat(L, 0) == -3
    with pytest.raises(RuntimeError):
        abi.trim_read(L, 0)


@pytest.mark.parametrize("case", gu.load("calls"), ids=lambda c: c["name"])
def test_extract_calls_matches_reference(case):
    S, H = case["S"], case["H"]
    post = gu.unhex(case["out_log_sample_posteriors"], (S, H, H))
    totals = gu.unhex(case["out_sample_total_lls"])
    got = abi.extract_calls(post, totals, haploid=case["haploid"])
    assert list(got["best_gts"].ravel()) == case["out_best_gts"]
    assert list(got["pls"].ravel()) == case["out_pls"]
    for k in ("log_phased_posteriors", "log_unphased_posteriors", "hap_log_phased_posteriors",
              "hap_log_unphased_posteriors", "gls", "phased_gls", "gl_diffs"):
        assert np.array_equal(got[k].ravel(), gu.unhex(case["out_" + k])), k


def test_flatten_loci_feeds_the_oracle_to_the_reference_values():
    """ltr_flatten_loci (host mirror: Haplotype column order, realign masks, trim_alignment, 10 bp pseudo reads) on the
    golden long-path loci: the oracle's Viterbi on the flattened batch, scattered by hap_col / read_row, reproduces the
    aln_probs matrices recorded from the unmodified reference bit for bit -- one locus at a time and all loci that share
    their parameters in one batch."""
    from longtr_b200 import abi
    from oracle import pyoracle as po
    cases = [c for c in gu.load("appendix_a") + gu.load("process_reads_long") if not (c["switch"] != 0 and c["period"] == 1)]
    assert len(cases) >= 30

    def check(group):
        loci, keeps = [], []
        for c in group:
            L, keep = gu.flat_locus(c)
            loci.append(L)
            keeps.append(keep)
        batch, hap_col, read_row, params, flank = abi.flatten_loci(loci)
        ll, _ = po.viterbi_batch(batch, aln_params=params, indel_flank_len=flank)
        pos = 0
        for i, c in enumerate(group):
            P, H = len(c["reads"]), len(c["alleles"])
            h0, h1 = int(batch["locus_hap_begin"][i]), int(batch["locus_hap_begin"][i + 1])
            r0, r1 = int(batch["locus_read_begin"][i]), int(batch["locus_read_begin"][i + 1])
            got = np.full((P, H), c.get("fill", 0.0))
            block = ll[pos:pos + (h1 - h0) * (r1 - r0)].reshape(r1 - r0, h1 - h0)
            pos += (h1 - h0) * (r1 - r0)
            for rr in range(r1 - r0):
                for hh in range(h1 - h0):
                    got[read_row[r0 + rr], hap_col[h0 + hh]] = block[rr, hh]
            assert np.array_equal(got, gu.unhex(c["ll"], (P, H))), c["name"]

    for c in cases:
        check([c])
    groups = {}
    for c in cases:
        groups.setdefault((tuple(c.get("aln_params") or ()), c.get("indel_flank_len", 5)), []).append(c)
    for g in groups.values():
        check(g)


@pytest.mark.parametrize("seed", range(10))
def test_flatten_loci_matches_the_oracle_flattening_on_fresh_loci(seed):
    """Same check against the oracle's own per-locus restatement (trim + haplotype order + Viterbi) on seeded loci with
    indels near the repeat boundaries, so that trim_alignment's un-trimming rules matter."""
    from longtr_b200 import abi
    from oracle import pyoracle as po
    loci, shapes, keeps = [], [], []
    for k in range(6):
        loc = synth.make_locus(20000 + 10 * seed + k, n_reads=10, sub=0.01, indel=0.03, ref_len=30 + 17 * k)
        L, keep = synth.to_flat(loc)
        loci.append(L)
        keeps.append(keep)
        shapes.append((len(loc["reads"]), len(loc["alleles"])))
    batch, hap_col, read_row, params, flank = abi.flatten_loci(loci)
    ll, _ = po.viterbi_batch(batch, aln_params=params, indel_flank_len=flank)
    pos = 0
    for i, (P, H) in enumerate(shapes):
        want, _seeds, _ = po.process_reads(loci[i], P, H)
        h0, h1 = int(batch["locus_hap_begin"][i]), int(batch["locus_hap_begin"][i + 1])
        r0, r1 = int(batch["locus_read_begin"][i]), int(batch["locus_read_begin"][i + 1])
        assert (h1 - h0, r1 - r0) == (H, P)
        got = np.zeros((P, H))
        block = ll[pos:pos + H * P].reshape(P, H)
        pos += H * P
        for rr in range(P):
            for hh in range(H):
                got[read_row[r0 + rr], hap_col[h0 + hh]] = block[rr, hh]
        assert np.array_equal(got, want)


@pytest.mark.parametrize("case", gu.load("pooling"), ids=lambda c: c["name"])
def test_read_pooler_matches_reference(case):
    """ltr::ReadPooler + BaseQuality::median_base_qualities (csrc/host/host_types.cpp) against the reference's ReadPooler
    (src/read_pooler.cpp:3-20, src/read_pooler.h:42-48, src/base_quality.cpp:11-28; fixture tests/golden/pooling.json
    recorded through oracle/_ref): same pool of every read, same per-position median qualities."""
    from longtr_b200 import abi
    pool, meds = abi.pool_reads(case["seqs"], case["quals"])
    assert pool == case["pool_index"]
    assert meds == case["median_quals"]
