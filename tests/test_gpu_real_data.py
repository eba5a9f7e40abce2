"""BASELINE.json configs[0] / configs[1] through the library's OWN pipeline (not through LongTR's classes as the drop-in test
does): the reads of the shipped HG002 / trio BAMs on the shipped regions (tests/golden/real_cases.json.gz, decoded by
tools/real_cases.py) -> ltr_candidate_alleles (exact candidates + assembly branch) -> ltr_genotyper_run (pooling, trimming,
Viterbi, posteriors, removal of uncalled alleles, call extraction) against the VCF record the all-CPU reference wrote for the
same reads: per sample the genotype as base-pair differences (GB), Q, PQ, DP and GLDIFF."""
import numpy as np
import pytest

import golden_util as gu
from longtr_b200 import Genotyper, abi, build_locus_batch

pytestmark = pytest.mark.gpu


def _fields(record):
    f = record.split("\t")
    fmt = f[8].split(":")
    return f, [dict(zip(fmt, s.split(":"))) if s != "." else None for s in f[9:]]


def _extras(out, s0, s1):
    """GL / PL / PHASEDGL slices of samples [s0, s1) of a Genotyper.run result, re-based for ltr_vcf_record_ex."""
    glb, pglb = out["gl_begin"][s0:s1 + 1], out["pgl_begin"][s0:s1 + 1]
    return dict(gl_begin=glb - glb[0], gls=out["gls"][glb[0]:glb[-1]], pls=out["pls"][glb[0]:glb[-1]],
                pgl_begin=pglb - pglb[0], phased_gls=out["phased_gls"][pglb[0]:pglb[-1]])


def test_real_loci_through_the_library_pipeline():
    cases = gu.load_real_cases()
    loci, keep = [], []
    for c in cases:
        cand = abi.candidate_alleles_from_reads(c["reads"], len(c["samples"]), c["region_start"], c["region_stop"],
                                                len(c["motif"]), c["chrom_seq"])
        assert cand["status"] == 0, c["name"]
        loci.append(dict(lflank=cand["lflank"], rflank=cand["rflank"], alleles=cand["alleles"], repeat_start=cand["block_start"],
                         repeat_end=cand["block_end"], n_samples=len(c["samples"]),
                         reads=[dict(start=r["start"], stop=r["stop"], seq=r["seq"], cigar=r["cigar"], sample=r["sample"],
                                     log_p1=r["log_p1"], log_p2=r["log_p2"]) for r in c["reads"]]))
        keep.append(cand)
    g = Genotyper(devices=(0,), host_threads=8, chunk_loci=16)
    try:
        g.set_read_alleles(True)
        g.set_phased_gls(True)
        out = g.run(build_locus_batch(loci))
    finally:
        g.close()
    assert (out["status"] == 0).all()
    # ---- the whole record as text: ltr_vcf_record on the device's numbers = the reference's record, character for character
    import test_vcf_writer as tw
    rb = np.concatenate([[0], np.cumsum([len(c["reads"]) for c in cases])])
    switch_gold = tw.load_switch_records()
    n_text = n_switched = 0
    for l, (c, cand) in enumerate(zip(cases, keep)):
        s0, s1 = out["locus_sample_begin"][l], out["locus_sample_begin"][l + 1]
        a0, a1 = out["locus_allele_begin"][l], out["locus_allele_begin"][l + 1]
        calls = dict(gts=out["gts"][s0:s1], lup=out["log_unphased_posteriors"][s0:s1], lpp=out["log_phased_posteriors"][s0:s1],
                     gld=out["gl_diffs"][s0:s1], kept=out["kept_mask"][a0:a1], read_allele=out["read_allele"][rb[l]:rb[l + 1]])
        got = abi.vcf_record(**tw.record_inputs(c, cand, calls))
        assert got == c["record"], c["name"]
        n_text += 1
        # ... and under the reference's output switches (GL / PL / PHASEDGL computed on the device's posteriors)
        extras = _extras(out, s0, s1)
        for mask, recs in switch_gold.items():
            got = abi.vcf_record(switches=mask, **tw.record_inputs(dict(c, record=recs[c["name"]]), cand, calls), **extras)
            assert got == recs[c["name"]], (c["name"], mask)
            n_switched += 1
    assert n_text == len(cases) and n_switched == 4 * len(cases)
    n_samples = n_het = n_inexact = 0
    for l, (c, cand) in enumerate(zip(cases, keep)):
        f, samples = _fields(c["record"])
        info = dict(kv.split("=") for kv in f[7].split(";"))
        s0 = out["locus_sample_begin"][l]
        lens = [len(a) for a in cand["alleles"]]
        n_inexact += sum(cand["inexact"])
        # INFO/INEXACT_ALLELE lists the alternate alleles that survive, in order of length
        a0, a1 = out["locus_allele_begin"][l], out["locus_allele_begin"][l + 1]
        kept = [k for k in range(a1 - a0) if out["kept_mask"][a0 + k]]
        want_inexact = [] if info["INEXACT_ALLELE"] == "." else [int(x) for x in info["INEXACT_ALLELE"].split(",")]
        alts = sorted((k for k in kept if k != 0), key=lambda k: (lens[k], cand["alleles"][k]))
        assert [cand["inexact"][k] for k in alts] == want_inexact, c["name"]
        if "BPDIFFS" in info:
            assert [lens[k] - lens[0] for k in alts] == [int(x) for x in info["BPDIFFS"].split(",")], c["name"]
        for s, want in enumerate(samples):
            if want is None:
                continue
            ga, gb = out["gts"][s0 + s]
            assert "%d|%d" % (lens[ga] - lens[0], lens[gb] - lens[0]) == want["GB"], (c["name"], s)
            assert abs(np.exp(out["log_unphased_posteriors"][s0 + s]) - float(want["Q"])) <= 0.0051, (c["name"], s)
            assert abs(np.exp(out["log_phased_posteriors"][s0 + s]) - float(want["PQ"])) <= 0.0051, (c["name"], s)
            assert int(out["n_reads"][s0 + s]) == int(want["DP"]), (c["name"], s)
            if want["GLDIFF"] != ".":
                assert abs(out["gl_diffs"][s0 + s] - float(want["GLDIFF"])) <= 0.0051 + 1e-4 * abs(float(want["GLDIFF"])), (c["name"], s)
            n_samples += 1
            n_het += ga != gb
    assert len(cases) >= 50 and n_samples >= 100 and n_het >= 10 and n_inexact >= 2


def test_seeded_loci_records_character_for_character():
    """SURVEY Appendix A4 and the ten seeded drop-in loci (tests/dropin_cases.py; haploid loci and custom alignment parameters
    among them) through the same route: the library's record = the reference's (tests/golden/vcf_records.json)."""
    import dropin_cases as dc
    import test_vcf_writer as tw
    want = {c["name"]: c["record"] for c in gu.load("vcf_records")}
    cases = [dc.case_a4()] + dc.seeded_cases()
    g = Genotyper(devices=(0,), host_threads=4, chunk_loci=16)
    g.set_read_alleles(True)
    g.set_phased_gls(True)
    switch_gold = tw.load_switch_records()
    n = n_switched = 0
    try:
        for c in cases:
            if c["name"] not in want:
                continue
            cand = abi.candidate_alleles_from_reads(c["reads"], len(c["samples"]), c["region_start"], c["region_stop"],
                                                    len(c["motif"]), c["chrom_seq"])
            assert cand["status"] == 0, c["name"]
            locus = dict(lflank=cand["lflank"], rflank=cand["rflank"], alleles=cand["alleles"], repeat_start=cand["block_start"],
                         repeat_end=cand["block_end"], n_samples=len(c["samples"]), haploid=bool(c.get("haploid")),
                         reads=[dict(start=r["start"], stop=r["stop"], seq=r["seq"], cigar=r["cigar"], sample=r["sample"],
                                     log_p1=r["log_p1"], log_p2=r["log_p2"]) for r in c["reads"]])
            out = g.run(build_locus_batch([locus]), aln_params=c.get("aln_params"))
            assert out["status"][0] == 0, c["name"]
            calls = dict(gts=out["gts"], lup=out["log_unphased_posteriors"], lpp=out["log_phased_posteriors"],
                         gld=out["gl_diffs"], kept=out["kept_mask"], read_allele=out["read_allele"])
            inp = tw.record_inputs(dict(c, record=want[c["name"]]), cand, calls)
            got = abi.vcf_record(haploid=bool(c.get("haploid")), **inp)
            assert got == want[c["name"]], c["name"]
            n += 1
            for mask, recs in switch_gold.items():   # the reference's output switches, haploid loci among them
                inp = tw.record_inputs(dict(c, record=recs[c["name"]]), cand, calls)
                got = abi.vcf_record(haploid=bool(c.get("haploid")), switches=mask, **inp, **_extras(out, 0, len(c["samples"])))
                assert got == recs[c["name"]], (c["name"], mask)
                n_switched += 1
        # the same loci as a haploid chromosome (--haploid-chrs): GT / GL / PL over single alleles, no PHASEDGL
        hap_gold = tw.load_switch_records("haploid")
        n_hap = 0
        for c in dc.haploid_cases():
            cand = abi.candidate_alleles_from_reads(c["reads"], len(c["samples"]), c["region_start"], c["region_stop"],
                                                    len(c["motif"]), c["chrom_seq"])
            locus = dict(lflank=cand["lflank"], rflank=cand["rflank"], alleles=cand["alleles"], repeat_start=cand["block_start"],
                         repeat_end=cand["block_end"], n_samples=len(c["samples"]), haploid=True,
                         reads=[dict(start=r["start"], stop=r["stop"], seq=r["seq"], cigar=r["cigar"], sample=r["sample"],
                                     log_p1=r["log_p1"], log_p2=r["log_p2"]) for r in c["reads"]])
            out = g.run(build_locus_batch([locus]), aln_params=c.get("aln_params"))
            assert out["status"][0] == 0, c["name"]
            calls = dict(gts=out["gts"], lup=out["log_unphased_posteriors"], lpp=out["log_phased_posteriors"],
                         gld=out["gl_diffs"], kept=out["kept_mask"], read_allele=out["read_allele"])
            for mask, recs in hap_gold.items():
                inp = tw.record_inputs(dict(c, record=recs[c["name"]]), cand, calls)
                got = abi.vcf_record(haploid=True, switches=mask, **inp, **_extras(out, 0, len(c["samples"])))
                assert got == recs[c["name"]], (c["name"], mask)
                n_hap += 1
    finally:
        g.close()
    assert n >= 10 and n_switched == 4 * n and n_hap == 25
