"""ltr_vcf_record (csrc/host/vcf_writer.cpp; reference SeqStutterGenotyper::write_vcf_record, get_alleles, reorder_alleles,
ExtractCigar, condense_read_counts) on the 52 shipped HG002 / trio loci of tests/golden/real_cases.json.gz and the seeded
drop-in loci.  Host only: candidate alleles come from ltr_candidate_alleles, the numbers a GPU would deliver (genotypes,
posteriors, GLDIFF) are read back from the reference's record, and everything else of the record -- position, trimmed / padded
alleles and their order, INFO counts, depth / phasing fields, ALLREADS, MALLREADS of homozygous samples -- must come out
character for character.  tests/test_gpu_real_data.py does the same with the numbers computed on the device (including the
per-read allele assignment behind MALLREADS of heterozygous samples)."""
import numpy as np
import pytest

import golden_util as gu
from longtr_b200 import abi


def record_inputs(c, cand, calls=None):
    """Everything ltr_vcf_record needs for real case c.  calls = None: read the genotypes / posteriors / GLDIFF back from the
    reference's record; otherwise dict(gts [S, 2] candidate indices, lup, lpp, gld, read_allele, kept)."""
    f = c["record"].split("\t")
    fmt = f[8].split(":")
    cols = [dict(zip(fmt, x.split(":"))) if x != "." else None for x in f[9:]]
    S = len(c["samples"])
    lens = [len(a) for a in cand["alleles"]]
    reads = c["reads"]
    if calls is None:
        gts, lup, lpp, gld, kept = [], [], [], [], [0] * len(lens)
        kept[0] = 1
        for s in range(S):
            w = cols[s]
            d = [int(x) for x in w["GB"].split("|")]
            idx = []
            for x in d:   # the allele of that size; ties are told apart by the record's allele order
                k = [k for k in range(len(lens)) if lens[k] - lens[0] == x]
                idx.append(k)
            gts.append(idx)
            lup.append(np.log(max(float(w["Q"]), 1e-300)))
            lpp.append(np.log(max(float(w["PQ"]), 1e-300)))
            gld.append(0.0 if w["GLDIFF"] == "." else float(w["GLDIFF"]))
        if any(len(k) != 1 for pair in gts for k in pair):
            return None   # two candidates of the same length: the sizes alone do not name the allele
        gts = [[pair[0][0], pair[1][0]] for pair in gts]
        alts = f[4].split(",") if f[4] != "." else []
        n_alt = len(alts)
        for pair in gts:
            for k in pair:
                kept[k] = 1
        if sum(kept) - 1 != n_alt:
            return None   # a surviving allele nobody carries in the reported genotype (kept by another sample's phase)
        read_allele = None
        calls = dict(gts=gts, lup=lup, lpp=lpp, gld=gld, kept=kept, read_allele=read_allele)
    bp = [abi.extract_cigar_bp_diff(r["cigar"], r["start"], c["region_start"] - 5, c["region_stop"] + 5) for r in reads]
    return dict(chrom=c["chrom_name"], name=c["region_name"], motif=c["motif"], region_start=c["region_start"],
                region_stop=c["region_stop"], chrom_seq=c["chrom_seq"], chrom_seq_start=0, block_start=cand["block_start"],
                block_end=cand["block_end"], alleles=cand["alleles"], inexact=cand["inexact"], kept_mask=calls["kept"],
                gts=np.array(calls["gts"]).ravel(), log_unphased=calls["lup"], log_phased=calls["lpp"], gl_diffs=calls["gld"],
                n_p1=c["n_p1s"], n_p2=c["n_p2s"], read_sample=[r["sample"] for r in reads],
                log_p1=[r["log_p1"] for r in reads], log_p2=[r["log_p2"] for r in reads], read_bp_diff=bp,
                read_allele=calls["read_allele"], column_sample=list(range(S)))


def strip_het_mallreads(record):
    f = record.split("\t")
    fmt = f[8].split(":")
    gi, mi = fmt.index("GT"), fmt.index("MALLREADS")
    for k in range(9, len(f)):
        if f[k] == ".":
            continue
        x = f[k].split(":")
        a, b = x[gi].split("|")
        if a != b:
            x[mi] = "?"
        f[k] = ":".join(x)
    return "\t".join(f)


def test_records_of_the_real_loci_character_for_character():
    cases = gu.load_real_cases()
    n, n_multi, n_pad = 0, 0, 0
    for c in cases:
        cand = abi.candidate_alleles_from_reads(c["reads"], len(c["samples"]), c["region_start"], c["region_stop"],
                                                len(c["motif"]), c["chrom_seq"])
        inp = record_inputs(c, cand)
        if inp is None:
            continue
        het = any(inp["gts"][2 * s] != inp["gts"][2 * s + 1] for s in range(len(c["samples"])))
        if het:   # the per-read assignment needs the LL matrix: give every read its sample's first allele, compare the rest
            inp["read_allele"] = [inp["gts"][2 * s] for s in inp["read_sample"]]
        got = abi.vcf_record(**inp)
        assert strip_het_mallreads(got) == strip_het_mallreads(c["record"]), c["name"]
        n += 1
        n_multi += got.split("\t")[4] != "."
        n_pad += int(got.split("\t")[1]) < c["region_start"] + 1
    assert n >= 45 and n_multi >= 15, (n, n_multi, n_pad)


def test_extract_cigar_bp_diff():
    assert abi.extract_cigar_bp_diff("200=9I245=", 800, 995, 1050) == 9
    assert abi.extract_cigar_bp_diff("200=6D239=", 800, 995, 1050) == -6
    assert abi.extract_cigar_bp_diff("445=", 800, 995, 1050) == 0
    assert abi.extract_cigar_bp_diff("445=", 800, 700, 1050) is None      # the region starts before the read
    assert abi.extract_cigar_bp_diff("445=", 800, 995, 1245) is None      # ... or ends behind it
    assert abi.extract_cigar_bp_diff("100=3I50=2D20=4I100=", 800, 890, 1000) == 5
    assert abi.extract_cigar_bp_diff("5I440=", 800, 800, 1000) is None    # nothing aligned in front of the region


def test_vcf_record_errors():
    c = gu.load_real_cases()[0]
    cand = abi.candidate_alleles_from_reads(c["reads"], len(c["samples"]), c["region_start"], c["region_stop"], len(c["motif"]),
                                            c["chrom_seq"])
    inp = record_inputs(c, cand)
    bad = dict(inp, gts=np.array([99, 99] * len(c["samples"])))
    with pytest.raises(RuntimeError):
        abi.vcf_record(**bad)
    bad = dict(inp, read_sample=[7] * len(inp["read_sample"]))
    with pytest.raises(RuntimeError):
        abi.vcf_record(**bad)
