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


# ---- output switches (ltr_vcf_record_ex / ltr_vcf_header_ex) -------------------------------------------------------------
def load_switch_records(key="masks"):
    import gzip
    import json
    import os
    with gzip.open(os.path.join(gu.GOLD, "vcf_switches.json.gz"), "rt") as f:
        return {int(m): v for m, v in json.load(f)[key].items()}


def extras_from_record(record, cand, kept, n_samples, haploid=False):
    """GL / PL / PHASEDGL of a reference record, moved from the record's allele order back into the layout ltr_batch_calls
    delivers (kept alleles in candidate order; pair (a <= b) at b(b+1)/2 + a; phased [a * K + b])."""
    f = record.split("\t")
    fmt = f[8].split(":")
    kept_idx = [k for k in range(len(kept)) if kept[k]]
    K = len(kept_idx)
    al = [cand["alleles"][k] for k in kept_idx]
    new_to_old = [0] + sorted(range(1, K), key=lambda i: (len(al[i]), al[i]))
    n_gl, n_pgl = (K if haploid else K * (K + 1) // 2), (K if haploid else K * K)
    gls, pls, pgls = np.zeros(n_samples * n_gl), np.zeros(n_samples * n_gl, dtype=np.int32), np.zeros(n_samples * n_pgl)
    for s in range(n_samples):
        w = dict(zip(fmt, f[9 + s].split(":")))
        for key, dst, conv in (("GL", gls, float), ("PL", pls, int)):
            if key not in w:
                continue
            v = [conv(x) for x in w[key].split(",")]
            assert len(v) == n_gl
            if haploid:
                for i in range(K):
                    dst[s * n_gl + new_to_old[i]] = v[i]
            else:
                k = 0
                for i in range(K):
                    for j in range(i + 1):
                        lo, hi = sorted((new_to_old[i], new_to_old[j]))
                        dst[s * n_gl + hi * (hi + 1) // 2 + lo] = v[k]
                        k += 1
        if "PHASEDGL" in w:
            v = [float(x) for x in w["PHASEDGL"].split(",")]
            assert len(v) == K * K
            for i in range(K):
                for j in range(K):
                    pgls[s * n_pgl + new_to_old[i] * K + new_to_old[j]] = v[i * K + j]
    return dict(gl_begin=[s * n_gl for s in range(n_samples + 1)], gls=gls, pls=pls,
                pgl_begin=[s * n_pgl for s in range(n_samples + 1)], phased_gls=pgls)


def strip_het_mallreads_any(record):
    return strip_het_mallreads(record) if "MALLREADS" in record.split("\t")[8].split(":") else record


def test_records_under_the_output_switches():
    """Every switch combination of tests/golden/vcf_switches.json.gz (the reference run with --hide-allreads / --hide-mallreads /
    --output-gls / --output-pls / --output-phased-gls / --output-filters): the FORMAT column, the order of GL / PL / PHASEDGL
    after the alleles are re-ordered, PASS -- character for character (the likelihoods themselves are read back from the
    record here; tests/test_gpu_real_data.py computes them on the device)."""
    gold = load_switch_records()
    cases = gu.load_real_cases()
    n, n_multi = 0, 0
    for c in cases:
        cand = abi.candidate_alleles_from_reads(c["reads"], len(c["samples"]), c["region_start"], c["region_stop"],
                                                len(c["motif"]), c["chrom_seq"])
        for mask, recs in sorted(gold.items()):
            want = recs[c["name"]]
            inp = record_inputs(dict(c, record=want), cand)
            if inp is None:
                continue
            S = len(c["samples"])
            if any(inp["gts"][2 * s] != inp["gts"][2 * s + 1] for s in range(S)):
                inp["read_allele"] = [inp["gts"][2 * s] for s in inp["read_sample"]]
            got = abi.vcf_record(switches=mask, **inp, **extras_from_record(want, cand, inp["kept_mask"], S))
            assert strip_het_mallreads_any(got) == strip_het_mallreads_any(want), (c["name"], mask)
            n += 1
            n_multi += got.split("\t")[4].count(",") >= 1
    assert n >= 4 * 45 and n_multi >= 8, (n, n_multi)


def test_default_switches_and_missing_columns():
    c = gu.load_real_cases()[0]
    cand = abi.candidate_alleles_from_reads(c["reads"], len(c["samples"]), c["region_start"], c["region_stop"], len(c["motif"]),
                                            c["chrom_seq"])
    inp = record_inputs(c, cand)
    S = len(c["samples"])
    assert abi.vcf_record(switches=abi.VCF_DEFAULT, **inp) == abi.vcf_record(**inp) == c["record"]
    # a column whose sample has no reads at the locus: "." or, with FILTER, the empty fields + NO_READS
    # (seq_stutter_genotyper.cpp:1196-1215)
    inp2 = dict(inp, column_sample=list(range(S)) + [-1])
    assert abi.vcf_record(**inp2).split("\t")[-1] == "."
    got = abi.vcf_record(switches=abi.VCF_DEFAULT | abi.VCF_FILTERS, **inp2).split("\t")
    assert got[8].endswith(":MALLREADS:FILTER") and got[-1] == ".:" * 12 + "NO_READS" and got[-2].endswith(":PASS")
    with pytest.raises(RuntimeError):   # GL switched on without the likelihoods
        abi.vcf_record(switches=abi.VCF_GLS, **inp)
    with pytest.raises(RuntimeError):   # unknown switch
        abi.vcf_record(switches=1 << 9, **inp)


def test_haploid_records_under_the_output_switches():
    """The seeded loci as a haploid chromosome (tests/dropin_cases.py::haploid_cases) under every switch combination: FORMAT
    GT:GB:Q:DP:DFLANKINDEL:GLDIFF, GL / PL per allele, PHASEDGL never shown (seq_stutter_genotyper.cpp:1186, :1311-1323)."""
    import dropin_cases as dc
    gold = load_switch_records("haploid")
    n = 0
    for c in dc.haploid_cases():
        cand = abi.candidate_alleles_from_reads(c["reads"], len(c["samples"]), c["region_start"], c["region_stop"],
                                                len(c["motif"]), c["chrom_seq"])
        lens = [len(a) for a in cand["alleles"]]
        for mask, recs in sorted(gold.items()):
            want = recs[c["name"]]
            f = want.split("\t")
            fmt = f[8].split(":")
            assert "PHASEDGL" not in fmt and "PQ" not in fmt
            S = len(c["samples"])
            gts, lup, gld, kept = [], [], [], [1] + [0] * (len(lens) - 1)
            for s in range(S):
                w = dict(zip(fmt, f[9 + s].split(":")))
                k = [k for k in range(len(lens)) if lens[k] - lens[0] == int(w["GB"])]
                assert len(k) == 1
                gts.append([k[0], k[0]])
                kept[k[0]] = 1
                lup.append(np.log(max(float(w["Q"]), 1e-300)))
                gld.append(0.0 if w["GLDIFF"] == "." else float(w["GLDIFF"]))
            if sum(kept) - 1 != (0 if f[4] == "." else f[4].count(",") + 1):
                continue
            calls = dict(gts=gts, lup=lup, lpp=lup, gld=gld, kept=kept, read_allele=None)
            inp = record_inputs(dict(c, record=want), cand, calls)
            got = abi.vcf_record(haploid=True, switches=mask, **inp, **extras_from_record(want, cand, kept, S, haploid=True))
            assert got == want, (c["name"], mask)
            n += 1
    assert n >= 20, n
