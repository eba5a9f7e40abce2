#!/usr/bin/env python
"""Generates tests/golden/*.json from the UNMODIFIED reference sources (oracle/_ref).

Run in the build container (needs /root/reference to build oracle/_ref):

    python tools/make_golden.py

The reference ships no golden vectors for this path, so the fixtures are
outputs of the reference's own HapAligner::process_reads / Genotyper::calc_log_sample_posteriors
on seeded inputs.  Doubles are stored as C99 hex-float strings (bit exact).  The inputs are stored
next to the outputs so the fixtures can be replayed anywhere (GPU box included) without the
reference tree.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import synth  # noqa: E402
from longtr_b200.flat import make_flat_locus  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

ONT = (-1.0, -0.458675, -1.0, -0.458675, -0.0202027, -4.60517, -4.60517)
ODD = (-0.7, -0.61, -0.35, -1.3, -0.013, -3.9, -4.4)


def hexf(a):
    return [float(x).hex() for x in np.asarray(a, dtype=np.float64).ravel()]


def appendix_a():
    """SURVEY.md Appendix A1-A3 inputs."""
    lf = "GATTACAGGCTTAACGTGCATCGATCGGATCCATG"
    rf = "TTGACCGTAGGCTAATCGGATATCGCGTATAGCCA"
    pr = "GGTCT"
    out = []
    pl = "CTGAA"
    alleles = [pl + "AC" * 10 + pr, pl + "AC" * 8 + pr, pl + "AC" * 12 + pr]
    seq = "T" * 50 + lf + pl + "AC" * 12 + pr + rf + "G" * 50
    out.append(dict(name="A1", lflank=lf, rflank=rf, alleles=alleles, repeat_start=1035, repeat_end=1065,
                    period=2, motif="AC", switch=0, aln_params=None,
                    reads=[dict(start=950, stop=1149, seq=seq, qual="I" * len(seq), cigar="110=4I90=")]))
    pl = "CTGAC"
    alleles = [pl + "A" * 12 + pr, pl + "A" * 11 + pr, pl + "A" * 13 + pr]
    seq = "T" * 50 + lf + pl + "A" * 13 + pr + rf + "G" * 50
    for name, sw in (("A2", 0), ("A3", 20)):
        out.append(dict(name=name, lflank=lf, rflank=rf, alleles=alleles, repeat_start=1035, repeat_end=1057,
                        period=1, motif="A", switch=sw, aln_params=None,
                        reads=[dict(start=950, stop=1141, seq=seq, qual="?" * len(seq), cigar="102=1I90=")]))
    return out


def run_ref(case):
    reads = [(r["start"], r["stop"], r["seq"], r["qual"], r["cigar"]) for r in case["reads"]]
    L, keep = make_flat_locus(case["lflank"], case["alleles"], case["rflank"], case["repeat_start"],
                              case["repeat_end"], case["period"], reads, motif=case["motif"],
                              switch_old_align_len=case["switch"], aln_params=case["aln_params"],
                              realign_to_hap=case.get("realign_to_hap"), realign_read=case.get("realign_read"))
    ll, seeds, _ = po.process_reads(L, len(reads), len(case["alleles"]), fill=case.get("fill", 0.0), which="ref")
    case["ll"] = hexf(ll)
    case["seeds"] = [int(s) for s in seeds]
    # HapAligner::calc_seed_base of every read (used by the homopolymer path; recorded for all loci)
    case["seed_bases"] = [int(s) for s in po.ref_seed_bases(L, len(reads))]
    return case


def long_path_cases():
    cases = []
    for seed in range(40):
        kw = dict(n_reads=8)
        params = None
        if seed % 4 == 1:
            kw.update(sub=0.01, indel=0.02)
            params = ONT
        if seed % 4 == 2:
            kw.update(ref_len=int(150 + 10 * seed), n_reads=5)
        if seed % 4 == 3:
            params = ODD
            kw.update(sub=0.03, indel=0.03)
        if seed % 10 == 9:
            kw.update(homopolymer=True)
        loc = synth.make_locus(1000 + seed, **kw)
        case = dict(name="long%02d" % seed, lflank=loc["lflank"], rflank=loc["rflank"], alleles=loc["alleles"],
                    repeat_start=loc["repeat_start"], repeat_end=loc["repeat_end"], period=loc["period"],
                    motif=loc["motif"], switch=0, aln_params=list(params) if params else None, reads=loc["reads"])
        if seed % 7 == 3:
            rng = np.random.default_rng(seed)
            case["realign_to_hap"] = [bool(x) for x in (rng.random(len(loc["alleles"])) < 0.7)]
            case["realign_read"] = [bool(x) for x in (rng.random(len(loc["reads"])) < 0.7)]
            case["fill"] = 7.25
        cases.append(run_ref(case))
    return cases


def short_path_cases():
    """Homopolymer loci with --stutter-align-len (HapAligner::process_reads short path)."""
    cases = []
    for seed in range(24):
        kw = dict(n_reads=6, homopolymer=True, ref_len=int(8 + (5 * seed) % 30), sub=0.004 * (seed % 3),
                  indel=0.01 * (seed % 4), ctx=40 + 10 * (seed % 5))
        params = ONT if seed % 5 == 4 else None
        loc = synth.make_locus(3000 + seed, **kw)
        case = dict(name="short%02d" % seed, lflank=loc["lflank"], rflank=loc["rflank"], alleles=loc["alleles"],
                    repeat_start=loc["repeat_start"], repeat_end=loc["repeat_end"], period=1, motif=loc["motif"],
                    switch=20, aln_params=list(params) if params else None, reads=loc["reads"])
        if seed % 6 == 5:
            rng = np.random.default_rng(seed)
            case["realign_to_hap"] = [bool(x) for x in (rng.random(len(loc["alleles"])) < 0.7)]
            case["realign_read"] = [bool(x) for x in (rng.random(len(loc["reads"])) < 0.7)]
            case["fill"] = 7.25
        if seed % 8 == 7:  # a read without any '=' run long enough -> no seed -> all-zero row
            r = case["reads"][0]
            n = len(r["seq"])
            r["cigar"] = "%dX" % n
            r["stop"] = r["start"] + n - 1
        cases.append(run_ref(case))
    return cases


def posterior_cases():
    rng = np.random.default_rng(20260117)
    out = []
    for t in range(24):
        S, H = int(rng.integers(1, 4)), int(rng.integers(1, 9))
        haploid = (t % 5 == 0)
        rps = rng.integers(1, 15, size=S)
        lab = np.repeat(np.arange(S), rps).astype(np.int32)
        R = len(lab)
        ll = -rng.exponential(20, size=(R, H))
        ll[rng.random((R, H)) < 0.05] = -700
        ll[rng.random((R, H)) < 0.02] = -1e9
        hp = rng.integers(0, 3, size=R)
        p1 = np.where(hp == 0, -1e-6, np.where(hp == 1, -1000.0, 0.0))
        p2 = np.where(hp == 0, -1000.0, np.where(hp == 1, -1e-6, 0.0))
        cl, post, tot, total, best = po.log_sample_posteriors(ll, p1, p2, lab, S, haploid=haploid, which="ref")
        out.append(dict(name="post%02d" % t, S=S, H=H, haploid=haploid, label=[int(x) for x in lab], ll=hexf(ll),
                        log_p1=hexf(p1), log_p2=hexf(p2), ll_clamped=hexf(cl), post=hexf(post), totals=hexf(tot),
                        total=float(total).hex(), best=[int(x) for x in best.ravel()]))
    return out


def calls_cases():
    """calc_log_sample_posteriors + extract_genotypes_and_likelihoods (GT, Q, PQ, GL, PL, PHASEDGL, GLDIFF)."""
    rng = np.random.default_rng(20260118)
    out = []
    for t in range(30):
        S, H = int(rng.integers(1, 4)), int(rng.integers(1, 7))
        haploid = (t % 6 == 0)
        rps = [int(x) for x in rng.integers(1, 15, size=S)]
        R = sum(rps)
        true = rng.integers(0, H, size=(S, 2))
        lab = np.repeat(np.arange(S), rps)
        ll = -rng.exponential(25, size=(R, H)) - 5
        hp = rng.integers(0, 3, size=R)
        for r in range(R):  # reads support one of the sample's two alleles
            ll[r, true[lab[r], hp[r] % 2]] = -rng.exponential(0.3)
        ll[rng.random((R, H)) < 0.03] = -700
        p1 = np.where(hp == 0, -1e-6, np.where(hp == 1, -1000.0, 0.0))
        p2 = np.where(hp == 0, -1000.0, np.where(hp == 1, -1e-6, 0.0))
        ref = po.ref_genotype_locus(ll, p1, p2, rps, haploid=haploid)
        case = dict(name="calls%02d" % t, S=S, H=H, haploid=haploid, reads_per_sample=rps, ll=hexf(ll),
                    log_p1=hexf(p1), log_p2=hexf(p2), total_ll=float(ref["total_ll"]).hex())
        for k, v in ref.items():
            if k in ("total_ll",):
                continue
            case["out_" + k] = [int(x) for x in v.ravel()] if v.dtype == np.int32 else hexf(v)
        out.append(case)
    return out


def pair_batch_cases():
    """Kernel-level batches (full haplotypes + trimmed reads) through the reference classes."""
    out = []
    for seed, kw, params in ((11, dict(n_loci=12, n_lo=20, n_hi=160, flank=30, weird=0.0), None),
                             (12, dict(n_loci=8, n_lo=100, n_hi=420, flank=30, weird=0.0, sub=0.02, indel=0.03), ONT),
                             (13, dict(n_loci=10, n_lo=20, n_hi=90, flank=30, weird=0.0, sub=0.05, indel=0.05), ODD)):
        b = synth.make_pair_batch(seed, **kw)
        # the reference driver needs reads >= 11 bp and haplotypes with 35 bp flanks (synth flank=30 -> 35)
        ll, _sec = po.ref_viterbi_batch(b, params, n_threads=4)
        out.append(dict(name="pairs%d" % seed, aln_params=list(params) if params else None,
                        locus_hap_begin=[int(x) for x in b["locus_hap_begin"]],
                        locus_read_begin=[int(x) for x in b["locus_read_begin"]],
                        haps=b["haps"], reads=b["reads"], ll=hexf(ll)))
    return out


def vcf_record_cases():
    """Whole loci through the reference's SeqStutterGenotyper (ctor -> genotype -> write_vcf_record), IO-less."""
    import dropin_cases as dc
    cases = [dc.case_a4()] + dc.seeded_cases()
    recs = po.full_locus_records(cases, "full")
    assert recs[0] == dc.A4_RECORD, "SURVEY Appendix A4 not reproduced"
    return [dict(name=c["name"], record=r) for c, r in zip(cases, recs)]


def pooling_cases():
    """ReadPooler + median base qualities as the reference computes them (oracle/_ref: ltr_ref_pool_reads)."""
    from longtr_b200 import abi
    out = []
    for seed in range(12):
        rng = np.random.default_rng(7700 + seed)
        n_distinct = int(rng.integers(1, 8))
        seqs_d = ["".join("ACGT"[int(x)] for x in rng.integers(0, 4, size=int(rng.integers(5, 60)))) for _ in range(n_distinct)]
        if seed % 4 == 0 and n_distinct > 1:
            seqs_d[1] = seqs_d[0][:-1] + ("A" if seqs_d[0][-1] != "A" else "C")   # differs in one base only
        n = int(rng.integers(1, 30))
        pick = rng.integers(0, n_distinct, size=n)
        seqs = [seqs_d[int(k)] for k in pick]
        quals = ["".join(chr(int(q)) for q in rng.integers(33, 75, size=len(s))) for s in seqs]
        pool, meds = abi.pool_reads(seqs, quals, lib=po.ref_lib(), fn="ltr_ref_pool_reads")
        out.append(dict(name="pool%02d" % seed, seqs=seqs, quals=quals, pool_index=pool, median_quals=meds))
    return out


def pruning_cases():
    """Both passes of SeqStutterGenotyper::genotype (src/seq_stutter_genotyper.cpp:634-645) as the reference ran them
    (oracle/_ref/ltr_ref_trace records every call of Genotyper::calc_log_sample_posteriors): the LL matrix of all candidate
    alleles going in, and -- after get_unused_alleles / remove_alleles -- the surviving alleles, their LL columns, the
    recomputed posteriors and the optimal pairs.  Seeded loci with an uncalled third allele + the shipped-data loci in
    which the reference pruned something + two loci in which it did not."""
    import dropin_cases as dc
    import golden_util as gu
    cases = dc.pruning_cases() + [dc.case_a4()] + dc.seeded_cases()[:1] + gu.load_real_cases()
    out = []
    for c, (rec, calls) in zip(cases, po.full_locus_traces(cases)):
        if not calls:
            continue
        real = not c["name"].startswith(("prune", "A4", "dropin"))
        if real and len(calls) < 2:
            continue
        first, last = calls[0], calls[-1]
        assert len(set(first["alleles"])) == len(first["alleles"])
        kept = [first["alleles"].index(a) for a in last["alleles"]]
        labels = first["labels"]
        S = first["S"]
        out.append(dict(name=c["name"], S=S, H=first["H"], R=first["R"], n_calls=len(calls), haploid=bool(c.get("haploid")),
                        reads_per_sample=[labels.count(s) for s in range(S)], seeds=first["seeds"],
                        ll=first["ll"], log_p1=first["p1"], log_p2=first["p2"], first_gts=first["gts"],
                        first_post=first["post"], kept=kept, out_ll=last["ll"], out_post=last["post"],
                        out_totals=last["totals"], out_gts=last["gts"],
                        # what the reference's HaplotypeGenerator built for the locus: the input of ltr_genotyper_run
                        lflank=first["lflank"], rflank=first["rflank"], repeat_start=first["repeat_start"],
                        repeat_end=first["repeat_end"], alleles=first["alleles"],
                        aln_params=list(c["aln_params"]) if c.get("aln_params") else None,
                        reads=[dict(start=r["start"], stop=r["stop"], seq=r["seq"], cigar=r["cigar"], sample=r["sample"],
                                    log_p1=r["log_p1"], log_p2=r["log_p2"]) for r in c["reads"]] if not real else None,
                        real_case=c["name"] if real else None))
    return out


def main():
    if not po.ref_available():
        raise SystemExit("oracle/_ref is not built (needs /root/reference)")
    os.makedirs(GOLD, exist_ok=True)
    makers = dict(appendix_a=lambda: [run_ref(c) for c in appendix_a()], process_reads_long=long_path_cases,
                  process_reads_short=short_path_cases, posteriors=posterior_cases, pair_batches=pair_batch_cases,
                  calls=calls_cases, vcf_records=vcf_record_cases, pruning=pruning_cases, pooling=pooling_cases)
    only = [a for a in sys.argv[1:] if not a.startswith("-")]  # `python tools/make_golden.py pruning`: that set only
    sets = {k: f() for k, f in makers.items() if not only or k in only}
    for name, cases in sets.items():
        path = os.path.join(GOLD, name + ".json")
        with open(path, "w") as f:
            json.dump(dict(generator="tools/make_golden.py", source="oracle/_ref (reference sources compiled in place)",
                           cases=cases), f, indent=0, separators=(",", ":"))
        print(path, len(cases), "cases", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
