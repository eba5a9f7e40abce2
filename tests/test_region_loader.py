"""N3: ltr_region_collect (csrc/host/region_loader.cpp) on the reference's shipped trio reads and BED regions
(BASELINE.json configs[0] / [1]) against (a) the reference's own input layer -- BamCramReader / BamAlignment::TrimAlignment
compiled in place, running on the library's BAM reader through integration/hts_compat.cpp -- and (b) the line-by-line
restatement of the region loop in oracle/pyregion.py.  Host only; runs where /root/reference is mounted."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from longtr_b200 import abi  # noqa: E402
from oracle import pyregion as pr  # noqa: E402

DATA = os.path.join(os.environ.get("LONGTR_REFERENCE", "/root/reference"), "test_data")
SAMPLES = ["HG002", "HG003", "HG004"]
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(DATA, "HG002_sample_reads.bam")),
                                reason="reference test data not mounted")


def bam_path(s):
    return os.path.join(DATA, s + "_sample_reads.bam")


@pytest.fixture(scope="module")
def world():
    import real_cases
    bams = [abi.BamFile(bam_path(s)) for s in SAMPLES]
    tid = [i for i, (n, _) in enumerate(bams[0].refs) if n == "chr1"][0]
    reads = [b.fetch(tid, 0, 1 << 29, keep_raw=True) for b in bams]
    allr = [r for rs in reads for r in rs]
    lo, hi = min(r["pos"] for r in allr), max(r["end"] for r in allr)
    ref = real_cases.build_pseudo_reference([dict(r, seq=r["seq"].upper()) for r in allr], lo, hi)
    regions = [r for r in real_cases.regions() if r["chrom"] == "chr1"]
    assert len(regions) >= 30
    return dict(bams=bams, tid=tid, reads=reads, ref=ref, ref_start=lo, regions=regions)


@pytest.mark.skipif(not pr.ref_io_available(), reason="oracle/_ref/libltr_ref_io.so not built")
def test_reference_reader_runs_on_our_bam_layer(world):
    """BamCramReader::SetRegion / GetNextAlignment of the reference, linked against hts_compat.cpp instead of htslib,
    yields the records ltr_bam_fetch yields (the reference stops at the first record starting behind end + 1)."""
    n = 0
    for reg in world["regions"][::3]:
        q0, q1 = max(0, reg["start"] - 1000), reg["stop"] + 1000
        for s, bam in zip(SAMPLES, world["bams"]):
            want = [r for r in bam.fetch(world["tid"], q0, q1) if r["pos"] <= q1 + 1]
            got = pr.ref_io_region(bam_path(s), "chr1", q0, q1)
            assert [(g["name"], g["pos"], g["end"], g["mapq"], g["hp"]) for g in got] == \
                   [(w["name"], w["pos"], w["end"], w["mapq"], w["hp"]) for w in want]
            for g, w in zip(got, want):
                assert g["cigar"] == "".join("%d%s" % (k, op) for op, k in w["cigar"])
                assert g["seq"] == w["seq"] and g["qual"] == w["qual"] and g["rev"] == ((w["flag"] >> 4) & 1)
            n += len(got)
    assert n > 500


@pytest.mark.skipif(not pr.ref_io_available(), reason="oracle/_ref/libltr_ref_io.so not built")
def test_trim_restatement_is_the_references(world):
    """oracle/pyregion.trim_alignment == BamAlignment::TrimAlignment on every spanning read of every region."""
    n = n_del = 0
    for reg in world["regions"]:
        lo, hi = (reg["start"] - 200 if reg["start"] > 200 else 1), reg["stop"] + 200
        for s, bam in zip(SAMPLES, world["bams"]):
            q0, q1 = max(0, reg["start"] - 1000), reg["stop"] + 1000
            mine = [pr.trim_alignment(r, lo, hi) for r in bam.fetch(world["tid"], q0, q1)
                    if r["pos"] <= q1 + 1 and r["pos"] <= reg["start"] and r["end"] >= reg["stop"]]
            ref = pr.ref_io_region(bam_path(s), "chr1", q0, q1, span=(reg["start"], reg["stop"]), trim=(lo, hi))
            assert len(mine) == len(ref)
            for m, g in zip(mine, ref):
                cig = "".join("%d%s" % (k, op) for op, k in m["cigar"]) or "*"
                assert (m["pos"], m["end"], cig, m["seq"], m["qual"], int(m["deleted"])) == \
                       (g["pos"], g["end"], g["cigar"], g["seq"], g["qual"], g["deleted"])
                n += 1
                n_del += g["deleted"]
    assert n > 800


@pytest.mark.parametrize("which", ["trio", "single"])
def test_region_collect_matches_the_restatement(world, which):
    files = [0, 1, 2] if which == "trio" else [0]
    bams = [world["bams"][f] for f in files]
    n_reads = n_regions = n_phased = 0
    for reg in world["regions"]:
        got = abi.region_collect(bams, "chr1", reg["start"], reg["stop"], world["ref"], world["ref_start"])
        q0, q1 = max(0, reg["start"] - 1000), reg["stop"] + 1000
        per_file = [[r for r in world["reads"][f] if r["pos"] < q1 and r["end"] > q0] for f in files]
        samples, by_sample, cnt = pr.filter_and_order(per_file, reg["start"], reg["stop"])
        terms = pr.phasing_terms(by_sample)
        want, failed = pr.left_align(samples, by_sample, terms, reg["start"], reg["stop"], world["ref"], world["ref_start"])
        assert got["samples"] == samples
        cnt["n_trim_failed"] = failed
        assert got["counters"] == cnt
        assert len(got["reads"]) == len(want)
        for g, w in zip(got["reads"], want):
            assert g == w
        n_reads += len(want)
        n_regions += bool(want)
        n_phased += sum(1 for w in want if w["log_p1"] != w["log_p2"])
    assert n_regions >= 25 and n_reads > (600 if which == "trio" else 200) and n_phased > 50


def test_region_collect_options_and_errors(world):
    reg = world["regions"][0]
    bams = world["bams"][:1]
    base = abi.region_collect(bams, "chr1", reg["start"], reg["stop"], world["ref"], world["ref_start"])
    strict = abi.region_collect(bams, "chr1", reg["start"], reg["stop"], world["ref"], world["ref_start"], min_mapq=61.0)
    assert strict["counters"]["n_low_mapq"] > 0 and len(strict["reads"]) < max(1, len(base["reads"]))
    unphased = abi.region_collect(bams, "chr1", reg["start"], reg["stop"], world["ref"], world["ref_start"], phased_bam=0)
    assert all(r["log_p1"] == 0.0 and r["log_p2"] == 0.0 for r in unphased["reads"])
    with pytest.raises(RuntimeError):
        abi.region_collect(bams, "chrNope", reg["start"], reg["stop"], world["ref"], world["ref_start"])
    with pytest.raises(RuntimeError):  # reference slice that does not cover the reads
        abi.region_collect(bams, "chr1", reg["start"], reg["stop"], "ACGT", world["ref_start"])


def test_a_record_whose_cigar_does_not_fit_its_sequence_refuses_the_region(tmp_path):
    """Damaged files must end in an error code, not in an exception across the C ABI: a record whose CIGAR describes more bases
    than it carries (the reference would read past its strings) makes ltr_region_collect return LTR_ERR_INVALID."""
    import bam_writer as bw
    seq = "ACGT" * 300
    good = bw.encode_record(0, 1000, "r1", 0, 60, [("=", 1200)], seq, "I" * 1200)
    bad = bw.encode_record(0, 1000, "r2", 0, 60, [("=", 2000)], seq, "I" * 1200)   # CIGAR longer than the sequence
    ref = "ACGT" * 2000
    for k, recs in enumerate(([good], [good, bad])):
        path = str(tmp_path / ("f%d.bam" % k))
        bw.write_bam(path, [("chrS", 8000)], recs)
        b = abi.BamFile(path)
        b.build_index()
        if k == 0:
            assert len(abi.region_collect([b], "chrS", 1500, 1560, ref, 0)["reads"]) == 1
        else:
            with pytest.raises(RuntimeError):
                abi.region_collect([b], "chrS", 1500, 1560, ref, 0)


def _random_cigar_world(tmp_path, seed, n_reads=60):
    """One BAM file of reads with random CIGARs around the region [3000, 3060): match runs, insertions and deletions of every
    size -- some exactly at the trimming boundaries (2800 / 3260), some deleting the whole repeat --, soft / hard clips."""
    import random
    import bam_writer as bw
    rng = random.Random(seed)
    ref = "".join(rng.choice("ACGT") for _ in range(6000))
    recs = []
    for k in range(n_reads):
        pos = rng.randrange(2300, 2990)
        cigar, seq, p = [], [], pos
        if rng.random() < 0.2:
            cigar.append(("H", rng.randrange(1, 30)))
        if rng.random() < 0.25:
            n = rng.randrange(1, 25)
            cigar.append(("S", n))
            seq.append("".join(rng.choice("ACGT") for _ in range(n)))
        want_end = rng.randrange(3070, 3700)
        last = None
        while p < want_end:
            x = rng.random()
            near = min(abs(p - b) for b in (2800, 3000, 3060, 3260)) < 15
            if last in (None, "I", "D") or x < (0.5 if near else 0.8):
                n = rng.randrange(1, 12 if near else 90)
                op = rng.choice("M=X") if rng.random() < 0.5 else "M"
                s = ref[p:p + n]
                if op != "=":
                    s = "".join(c if rng.random() < 0.9 else rng.choice("ACGT") for c in s)
                seq.append(s)
                p += n
            elif x < (0.75 if near else 0.9):
                op, n = "I", rng.randrange(1, 14)
                seq.append("".join(rng.choice("ACGT") for _ in range(n)))
            else:
                op, n = "D", (rng.randrange(70, 120) if rng.random() < 0.1 else rng.randrange(1, 14))
                p += n
            if op == last or (op in "M=X" and last in ("M", "=", "X") and op == cigar[-1][0]):
                cigar[-1] = (op, cigar[-1][1] + n)
            else:
                cigar.append((op, n))
            last = op
        if last not in ("M", "=", "X"):
            cigar.append(("M", 5))
            seq.append(ref[p:p + 5])
            p += 5
        if rng.random() < 0.25:
            n = rng.randrange(1, 25)
            cigar.append(("S", n))
            seq.append("".join(rng.choice("ACGT") for _ in range(n)))
        if rng.random() < 0.2:
            cigar.append(("H", rng.randrange(1, 30)))
        s = "".join(seq)
        recs.append((pos, bw.encode_record(0, pos, "rd%03d" % k, 16 if k % 3 == 0 else 0, 60, cigar, s, "I" * len(s),
                                           hp=(1 + k % 2) if k % 4 else None)))
    recs.sort(key=lambda t: t[0])
    path = str(tmp_path / ("cigars%d.bam" % seed))
    bw.write_bam(path, [("chrR", len(ref))], [r for _p, r in recs])
    return path, ref


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_trimming_on_random_cigars(tmp_path, seed):
    """The library's trimming consumes whole CIGAR operations where the reference walks base by base: on reads with random
    CIGARs (indels at the trimming boundaries, deleted repeats, clips) ltr_region_collect must leave the same reads as the
    base-by-base restatement (oracle/pyregion.trim_alignment, itself pinned by the reference's BamAlignment::TrimAlignment on
    the shipped reads above; the reference's reader needs a .bai on disk, which these generated files do not have)."""
    path, ref = _random_cigar_world(tmp_path, seed)
    bam = abi.BamFile(path)
    bam.build_index()
    start, stop = 3000, 3060
    reads = bam.fetch(0, 0, 1 << 29, keep_raw=True)
    assert sum(1 for r in reads if r["pos"] <= start and r["end"] >= stop) >= 50
    got = abi.region_collect([bam], "chrR", start, stop, ref, 0, min_mean_qual=0.0)
    samples, by_sample, cnt = pr.filter_and_order([reads], start, stop, min_mean_qual=0.0)
    terms = pr.phasing_terms(by_sample)
    want, failed = pr.left_align(samples, by_sample, terms, start, stop, ref, 0)
    cnt["n_trim_failed"] = failed
    assert got["counters"] == cnt and len(got["reads"]) == len(want) >= 15
    for g, w in zip(got["reads"], want):
        assert g == w
    bam.close()
