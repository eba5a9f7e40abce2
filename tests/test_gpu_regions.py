"""N3 end to end on the GPU: synthetic BAM files (tests/bam_writer.py, rebuilt here from the generator's seeds) ->
ltr_regions_run (BAM reader, region loop, candidate alleles on host threads; alignment, posteriors, removal of uncalled
alleles on the device) against what the reference's own genotyper made of the same reads (tests/golden/regions.json,
recorded by tools/make_region_golden.py from oracle/_ref/ltr_ref_trace): same verdict per region, same candidate alleles,
same surviving alleles, same optimal pairs, same posteriors -- including the regions whose candidate alleles come from the
assembly branch (clustering + partial-order consensus; the reference ran it on the restated spoa, see tools/make_region_golden.py)."""
import json
import os

import numpy as np
import pytest

import bam_writer as bw
import golden_util as gu
from longtr_b200 import abi
from longtr_b200.locus_batch import Genotyper

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "regions.json")


@pytest.fixture(scope="module")
def genotyper():
    g = Genotyper(devices=(0,), host_threads=8, chunk_loci=16)
    yield g
    g.close()


@pytest.mark.parametrize("wi", [0, 1])
def test_regions_run_matches_the_reference(genotyper, tmp_path, wi):
    W = json.load(open(GOLD))["worlds"][wi]
    world = bw.synthetic_world(W["n_loci"], config=3, first_locus=W["first_locus"], n_samples=W["n_samples"])
    bams = [abi.BamFile(p) for p in bw.write_world(world, str(tmp_path))]
    for b in bams:
        b.build_index()
    motifs = [world["chrom_seq"][s0:s0 + per] for s0, _e, per in world["regions"]]
    names = ["R%d" % k for k in range(len(world["regions"]))]
    out = genotyper.run_regions(bams, "chrS", world["regions"], world["chrom_seq"], 0, motifs=motifs, names=names)
    calls = out["calls"]
    n_checked = n_records = 0
    for g in W["regions"]:
        r = g["region"]
        if g["status"] != 0:
            assert out["status"][r] == {2: 5, 3: 6, 1: 3}[g["status"]], (r, out["status"][r], g["status"])
            assert out["locus_index"][r] == -1
            continue
        assert out["status"][r] == 0
        assert out["alleles"][r] == g["alleles"] and list(out["block"][r]) == g["block"] and out["samples"][r] == g["samples"]
        assert out["inexact"][r] == g["inexact"]
        if g.get("reference_failed"):
            continue
        l = out["locus_index"][r]
        assert calls["status"][l] == 0
        a0, a1 = calls["locus_allele_begin"][l], calls["locus_allele_begin"][l + 1]
        kept = list(np.nonzero(calls["kept_mask"][a0:a1])[0])
        assert kept == g["kept"], r
        s0, s1 = calls["locus_sample_begin"][l], calls["locus_sample_begin"][l + 1]
        S, K = g["S"], len(kept)
        assert s1 - s0 == S
        assert list(calls["gts"][s0:s1].ravel()) == [kept[x] for x in g["out_gts"]], r
        post = gu.unhex(g["out_post"], (S, K, K))
        want = [post[s, g["out_gts"][2 * s], g["out_gts"][2 * s + 1]] for s in range(S)]
        np.testing.assert_allclose(calls["log_phased_posteriors"][s0:s1], want, rtol=1e-10, atol=1e-9, err_msg=str(r))
        np.testing.assert_allclose(calls["sample_total_lls"][s0:s1], gu.unhex(g["out_totals"]), rtol=1e-10, atol=1e-9)
        n_checked += 1
        # the VCF record: the reference lists the samples in the region's order, the library one column per BAM file
        f = g["record"].split("\t")
        by_file = {g["samples"][k]: f[9 + k] for k in range(S)}
        want_rec = "\t".join(f[:9] + [by_file.get(b, ".") for b in range(len(bams))])
        assert out["records"][r] == want_rec, r
        n_records += 1
    assert n_checked >= (100 if wi == 0 else 60) and n_records == n_checked
    assert out["n_assembled"] >= 40 and sum(map(sum, out["inexact"])) >= 10
    # the same regions with the assembly switched off: those that needed it are reported, the others are unchanged
    off = genotyper.run_regions(bams, "chrS", world["regions"], world["chrom_seq"], 0, no_assembly=1)
    n_needs = sum(1 for st in off["status"] if st == 6)
    assert n_needs == out["n_assembled"] and off["n_assembled"] == 0


def test_regions_run_reports_skipped_regions(genotyper, tmp_path):
    world = bw.synthetic_world(6, config=3, n_samples=1)
    bams = [abi.BamFile(p) for p in bw.write_world(world, str(tmp_path))]
    s, e, per = world["regions"][2]
    regions = [(s, e, per), (s, s + 2000, per), (10, 40, 2), (s + 1500, s + 1600, 3), (e, s, per)]
    out = genotyper.run_regions(bams, "chrS", regions, world["chrom_seq"], 0)  # files without index: scanned
    assert out["status"][1] == 2 and out["status"][2] == 3 and out["status"][3] == 4 and out["status"][4] == 1
    assert out["status"][0] in (0, 6)
    strict = genotyper.run_regions(bams, "chrS", regions[:1], world["chrom_seq"], 0, min_total_reads=31)
    assert strict["status"] == [4] and strict["calls"] is None


def test_run_bed_equals_regions_run(genotyper, tmp_path):
    """ltr_run_bed (FASTA file + region file + BAM files -> calls) against ltr_regions_run on the same
    regions with the sequence handed over directly."""
    world = bw.synthetic_world(12, config=3, first_locus=40, n_samples=1)
    paths = bw.write_world(world, str(tmp_path))
    bams = [abi.BamFile(p) for p in paths]
    for b in bams:
        b.build_index()
    fa = tmp_path / "ref.fa"
    with open(fa, "w") as f:
        f.write(">chrOther\n" + "ACGT" * 50 + "\n>chrS description\n")
        s = world["chrom_seq"]
        for k in range(0, len(s), 60):
            f.write(s[k:k + 60] + "\n")
    bed = tmp_path / "regions.bed"
    order = list(range(len(world["regions"])))[::-1]                  # unsorted on purpose
    with open(bed, "w") as f:
        for r in order:
            s0, e0, per = world["regions"][r]
            f.write("chrS\t%d\t%d\t%s\tR%d\n" % (s0 + 1, e0, world["chrom_seq"][s0:s0 + per], r))
    motifs = [world["chrom_seq"][s0:s0 + per] for s0, _e, per in world["regions"]]
    want = genotyper.run_regions(bams, "chrS", world["regions"], world["chrom_seq"], 0, motifs=motifs,
                                 names=["R%d" % r for r in range(len(order))])
    got = genotyper.run_bed(bams, abi.FastaFile(str(fa)), str(bed), vcf_records=True)
    assert got["chroms"] == ["chrS"] and [b[4] for b in got["bed"]] == ["R%d" % r for r in range(len(order))]
    g = got["per_chrom"][0]
    for key in ("status", "locus_index", "block", "alleles", "samples", "inexact", "records"):
        assert g[key] == want[key], key
    assert sum(1 for x in g["records"] if x.startswith("chrS\t")) == sum(1 for st in g["status"] if st == 0) >= 8
    for key in ("gts", "kept_mask", "log_phased_posteriors", "gl_diffs"):
        np.testing.assert_array_equal(g["calls"][key], want["calls"][key])
    # ltr_run_bed_stream: the same regions three at a time through a sink -- same status, alleles and records in region order
    chunks = []
    genotyper.run_bed_stream(bams, abi.FastaFile(str(fa)), str(bed), lambda c, first, res: chunks.append((c, first, res)) and None,
                             chunk_regions=3, vcf_records=True)
    assert [first for _c, first, _r in chunks] == list(range(0, len(order), 3)) and all(c == 0 for c, _f, _r in chunks)
    for key in ("status", "block", "alleles", "samples", "inexact", "records"):
        assert [x for _c, _f, res in chunks for x in res[key]] == g[key], key
    stopped = []
    with pytest.raises(Exception):   # a sink that returns non-zero stops the run
        genotyper.run_bed_stream(bams, abi.FastaFile(str(fa)), str(bed), lambda c, first, res: stopped.append(first) or True,
                                 chunk_regions=3)
    assert stopped == [0]
    with open(bed, "a") as f:
        f.write("chrMissing\t100\t130\tAC\n")
    with pytest.raises(Exception):
        genotyper.run_bed(bams, abi.FastaFile(str(fa)), str(bed))     # chromosome absent from the FASTA / BAM files
    with pytest.raises(Exception):
        genotyper.run_bed_stream(bams, abi.FastaFile(str(fa)), str(bed), lambda *a: None)


def test_example_flow_writes_the_vcf_file(tmp_path):
    """tools/run_bed_to_vcf.py: BAM + FASTA + region file -> VCF file; its records are the golden ones of the world."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import run_bed_to_vcf
    W = json.load(open(GOLD))["worlds"][0]
    world = bw.synthetic_world(W["n_loci"], config=3, first_locus=W["first_locus"], n_samples=W["n_samples"])
    paths = bw.write_world(world, str(tmp_path))
    fa, bed, vcf = tmp_path / "ref.fa", tmp_path / "regions.bed", tmp_path / "calls.vcf"
    with open(fa, "w") as f:
        f.write(">chrS\n")
        for k in range(0, len(world["chrom_seq"]), 80):
            f.write(world["chrom_seq"][k:k + 80] + "\n")
    with open(bed, "w") as f:
        for r, (s0, e0, per) in enumerate(world["regions"]):
            f.write("chrS\t%d\t%d\t%s\tR%d\n" % (s0 + 1, e0, world["chrom_seq"][s0:s0 + per], r))
    n = run_bed_to_vcf.main(["--bams", ",".join(paths), "--fasta", str(fa), "--regions", str(bed), "--out", str(vcf),
                             "--samples", "S0"])
    lines = open(vcf).read().splitlines()
    body = [x for x in lines if not x.startswith("#")]
    assert lines[0] == "##fileformat=VCFv4.1" and lines[len(lines) - len(body) - 1].endswith("FORMAT\tS0")
    want = [g["record"] for g in W["regions"] if g["status"] == 0]
    assert n == len(body) == len(want) and body == want


def test_region_records_under_the_output_switches(genotyper, tmp_path):
    """opts->vcf_switches: the region loop's records with GL / PL / PHASEDGL / FILTER (the reference's --output-gls ... switches;
    the record composer itself is pinned by the reference's records in tests/test_vcf_writer.py / test_gpu_real_data.py): the
    default fields are unchanged, the extra ones have one entry per genotype of the record's alleles, the PL of the most likely
    genotype is 0, and hiding ALLREADS / MALLREADS removes exactly those."""
    world = bw.synthetic_world(12, config=3, first_locus=500, n_samples=2)
    bams = [abi.BamFile(p) for p in bw.write_world(world, str(tmp_path))]
    motifs = [world["chrom_seq"][s0:s0 + per] for s0, _e, per in world["regions"]]
    base = genotyper.run_regions(bams, "chrS", world["regions"], world["chrom_seq"], 0, motifs=motifs)
    full = genotyper.run_regions(bams, "chrS", world["regions"], world["chrom_seq"], 0, motifs=motifs, vcf_switches=63)
    bare = genotyper.run_regions(bams, "chrS", world["regions"], world["chrom_seq"], 0, motifs=motifs, vcf_switches=0)
    n = n_multi = 0
    for r, st in enumerate(base["status"]):
        if st != 0:
            continue
        b, f, z = (x["records"][r].split("\t") for x in (base, full, bare))
        assert f[:8] == b[:8] == z[:8]
        assert f[8] == b[8] + ":GL:PL:PHASEDGL:FILTER" and z[8] + ":ALLREADS:MALLREADS" == b[8]
        K = 1 + (0 if b[4] == "." else b[4].count(",") + 1)
        for cb, cf, cz in zip(b[9:], f[9:], z[9:]):
            if cb == ".":
                assert cf == ".:" * 15 + "NO_READS" and cz == "."
                continue
            wb, wf = cb.split(":"), cf.split(":")
            assert wf[:len(wb)] == wb and wf[-1] == "PASS" and cz.split(":") == wb[:-2]
            gl, pl, pgl = ([float(x) for x in wf[len(wb) + k].split(",")] for k in range(3))
            assert len(gl) == len(pl) == K * (K + 1) // 2 and len(pgl) == K * K
            assert min(pl) == 0 and pl[gl.index(max(gl))] == 0   # PL = min(999, -10 (GL - max GL)), genotyper.cpp:86-89
            n += 1
            n_multi += K > 1
    assert n >= 16 and n_multi >= 8


def test_haploid_chromosome(genotyper, tmp_path):
    """opts->haploid (--haploid-chrs): homozygous calls only, haploid FORMAT of the records."""
    world = bw.synthetic_world(10, config=3, first_locus=300, n_samples=1)
    bams = [abi.BamFile(p) for p in bw.write_world(world, str(tmp_path))]
    motifs = [world["chrom_seq"][s0:s0 + per] for s0, _e, per in world["regions"]]
    out = genotyper.run_regions(bams, "chrS", world["regions"], world["chrom_seq"], 0, motifs=motifs, haploid=True)
    calls = out["calls"]
    n = 0
    for r, st in enumerate(out["status"]):
        if st != 0:
            continue
        l = out["locus_index"][r]
        s0 = calls["locus_sample_begin"][l]
        assert calls["gts"][s0][0] == calls["gts"][s0][1]
        f = out["records"][r].split("\t")
        assert f[8] == "GT:GB:Q:DP:DFLANKINDEL:GLDIFF:ALLREADS:MALLREADS" and "|" not in f[9].split(":")[0]
        n += 1
    assert n >= 8
