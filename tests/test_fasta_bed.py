"""N3: ltr_fasta_* (indexed FASTA access in place of htslib's faidx behind the reference's FastaReader, src/fasta_reader.{h,cpp})
and ltr_bed_read (readRegions + orderRegions, src/region.cpp:26-75).  Host only.  The FASTA reader is held to Python's own
parsing of the files written here (with a samtools-style .fai, without one, wrapped at several widths, CRLF, a directory of
files); the region reader to the reference's rules restated line by line."""
import os
import random

import pytest

from longtr_b200 import abi
from oracle import pyregion as pr


def _write_fasta(path, seqs, width, fai=True, newline="\n"):
    entries = []
    with open(path, "wb") as f:
        for name, s in seqs:
            f.write((">%s some description%s" % (name, newline)).encode())
            off = f.tell()
            for k in range(0, len(s), width):
                f.write((s[k:k + width] + newline).encode())
            entries.append((name, len(s), off, width, width + len(newline)))
    if fai:
        with open(path + ".fai", "w") as f:
            for e in entries:
                f.write("%s\t%d\t%d\t%d\t%d\n" % e)


def _seqs(rng, n):
    return [("chr%s" % (k + 1), "".join(rng.choice("ACGTNacgt") for _ in range(rng.randint(1, 5000)))) for k in range(n)]


@pytest.mark.parametrize("fai", [True, False])
@pytest.mark.parametrize("width,newline", [(60, "\n"), (7, "\n"), (10000, "\n"), (50, "\r\n")])
def test_fasta_fetch(tmp_path, fai, width, newline):
    rng = random.Random(width * 2 + fai)
    seqs = _seqs(rng, 5)
    path = str(tmp_path / "ref.fa")
    _write_fasta(path, seqs, width, fai=fai, newline=newline)
    fa = abi.FastaFile(path)
    assert fa.names == [n for n, _ in seqs]
    for name, s in seqs:
        assert fa.length(name) == len(s)
        assert fa.fetch(name) == s
        for _ in range(20):
            a = rng.randint(0, len(s))
            b = rng.randint(a, len(s))
            assert fa.fetch(name, a, b) == s[a:b]
    assert fa.length("chrNope") == -1
    with pytest.raises(RuntimeError):
        fa.fetch("chr1", 0, len(seqs[0][1]) + 1)
    fa.close()


def test_fasta_directory_and_errors(tmp_path):
    rng = random.Random(3)
    d = tmp_path / "genome"
    d.mkdir()
    a, b = _seqs(rng, 2), [("chrX", "ACGT" * 100), ("chrY", "TTGA" * 33)]
    _write_fasta(str(d / "a.fa"), a, 60)
    _write_fasta(str(d / "b.fa"), b, 80, fai=False)
    (d / "notes.txt").write_text("not a fasta")
    fa = abi.FastaFile(str(d))
    assert sorted(fa.names) == sorted(n for n, _ in a + b)
    assert fa.fetch("chrY", 3, 50) == ("TTGA" * 33)[3:50]
    fa.close()
    _write_fasta(str(d / "c.fa"), [("chrX", "AAAA")], 60)          # the same name in two files
    with pytest.raises(RuntimeError):
        abi.FastaFile(str(d))
    with pytest.raises(RuntimeError):
        abi.FastaFile(str(tmp_path / "missing.fa"))
    empty = tmp_path / "empty_dir"
    empty.mkdir()
    with pytest.raises(RuntimeError):
        abi.FastaFile(str(empty))
    import gzip
    with gzip.open(str(tmp_path / "z.fa"), "wb") as f:
        f.write(b">chr1\nACGT\n")
    with pytest.raises(RuntimeError):
        abi.FastaFile(str(tmp_path / "z.fa"))
    ragged = tmp_path / "ragged.fa"
    ragged.write_text(">chr1\nACGT\nAC\nACGT\n")                   # a short line in the middle cannot be indexed
    with pytest.raises(RuntimeError):
        abi.FastaFile(str(ragged))


def test_bed_read(tmp_path):
    p = tmp_path / "regions.bed"
    p.write_text("chr2\t500\t530\tAC\tlocusB\n"
                 "chr1\t1000\t1045\tCAG\tlocus1\n"
                 "chr1\t200\t212\tA\n"
                 "chr1\t1000\t1030\tAAAG,AAAC\tlocus2\n"
                 "chr10\t7\t99\tAT,GGC\tmixed\n")
    b = abi.bed_read(str(p))
    assert b["chroms"] == ["chr1", "chr10", "chr2"]                 # sorted by chromosome name, then start, then stop
    assert b["regions"] == [(0, 199, 212, 1, "", "A"), (0, 999, 1030, 4, "locus2", "AAAG,AAAC"),
                            (0, 999, 1045, 3, "locus1", "CAG"), (1, 6, 99, -1, "mixed", "AT,GGC"),
                            (2, 499, 530, 2, "locusB", "AC")]
    only = abi.bed_read(str(p), chrom_limit="chr2")
    assert only["chroms"] == ["chr2"] and len(only["regions"]) == 1
    assert len(abi.bed_read(str(p), max_regions=2)["regions"]) == 2
    with pytest.raises(RuntimeError):
        abi.bed_read(str(p), chrom_limit="chr7")                    # no region on the requested chromosome
    for bad in ("chr1\t0\t10\tA\n", "chr1\t10\t10\tA\n", "chr1\t10\t20\tA1\n", "chr1\t10\n", "chr1\tx\t20\tA\n"):
        q = tmp_path / "bad.bed"
        q.write_text(bad)
        with pytest.raises(RuntimeError):
            abi.bed_read(str(q))
    with pytest.raises(RuntimeError):
        abi.bed_read(str(tmp_path / "missing.bed"))


@pytest.mark.skipif(not pr.ref_hapgen_available(), reason="oracle/_ref/libltr_ref_hapgen.so not built")
def test_bed_read_matches_the_reference(tmp_path):
    """Against readRegions + orderRegions compiled in place (oracle/hapgen_driver.cpp) on seeded region files."""
    rng = random.Random(11)
    for it in range(8):
        lines = []
        for _ in range(rng.randint(1, 60)):
            start = rng.randint(1, 5000)
            motif = ",".join("".join(rng.choice("ACGT") for _ in range(rng.randint(1, 6))) for _ in range(rng.choice([1, 1, 1, 2, 3])))
            line = "%s\t%d\t%d\t%s" % (rng.choice(["chr1", "chr2", "chr10", "chrX", "1"]), start, start + rng.randint(1, 400), motif)
            if rng.random() < 0.7:
                line += "\tL%d" % rng.randint(0, 999)
            lines.append(line)
        p = tmp_path / ("r%d.bed" % it)
        p.write_text("\n".join(lines) + "\n")
        for kw in (dict(), dict(chrom_limit="chr2"), dict(max_regions=5)):
            if kw.get("chrom_limit") and not any(l.startswith("chr2\t") for l in lines[:kw.get("max_regions", len(lines))]):
                with pytest.raises(RuntimeError):                        # the reference exits here
                    abi.bed_read(str(p), **kw)
                continue
            want = pr.ref_read_regions(str(p), kw.get("max_regions", 1000000000), kw.get("chrom_limit"))
            got = abi.bed_read(str(p), **kw)
            key = lambda r: (r[0], r[1], r[2])
            flat = [(got["chroms"][c], s, e, per, name, motif) for c, s, e, per, name, motif in got["regions"]]
            assert [key(r) for r in flat] == [key(r) for r in want]      # ties are in unspecified order in std::sort
            assert sorted(flat) == sorted(want)


@pytest.mark.skipif(not pr.ref_fasta_available(), reason="oracle/_ref/libltr_ref_fasta.so not built")
def test_reference_fasta_reader_runs_on_our_reader_and_header_matches(tmp_path):
    """integration/faidx_compat.cpp serves the faidx names LongTR binds: its own FastaReader (compiled in place) returns the
    sequences of the file, and Genotyper::get_vcf_header -- contig lines through that reader -- equals ltr_vcf_header."""
    rng = random.Random(21)
    seqs = _seqs(rng, 4)
    path = str(tmp_path / "ref.fa")
    _write_fasta(path, seqs, 70)
    for name, s in seqs:
        got, n = pr.ref_fasta_sequence(path, name)
        assert n == len(s) and got == s
    assert pr.ref_fasta_sequence(path, "chrNope")[1] == -1
    fa = abi.FastaFile(path)
    for samples in (["HG002"], ["HG002", "HG003", "HG004"], []):
        cmd = "LongTR --bams a.bam --fasta ref.fa --regions r.bed"
        assert abi.vcf_header(fa, path, cmd, samples) == pr.ref_vcf_header(path, cmd, samples)
    for mask in (0, 3, 7, 60, 63, 37):   # --hide-allreads / --hide-mallreads / --output-gls / -pls / -phased-gls / -filters
        assert abi.vcf_header(fa, path, cmd, ["HG002"], switches=mask) == pr.ref_vcf_header(path, cmd, ["HG002"], switches=mask)
    fa.close()
