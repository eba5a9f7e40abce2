"""N3: the library's BGZF / BAM / BAI reader (csrc/host/bam_reader.cpp, ltr_bam_*) on the reference's shipped reads
(test_data/HG00{2,3,4}_sample_reads.bam, BASELINE.json configs[0] / [1]) against the independent Python decoder of
tools/real_cases.py: every record of every file, and index-driven region queries against brute force.  Host only.
The BAM files are not copied into this repository: the tests run where /root/reference is mounted."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from longtr_b200 import abi  # noqa: E402

DATA = os.path.join(os.environ.get("LONGTR_REFERENCE", "/root/reference"), "test_data")
SAMPLES = ["HG002", "HG003", "HG004"]
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(DATA, "HG002_sample_reads.bam")),
                                reason="reference test data not mounted")


def bam_path(s):
    return os.path.join(DATA, s + "_sample_reads.bam")


@pytest.fixture(scope="module")
def decoded():
    import real_cases
    return {s: list(real_cases.read_bam(bam_path(s))) for s in SAMPLES}


@pytest.mark.parametrize("sample", SAMPLES)
def test_every_record(sample, decoded):
    bam = abi.BamFile(bam_path(sample))
    assert bam.has_index and len(bam.refs) > 0
    got = bam.fetch()
    want = decoded[sample]
    assert len(got) == len(want) > 100
    names = [n for n, _ in bam.refs]
    for g, w in zip(got, want):
        assert g["name"] == w["name"] and g["flag"] == w["flag"] and g["pos"] == w["pos"] and g["mapq"] == w["mapq"]
        assert (names[g["tid"]] if g["tid"] >= 0 else None) == w["chrom"]
        assert g["cigar"] == w["cigar"]
        assert g["seq"].upper() == w["seq"] and g["qual"] == w["qual"]
        assert g["hp"] == (w["hp"] or 0)
        ref_len = sum(n for op, n in g["cigar"] if op in "MDN=X")
        assert g["end"] == g["pos"] + (ref_len or 1)
    bam.close()


@pytest.mark.parametrize("sample", SAMPLES)
def test_region_queries_match_brute_force(sample):
    bam = abi.BamFile(bam_path(sample))
    everything = bam.fetch()
    rng = np.random.default_rng(5)
    tids = sorted({r["tid"] for r in everything if r["tid"] >= 0})
    checked = 0
    for tid in tids:
        recs = [r for r in everything if r["tid"] == tid]
        lo, hi = min(r["pos"] for r in recs), max(r["end"] for r in recs)
        queries = [(lo, hi), (lo - 1000, lo + 1), (hi - 1, hi + 5), (hi, hi + 100), (0, lo)]
        for _ in range(40):
            a = int(rng.integers(lo - 2000, hi + 2000))
            queries.append((a, a + int(rng.choice([1, 50, 500, 5000, 100000]))))
        for r in recs[:: max(1, len(recs) // 25)]:  # windows that start / end exactly on record boundaries
            queries += [(r["pos"], r["pos"] + 1), (r["end"] - 1, r["end"]), (r["end"], r["end"] + 1)]
        for beg, end in queries:
            want = [(r["name"], r["pos"], r["flag"]) for r in recs if r["pos"] < end and r["end"] > max(beg, 0)]
            got = [(r["name"], r["pos"], r["flag"]) for r in bam.fetch(tid, beg, end)]
            assert got == want, (tid, beg, end, len(got), len(want))
            checked += len(want)
    assert checked > 1000
    bam.close()


def test_without_index_and_raw_bytes(tmp_path):
    """A file without .bai is scanned; keep_raw hands out the record bytes (the layout the htslib binding wraps)."""
    src = bam_path("HG002")
    link = tmp_path / "noindex.bam"
    os.symlink(src, link)
    plain = abi.BamFile(str(link))
    indexed = abi.BamFile(src)
    assert not plain.has_index and indexed.has_index
    allr = indexed.fetch()
    tid = allr[len(allr) // 2]["tid"]
    mid = allr[len(allr) // 2]["pos"]
    a = [(r["name"], r["pos"]) for r in plain.fetch(tid, mid - 300, mid + 300)]
    b = indexed.fetch(tid, mid - 300, mid + 300, keep_raw=True)
    assert a == [(r["name"], r["pos"]) for r in b] and len(a) > 0
    for r in b:
        raw = r["raw"]
        assert int.from_bytes(raw[4:8], "little", signed=True) == r["pos"]
        assert raw[32:32 + raw[8] - 1].decode() == r["name"]
    plain.close()
    indexed.close()


def test_index_built_in_memory_equals_the_bai(tmp_path):
    """ltr_bam_build_index on a file without .bai answers region queries like the shipped index does."""
    src = bam_path("HG003")
    link = tmp_path / "noindex.bam"
    os.symlink(src, link)
    built, shipped = abi.BamFile(str(link)), abi.BamFile(src)
    assert not built.has_index
    built.build_index()
    assert built.has_index
    allr = shipped.fetch()
    rng = np.random.default_rng(9)
    n = 0
    for tid in sorted({r["tid"] for r in allr if r["tid"] >= 0}):
        recs = [r for r in allr if r["tid"] == tid]
        lo, hi = min(r["pos"] for r in recs), max(r["end"] for r in recs)
        for _ in range(60):
            a = int(rng.integers(lo - 1000, hi + 1000))
            b = a + int(rng.choice([1, 200, 3000, 60000]))
            want = [(r["name"], r["pos"]) for r in shipped.fetch(tid, a, b)]
            assert [(r["name"], r["pos"]) for r in built.fetch(tid, a, b)] == want
            n += len(want)
    assert n > 500
    built.close()
    shipped.close()


def test_open_errors(tmp_path):
    with pytest.raises(RuntimeError):
        abi.BamFile(str(tmp_path / "missing.bam"))
    junk = tmp_path / "junk.bam"
    junk.write_bytes(b"not a bam file at all" * 10)
    with pytest.raises(RuntimeError):
        abi.BamFile(str(junk))
    with pytest.raises(RuntimeError):
        abi.BamFile(bam_path("HG002"), index_path=str(junk))


def test_sequence_and_quality_decoding_edge_cases(tmp_path):
    """Every 4-bit base code, odd and even lengths around the 16-byte vector width, missing qualities (0xff -> '!') and
    qualities above 93 (-> '~'), as htslib prints them (SAM specification 4.2.3 / 4.2.4)."""
    import bam_writer as bw
    import struct
    codes = "=ACMGRSVTWYHKDBN"
    recs, want = [], []
    for n in (1, 2, 15, 16, 17, 31, 32, 33, 47, 64, 101):
        seq = "".join(codes[(3 * i + n) % 16] for i in range(n))
        raw_q = bytes((i * 37 + n) % 256 if i % 5 else 0xff for i in range(n))
        rec = bytearray(bw.encode_record(0, 100 + n, "r%d" % n, 0, 60, [("M", n)], seq, "!" * n))
        # overwrite the quality bytes (they sit right behind the packed sequence)
        l_name = rec[4 + 8]
        q_at = 4 + 32 + l_name + 4 + (n + 1) // 2
        rec[q_at:q_at + n] = raw_q
        recs.append(bytes(rec))
        want.append((seq, "".join("!" if q == 0xff else chr(126 if q > 93 else q + 33) for q in raw_q)))
    path = str(tmp_path / "edge.bam")
    bw.write_bam(path, [("chrE", 10000)], recs)
    b = abi.BamFile(path)
    got = b.fetch()
    assert [(r["seq"], r["qual"]) for r in got] == want
    b.close()


def test_read_name_without_terminator(tmp_path):
    """A damaged record whose read name lacks its NUL (l_read_name counts it, SAM specification 4.2): the reader writes the
    terminator itself instead of letting strlen run into the next record's name (found by tools/bam_fuzz.cpp under ASan)."""
    import bam_writer as bw
    recs = []
    for k, name in enumerate(("first", "second", "third")):
        rec = bytearray(bw.encode_record(0, 100 + k, name, 0, 60, [("M", 8)], "ACGTACGT", "I" * 8))
        if k < 2:
            rec[4 + 32 + len(name)] = ord("X")   # the terminator of the name
        recs.append(bytes(rec))
    path = str(tmp_path / "noterm.bam")
    bw.write_bam(path, [("chrE", 10000)], recs)
    b = abi.BamFile(path)
    got = b.fetch()
    assert [r["name"] for r in got] == ["first", "second", "third"] and [r["seq"] for r in got] == ["ACGTACGT"] * 3
    b.close()
