#!/usr/bin/env python
"""Builds tests/golden/real_cases.json from the reference's shipped trio reads (BASELINE.json configs[0] and [1]).

    python tools/real_cases.py          (in the build container: needs /root/reference/test_data and oracle/_ref)

LongTR's own IO layer (htslib) and the hg38 FASTA do not exist in this image, so the reads of
test_data/HG00{2,3,4}_sample_reads.bam are decoded here with a small BGZF/BAM reader (zlib only) and turned into the
IO-less per-locus inputs of SeqStutterGenotyper (the boundary the drop-in test uses):
  * pseudo reference: every base that some read reports with an '=' operation is known; the rest stays 'N'
    (both the reference run and the GPU run see the same sequence, so parity is meaningful);
  * per BED region (HipSTR 7-column file, converted as SURVEY Q6 describes): primary reads that span the region,
    cut to region +-200 bp (GenotyperBamProcessor::left_align_reads, src/genotyper_bam_processor.cpp:55-62),
    '=XID' CIGAR rebuilt against the pseudo reference (:75-128), soft-clipped reads dropped (:131-134);
  * phasing terms from the HP tag as --phased-bam does (src/snp_bam_processor.cpp:141-237, snp_bam_processor.h:16-18).
This is test tooling: an approximation of LongTR's read filters is good enough because the reference's genotyper
(oracle/_ref/ltr_ref_full) and the GPU drop-in build (ltr_ref_gpu) are fed the very same alignments.
Regions where the reference's candidate-allele code would need spoa (stubbed out) or genotype() fails are skipped.
"""
import gzip
import json
import os
import struct
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = os.environ.get("LONGTR_REFERENCE", "/root/reference")
DATA = os.path.join(REF, "test_data")
SAMPLES = ["HG002", "HG003", "HG004"]
FLANK = 200  # bam_io.h:28 FLANK_SIZE


def bgzf_decompress(path):
    raw = open(path, "rb").read()
    out, pos = [], 0
    while pos < len(raw):
        assert raw[pos:pos + 4] == b"\x1f\x8b\x08\x04"
        xlen = struct.unpack_from("<H", raw, pos + 10)[0]
        extra = raw[pos + 12:pos + 12 + xlen]
        bsize, p = None, 0
        while p < len(extra):
            si1, si2, slen = extra[p], extra[p + 1], struct.unpack_from("<H", extra, p + 2)[0]
            if si1 == 66 and si2 == 67:
                bsize = struct.unpack_from("<H", extra, p + 4)[0]
            p += 4 + slen
        cdata = raw[pos + 12 + xlen:pos + bsize + 1 - 8]
        out.append(zlib.decompress(cdata, -15))
        pos += bsize + 1
    return b"".join(out)


def read_bam(path):
    """Yields dicts (name, flag, tid, pos, mapq, cigar [(op, len)], seq, qual, hp)."""
    d = bgzf_decompress(path)
    assert d[:4] == b"BAM\x01"
    l_text = struct.unpack_from("<i", d, 4)[0]
    p = 8 + l_text
    n_ref = struct.unpack_from("<i", d, p)[0]
    p += 4
    refs = []
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", d, p)[0]
        refs.append(d[p + 4:p + 4 + l_name - 1].decode())
        p += 4 + l_name + 4
    while p < len(d):
        bs = struct.unpack_from("<i", d, p)[0]
        rec = d[p + 4:p + 4 + bs]
        p += 4 + bs
        tid, pos, l_read_name, mapq, _bin, n_cig, flag, l_seq = struct.unpack_from("<iiBBHHHi", rec, 0)
        q = 32
        name = rec[q:q + l_read_name - 1].decode()
        q += l_read_name
        cigar = []
        for k in range(n_cig):
            v = struct.unpack_from("<I", rec, q + 4 * k)[0]
            cigar.append(("MIDNSHP=X"[v & 15], v >> 4))
        q += 4 * n_cig
        sb = rec[q:q + (l_seq + 1) // 2]
        q += (l_seq + 1) // 2
        seq = "".join("=ACMGRSVTWYHKDBN"[(sb[i >> 1] >> (4 if i % 2 == 0 else 0)) & 15] for i in range(l_seq))
        qual = "".join(chr(min(c, 93) + 33) for c in rec[q:q + l_seq])
        q += l_seq
        hp = None
        while q < len(rec):  # aux fields
            tag, typ = rec[q:q + 2].decode(), chr(rec[q + 2])
            q += 3
            if typ in "cC":
                val = struct.unpack_from("<b" if typ == "c" else "<B", rec, q)[0]; q += 1
            elif typ in "sS":
                val = struct.unpack_from("<h" if typ == "s" else "<H", rec, q)[0]; q += 2
            elif typ in "iI":
                val = struct.unpack_from("<i" if typ == "i" else "<I", rec, q)[0]; q += 4
            elif typ == "f":
                val = None; q += 4
            elif typ == "A":
                val = None; q += 1
            elif typ in "ZH":
                e = rec.index(b"\x00", q); val = None; q = e + 1
            elif typ == "B":
                sub, cnt = chr(rec[q]), struct.unpack_from("<i", rec, q + 1)[0]
                q += 5 + cnt * {"c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}[sub]; val = None
            else:
                raise ValueError("aux type " + typ)
            if tag == "HP":
                hp = val
        yield dict(name=name, flag=flag, chrom=refs[tid] if tid >= 0 else None, pos=pos, mapq=mapq, cigar=cigar,
                   seq=seq.upper(), qual=qual, hp=hp)


def ref_end(r):
    return r["pos"] + sum(n for op, n in r["cigar"] if op in "MDN=X")


def build_pseudo_reference(reads, lo, hi):
    ref = bytearray(b"N" * (hi - lo))
    for r in reads:
        rp, sp = r["pos"], 0
        for op, n in r["cigar"]:
            if op == "=":
                a, b = max(rp, lo), min(rp + n, hi)
                if a < b:
                    ref[a - lo:b - lo] = r["seq"][sp + (a - rp):sp + (b - rp)].encode()
                rp += n; sp += n
            elif op in "MX":
                rp += n; sp += n
            elif op in "DN":
                rp += n
            elif op in "IS":
                sp += n
    return ref.decode()


def trim_read(r, ref, ref_lo, win_lo, win_hi):
    """Cut the read to [win_lo, win_hi) on the reference and rebuild an '=XID' CIGAR against ref.
    Returns dict(start, stop, seq, qual, aln, cigar) or None (soft clip inside the window / nothing left)."""
    rp, sp = r["pos"], 0
    ops, seq, qual, aln = [], [], [], []
    started = False
    for op, n in r["cigar"]:
        for _ in range(n):
            if op in "M=X":
                if win_lo <= rp < win_hi:
                    started = True
                    c = r["seq"][sp]
                    rc = ref[rp - ref_lo]
                    ops.append("=" if c == rc else "X")
                    seq.append(c); qual.append(r["qual"][sp]); aln.append(c)
                rp += 1; sp += 1
            elif op == "I":
                if win_lo < rp <= win_hi - 1 and started:
                    c = r["seq"][sp]
                    ops.append("I"); seq.append(c); qual.append(r["qual"][sp]); aln.append(c)
                sp += 1
            elif op in "DN":
                if win_lo <= rp < win_hi and started:
                    ops.append("D"); aln.append("-")
                rp += 1
            elif op == "S":
                if win_lo <= rp < win_hi:
                    return None
                sp += 1
            elif op == "H":
                pass
    # strip leading / trailing non-aligned operations so that start/stop are well defined
    while ops and ops[0] in "ID":
        if ops[0] == "I":
            seq.pop(0); qual.pop(0); aln.pop(0)
        else:
            aln.pop(0)
        ops.pop(0)
    while ops and ops[-1] in "ID":
        if ops[-1] == "I":
            seq.pop(); qual.pop(); aln.pop()
        else:
            aln.pop()
        ops.pop()
    if not ops:
        return None
    first_ref = max(win_lo, r["pos"])
    # first aligned base on the reference
    rp, k = r["pos"], 0
    start = None
    for op, n in r["cigar"]:
        if op in "M=X":
            if rp + n > win_lo:
                start = max(rp, win_lo); break
            rp += n
        elif op in "DN":
            rp += n
    n_ref = sum(1 for o in ops if o in "=XD")
    cig, prev, cnt = [], None, 0
    for o in ops:
        if o == prev:
            cnt += 1
        else:
            if prev:
                cig.append("%d%s" % (cnt, prev))
            prev, cnt = o, 1
    cig.append("%d%s" % (cnt, prev))
    return dict(start=start, stop=start + n_ref - 1, seq="".join(seq), qual="".join(qual), aln="".join(aln),
                cigar="".join(cig))


def regions():
    out = []
    for line in open(os.path.join(DATA, "test_regions_hg38.bed")):
        f = line.split()
        if len(f) < 7:
            continue
        motif = f[6].replace("/", ",")
        out.append(dict(chrom=f[0], start=int(f[1]) - 1, stop=int(f[2]), motif=motif, name=f[5]))
    return out


def main():
    from oracle import pyoracle as po
    reads = {s: [r for r in read_bam(os.path.join(DATA, s + "_sample_reads.bam"))
                 if r["chrom"] == "chr1" and not (r["flag"] & 0x904) and r["mapq"] >= 1] for s in SAMPLES}
    allr = [r for s in SAMPLES for r in reads[s]]
    lo = min(r["pos"] for r in allr)
    hi = max(ref_end(r) for r in allr)
    ref = build_pseudo_reference(allr, lo, hi)
    print("reads", {s: len(v) for s, v in reads.items()}, "window", lo, hi, "unknown bases", ref.count("N"))
    cases = []
    for reg in regions():
        if reg["stop"] - reg["start"] > 1000 or "," in reg["motif"]:
            continue
        win_lo, win_hi = max(reg["start"] - FLANK, lo), min(reg["stop"] + FLANK, hi)
        # chromosome slice handed to the genotyper: coordinates shifted so that the window starts at `pad`
        pad = 300
        c_lo = max(lo, win_lo - pad)
        chrom_seq = ref[c_lo - lo:min(hi, win_hi + pad) - lo]
        shift = c_lo
        for sample_set, tag in ((["HG002"], "hg002"), (SAMPLES, "trio")):
            rs, n1, n2 = [], [], []
            for si, s in enumerate(sample_set):
                c1 = c2 = 0
                for r in reads[s]:
                    if r["pos"] > reg["start"] or ref_end(r) < reg["stop"]:
                        continue
                    t = trim_read(r, ref, lo, win_lo, win_hi)
                    if t is None or len(t["seq"]) < 50 or "N" in ref[t["start"] - lo:t["stop"] + 1 - lo]:
                        continue
                    hp = r["hp"]
                    rs.append(dict(start=t["start"] - shift, stop=t["stop"] - shift, rev=bool(r["flag"] & 16), sample=si,
                                   name=r["name"].replace(" ", "_"), seq=t["seq"], qual=t["qual"], aln=t["aln"],
                                   cigar=t["cigar"], log_p1=-1e-6 if hp == 1 else (-1000.0 if hp == 2 else 0.0),
                                   log_p2=-1000.0 if hp == 1 else (-1e-6 if hp == 2 else 0.0)))
                    c1 += hp == 1
                    c2 += hp == 2
                n1.append(int(c1)); n2.append(int(c2))
            if len(rs) < 5 or any(sum(1 for x in rs if x["sample"] == si) == 0 for si in range(len(sample_set))):
                continue
            cases.append(dict(name="%s_%s" % (reg["name"], tag), chrom_name="chr1", chrom_seq=chrom_seq,
                              region_start=reg["start"] - shift, region_stop=reg["stop"] - shift, motif=reg["motif"],
                              region_name=reg["name"], samples=list(sample_set), n_p1s=n1, n_p2s=n2, reads=rs,
                              stutter_motif="A", stutter_period=len(reg["motif"])))
    print(len(cases), "candidate cases")
    good = []
    for c in cases:
        try:
            rec = po.full_locus_records([c], "full")[0]
        except Exception as e:  # spoa stub reached / reference abort
            print("skip", c["name"], str(e).splitlines()[-1][:80])
            continue
        if not rec:
            print("skip", c["name"], "genotype() returned false")
            continue
        c["record"] = rec
        good.append(c)
    path = os.path.join(ROOT, "tests", "golden", "real_cases.json.gz")
    with gzip.open(path, "wt") as f:
        json.dump(dict(generator="tools/real_cases.py", source="test_data/HG00{2,3,4}_sample_reads.bam through oracle/_ref/ltr_ref_full",
                       cases=good), f, separators=(",", ":"))
    print(path, len(good), "cases", os.path.getsize(path), "bytes")
    for c in good[:6]:
        print(c["name"], len(c["reads"]), c["record"].split("\t")[3][:30], c["record"].split("\t")[9:])


if __name__ == "__main__":
    main()
