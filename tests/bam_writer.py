"""Test utility: writes coordinate-sorted BGZF / BAM files (zlib only) and builds a synthetic chromosome + BAM files from
the raw loci of the workload generator (longtr_b200.workloads.generate_loci), so that the BAM -> calls path
(ltr_bam_* / ltr_region_collect / ltr_candidate_alleles / ltr_regions_run) can be exercised without the reference's data.
Not product code."""
import struct
import zlib

import numpy as np

_SEQ_CODE = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
_EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def _bgzf_block(data):
    comp = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = comp.compress(data) + comp.flush()
    bsize = 12 + 6 + len(body) + 8
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize - 1) + body +
            struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data)))


def _reg2bin(beg, end):
    end -= 1
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return base + (beg >> shift)
    return 0


def encode_record(tid, pos, name, flag, mapq, cigar, seq, qual, hp=None):
    """cigar: [(op char, len)]; qual: Phred+33 string."""
    ref_len = sum(n for op, n in cigar if op in "MDN=X")
    nm = name.encode() + b"\0"
    cig = b"".join(struct.pack("<I", (n << 4) | "MIDNSHP=X".index(op)) for op, n in cigar)
    codes = [_SEQ_CODE[c] for c in seq] + [0]
    packed = bytes((codes[i] << 4) | codes[i + 1] for i in range(0, len(seq), 2))
    q = bytes(ord(c) - 33 for c in qual)
    aux = b"" if hp is None else b"HPC" + bytes([hp])
    core = struct.pack("<iiBBHHHiiii", tid, pos, len(nm), mapq, _reg2bin(pos, pos + max(1, ref_len)), len(cigar), flag,
                       len(seq), -1, -1, 0)
    body = core + nm + cig + packed + q + aux
    return struct.pack("<i", len(body)) + body


def write_bam(path, refs, records, block=60000):
    """refs: [(name, length)]; records: encoded records, already sorted by (tid, pos)."""
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % r for r in refs)
    head = b"BAM\1" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(refs))
    for name, ln in refs:
        head += struct.pack("<i", len(name) + 1) + name.encode() + b"\0" + struct.pack("<i", ln)
    with open(path, "wb") as f:
        f.write(_bgzf_block(head))
        buf = b""
        for r in records:
            if len(buf) + len(r) > block and buf:
                f.write(_bgzf_block(buf))
                buf = b""
            buf += r
        if buf:
            f.write(_bgzf_block(buf))
        f.write(_EOF_BLOCK)


def _cigar_list(ops):
    return [("MIDNSHP=X"[int(v) & 15], int(v) >> 4) for v in ops]


def synthetic_world(n_loci, config=3, first_locus=0, n_samples=1, spacing=4000):
    """The raw loci of the generator laid out on one chromosome 'chrS'.  Returns dict(chrom_seq, regions [(start, stop,
    period)], records [per sample: encoded BAM records], refs).  Locus l sits at offset l * spacing; a read's sample is its
    index modulo n_samples; reads with phasing terms carry the HP tag; qualities are constant (Q40)."""
    from longtr_b200 import workloads
    w = workloads.generate_loci(config, n_loci, first_locus=first_locus)
    B = w.struct
    arr = np.ctypeslib.as_array
    lab = arr(B.locus_allele_begin, (n_loci + 1,)).copy()
    aoff = arr(B.allele_off, (int(lab[-1]) + 1,)).copy()
    abytes = arr(B.allele_bytes, (int(aoff[-1]),)).tobytes()
    lrb = arr(B.locus_read_begin, (n_loci + 1,)).copy()
    nr = int(lrb[-1])
    roff = arr(B.read_off, (nr + 1,)).copy()
    rbytes = arr(B.read_bytes, (int(roff[-1]),)).tobytes()
    coff = arr(B.cigar_off, (nr + 1,)).copy()
    cops = arr(B.cigar_ops, (int(coff[-1]),)).copy()
    rstart = arr(B.read_start, (nr,)).copy()
    p1 = arr(B.log_p1, (nr,)).copy()
    p2 = arr(B.log_p2, (nr,)).copy()
    rep0 = arr(B.repeat_start, (n_loci,)).copy()
    rep1 = arr(B.repeat_end, (n_loci,)).copy()
    lfo = arr(B.lflank_off, (n_loci + 1,)).copy()
    lfb = arr(B.lflank_bytes, (int(lfo[-1]),)).tobytes()
    rfo = arr(B.rflank_off, (n_loci + 1,)).copy()
    rfb = arr(B.rflank_bytes, (int(rfo[-1]),)).tobytes()
    chrom = bytearray(b"N" * (n_loci * spacing + spacing))
    regions, recs = [], [[] for _ in range(n_samples)]
    for l in range(n_loci):
        base = l * spacing
        ref_allele = abytes[aoff[lab[l]]:aoff[lab[l] + 1]]
        lf, rf = lfb[lfo[l]:lfo[l + 1]], rfb[rfo[l]:rfo[l + 1]]
        # known reference: flank blocks and the reference allele; the +-200 bp context from the reads' '=' runs
        for r in range(lrb[l], lrb[l + 1]):
            rp, sp = base + int(rstart[r]), int(roff[r])
            for op, n in _cigar_list(cops[coff[r]:coff[r + 1]]):
                if op == "=":
                    chrom[rp:rp + n] = rbytes[sp:sp + n]
                if op in "=XM":
                    rp += n; sp += n
                elif op == "D":
                    rp += n
                elif op == "I":
                    sp += n
        s0 = base + int(rep0[l])
        chrom[s0 - len(lf):s0] = lf
        chrom[s0:s0 + len(ref_allele)] = ref_allele
        chrom[base + int(rep1[l]):base + int(rep1[l]) + len(rf)] = rf
        core = ref_allele[5:-5]  # the allele block carries 5 bp of padding on either side
        period = next(p for p in range(1, len(core) + 1) if core[p:] == core[:-p] or p == len(core))
        regions.append((s0 + 5, base + int(rep1[l]) - 5, period))
        for k, r in enumerate(range(lrb[l], lrb[l + 1])):
            seq = rbytes[roff[r]:roff[r + 1]].decode()
            hp = 1 if (p1[r] > -1 and p2[r] < -1) else (2 if (p2[r] > -1 and p1[r] < -1) else None)
            recs[k % n_samples].append((base + int(rstart[r]),
                                        encode_record(0, base + int(rstart[r]), "L%dR%d" % (l, k), 16 if k % 3 == 0 else 0, 60,
                                                      _cigar_list(cops[coff[r]:coff[r + 1]]), seq, "I" * len(seq), hp)))
    w.close()
    out = [[rec for _, rec in sorted(rs, key=lambda t: t[0])] for rs in recs]
    return dict(chrom_seq=bytes(chrom).decode(), regions=regions, records=out, refs=[("chrS", len(chrom))])


def write_world(world, directory, prefix="synth"):
    paths = []
    for s, recs in enumerate(world["records"]):
        p = "%s/%s_%d.bam" % (directory, prefix, s)
        write_bam(p, world["refs"], recs)
        paths.append(p)
    return paths
