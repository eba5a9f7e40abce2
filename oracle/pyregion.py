"""TEST INFRASTRUCTURE ONLY -- plain-Python restatement of how LongTR's region loop prepares the reads of one region
(single-end reads), used to check ltr_region_collect.  Follows the reference line by line:

  filter_and_order   BamProcessor::read_and_filter_reads       src/bam_processor.cpp:188-487 (+ process_regions :584-596)
  phasing_terms      SNPBamProcessor::process_phased_reads     src/snp_bam_processor.cpp:141-232
  trim_alignment     BamAlignment::TrimAlignment               src/bam_io.cpp:267-372
  left_align         GenotyperBamProcessor::left_align_reads   src/genotyper_bam_processor.cpp:38-168

trim_alignment is pinned by the reference's own TrimAlignment (oracle/_ref/libltr_ref_io.so: bam_io.cpp compiled in place
on top of integration/hts_compat.cpp), see tests/test_region_loader.py.  Reads are dicts as longtr_b200.abi.BamFile.fetch
returns them (name, flag, pos, end, mapq, cigar [(op, len)], seq, qual, hp, raw)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_IO_SO = os.path.join(_HERE, "_ref", "libltr_ref_io.so")
FROM_HAP_LL, OTHER_HAP_LL = -0.000001, -1000.0  # snp_bam_processor.h:16-18


def has_tag(raw, tag):
    """Aux field present in the raw BAM record?"""
    l_name, n_cig = raw[8], int.from_bytes(raw[12:14], "little")
    l_seq = int.from_bytes(raw[16:20], "little")
    q = 32 + l_name + 4 * n_cig + (l_seq + 1) // 2 + l_seq
    size = {"A": 1, "c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4, "d": 8}
    while q + 3 <= len(raw):
        t, ty = raw[q:q + 2].decode(), chr(raw[q + 2])
        q += 3
        if t == tag:
            return True
        if ty in size:
            q += size[ty]
        elif ty in "ZH":
            q = raw.index(b"\0", q) + 1
        elif ty == "B":
            q += 5 + int.from_bytes(raw[q + 1:q + 5], "little") * {"c": 1, "C": 1, "s": 2, "S": 2}.get(chr(raw[q]), 4)
        else:
            return False
    return False


def filter_and_order(reads_by_file, start, stop, max_mate_dist=1000, min_mean_qual=30.0, min_mapq=20.0,
                     require_spanning=1):
    """-> (samples: [file index], reads_by_sample: [[read]], counters)."""
    cnt = dict(n_overlapping=0, n_hard_clipped=0, n_has_n=0, n_low_qual=0, n_low_mapq=0, n_not_spanning=0, n_not_unique=0)
    potential_strs = {}
    for f, reads in enumerate(reads_by_file):
        potential_mates = set()
        label = "%d_" % (f + 1)
        q1 = stop + max_mate_dist
        for r in reads:
            if r["pos"] > q1 + 1:  # bam_io.cpp:178
                break
            if r["pos"] > stop or r["end"] < start:  # :208-216
                continue
            if (r["flag"] & 4) or r["pos"] == 0 or not r["cigar"] or not r["seq"]:  # :224
                continue
            if not (r["pos"] < stop and r["end"] >= start):  # :255
                continue
            cnt["n_overlapping"] += 1
            if r["cigar"][0][0] == "H" or r["cigar"][-1][0] == "H":  # :231-237
                cnt["n_hard_clipped"] += 1
                continue
            ok = False
            if "N" in r["seq"]:
                cnt["n_has_n"] += 1
            elif sum(ord(c) - 33 for c in r["qual"]) / len(r["qual"]) < min_mean_qual:
                cnt["n_low_qual"] += 1
            elif r["mapq"] < min_mapq:
                cnt["n_low_mapq"] += 1
            elif require_spanning == 1 and not (r["pos"] <= start and r["end"] >= stop):
                cnt["n_not_spanning"] += 1
            else:
                ok = True
            name = r["name"]
            if len(name) > 2 and name[-2] == "/":
                name = name[:-2]
            key = label + name
            if not ok:
                potential_mates.add(key)
                continue
            potential_mates.discard(key)
            if key not in potential_strs:  # std::map::insert keeps the first
                potential_strs[key] = dict(r, file=f)
    unpaired = []
    for key in sorted(potential_strs):  # std::map order: byte-wise
        r = potential_strs[key]
        if has_tag(r["raw"], "XA"):
            cnt["n_not_unique"] += 1
        else:
            unpaired.append(r)
    cnt["n_passed"] = len(unpaired)
    samples, by_sample = [], []
    for r in reversed(unpaired):  # :453-483 pop_back
        if r["file"] not in samples:
            samples.append(r["file"])
            by_sample.append([])
        by_sample[samples.index(r["file"])].append(r)
    return samples, by_sample, cnt


def phasing_terms(by_sample, phased_bam=True):
    total = h1 = h2 = 0
    not_enough = False
    out = []
    for reads in by_sample:
        haps = [(r["hp"] if has_tag(r["raw"], "HP") else -1) for r in reads]
        for h in haps:
            total += 1
            h1 += h == 1
            h2 += h == 2
        if (total - (h1 + h2)) / total > 0.2 or h2 <= 1 or h1 <= 1:
            not_enough = True
        terms = []
        for h in haps:
            if phased_bam and h != -1 and not not_enough:
                terms.append((FROM_HAP_LL if h == 1 else OTHER_HAP_LL, FROM_HAP_LL if h == 2 else OTHER_HAP_LL))
            else:
                terms.append((0.0, 0.0))
        out.append(terms)
    return out


def trim_alignment(r, min_read_start, max_read_stop, flank=200):
    """-> dict(pos, end, cigar, seq, qual, deleted) (bam_io.cpp:267-372)."""
    cig = [[op, n] for op, n in r["cigar"]]
    ltrim, start_pos = 0, r["pos"]
    while start_pos < min_read_start and cig:
        op = cig[0][0]
        if op in "M=X":
            ltrim += 1
            start_pos += 1
        elif op == "D":
            start_pos += 1
        elif op in "IS":
            ltrim += 1
        if cig[0][1] == 1:
            cig.pop(0)
        else:
            cig[0][1] -= 1
    ptr, rep_start, rep_end, dele = start_pos, min_read_start + flank, max_read_stop - flank, 0
    tmp = [[op, n] for op, n in cig]
    while min_read_start <= ptr < rep_end and tmp:
        op = tmp[0][0]
        if op in "M=X":
            ptr += 1
        elif op == "D":
            if ptr >= rep_start:
                dele += 1
            ptr += 1
        if tmp[0][1] == 1:
            tmp.pop(0)
        else:
            tmp[0][1] -= 1
    deleted = dele >= rep_end - rep_start
    rtrim, end_pos = 0, r["end"]
    while end_pos > max_read_stop and cig:
        op = cig[-1][0]
        if op in "M=X":
            rtrim += 1
            end_pos -= 1
        elif op == "D":
            end_pos -= 1
        elif op in "IS":
            rtrim += 1
        if cig[-1][1] == 1:
            cig.pop()
        else:
            cig[-1][1] -= 1
    n = len(r["seq"])
    return dict(pos=start_pos, end=end_pos, cigar=[(op, k) for op, k in cig], seq=r["seq"][ltrim:n - rtrim],
                qual=r["qual"][ltrim:n - rtrim], deleted=deleted)


def left_align(samples, by_sample, terms, start, stop, ref, ref_start, flank=200, min_flank=5):
    """-> (reads [dict as longtr_b200.abi.region_collect], n_trim_failed)."""
    out, failed = [], 0
    for s, reads in enumerate(by_sample):
        for j, r in enumerate(reads):
            if r["pos"] > start or r["end"] < stop:
                failed += 1
                continue
            t = trim_alignment(r, start - flank if start > flank else 1, stop + flank, flank)
            hap_ok = int(not (min_flank > 0 and (r["pos"] > start - min_flank or r["end"] < stop + min_flank)))
            base = dict(name=r["name"], sample=s, log_p1=terms[s][j][0], log_p2=terms[s][j][1],
                        hp=(r["hp"] if has_tag(r["raw"], "HP") else 0))   # genotyper_bam_processor.cpp:146-151 counts these
            if not t["seq"]:
                out.append(dict(base, start=start, stop=stop, seq="", qual="", cigar="", hap_gen_ok=1, deleted=1))
                continue
            ops, soft, si, ri = [], False, 0, t["pos"]
            for op, n in t["cigar"]:
                if op == "H":
                    continue
                if op == "S":
                    ops.append((op, n)); si += n; soft = True
                elif op == "I":
                    ops.append((op, n)); si += n
                elif op == "D":
                    ops.append((op, n)); ri += n
                else:
                    prev, num = "=", 0
                    for _ in range(n):
                        c = "=" if t["seq"][si].upper() == ref[ri - ref_start].upper() else "X"
                        if c == prev:
                            num += 1
                        else:
                            if num:
                                ops.append((prev, num))
                            prev, num = c, 1
                        si += 1; ri += 1
                    if num:
                        ops.append((prev, num))
            if soft:
                failed += 1
                continue
            out.append(dict(base, start=t["pos"], stop=t["end"] - 1, seq=t["seq"].upper(), qual=t["qual"],
                            cigar="".join("%d%s" % (n, op) for op, n in ops), hap_gen_ok=hap_ok, deleted=int(t["deleted"])))
    return out, failed


# ---- the reference's own input layer over the library's BAM reader (oracle/io_driver.cpp) ------------------------------
def ref_io_available():
    return os.path.exists(_IO_SO)


def ref_io_region(path, chrom, start, end, span=None, trim=None):
    """BamCramReader on path: every alignment of chrom:[start, end); with span=(lo, hi), trim=(lo, hi): the spanning reads
    after BamAlignment::TrimAlignment(lo, hi)."""
    lib = C.CDLL(_IO_SO)
    lib.ltr_ref_io_region.restype = C.c_void_p
    lib.ltr_ref_io_region.argtypes = [C.c_char_p, C.c_char_p] + [C.c_int32] * 6
    lib.ltr_ref_io_free.argtypes = [C.c_void_p]
    s_lo, s_hi = span if span else (0, 0)
    t_lo, t_hi = trim if trim else (1, 0)
    p = lib.ltr_ref_io_region(path.encode(), chrom.encode(), start, end, s_lo, s_hi, t_lo, t_hi)
    text = C.string_at(p).decode()
    lib.ltr_ref_io_free(p)
    out = []
    for line in text.splitlines():
        f = line.split(" ")
        d = dict(name=f[0], pos=int(f[1]), end=int(f[2]), rev=int(f[3]), mapq=int(f[4]), cigar=f[5],
                 seq="" if f[6] == "*" else f[6], qual="" if f[7] == "*" else f[7], hp=int(f[8]))
        if trim:
            d["deleted"] = int(f[9])
        out.append(d)
    return out


# ---- the reference's own HaplotypeGenerator on flat reads (oracle/hapgen_driver.cpp) ------------------------------------
_HAPGEN_SO = os.path.join(_HERE, "_ref", "libltr_ref_hapgen.so")
_HAPGEN_POA_SO = os.path.join(_HERE, "_ref", "libltr_ref_hapgen_poa.so")  # spoa names served by oracle/poa_restatement.hpp


def ref_hapgen_available():
    return os.path.exists(_HAPGEN_SO)


def ref_hapgen_poa_available():
    return os.path.exists(_HAPGEN_POA_SO)


def ref_read_regions(path, max_regions=1000000000, chrom_limit=None):
    """readRegions + orderRegions of the reference -> [(chrom, start, stop, period, name, motif)]."""
    lib = C.CDLL(_HAPGEN_SO)
    lib.ltr_ref_read_regions.restype = C.c_void_p
    lib.ltr_ref_read_regions.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p]
    p = lib.ltr_ref_read_regions(path.encode(), max_regions, chrom_limit.encode() if chrom_limit else None)
    text = C.string_at(p).decode()
    C.CDLL(None).free(C.c_void_p(p))
    out = []
    for line in text.splitlines():
        c, s, e, per, name, motif = line.split(" ")
        out.append((c, int(s), int(e), int(per), "" if name == "." else name, motif))
    return out


_FASTA_SO = os.path.join(_HERE, "_ref", "libltr_ref_fasta.so")  # FastaReader + get_vcf_header on integration/faidx_compat.cpp


def ref_fasta_available():
    return os.path.exists(_FASTA_SO)


def ref_fasta_sequence(path, chrom):
    """FastaReader(path).get_sequence(chrom) of the reference, running on the library's FASTA reader -> (sequence, length)."""
    lib = C.CDLL(_FASTA_SO)
    lib.ltr_ref_fasta_sequence.restype = C.c_void_p
    lib.ltr_ref_fasta_sequence.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_longlong)]
    n = C.c_longlong(0)
    p = lib.ltr_ref_fasta_sequence(path.encode(), chrom.encode(), C.byref(n))
    text = C.string_at(p).decode()
    C.CDLL(None).free(C.c_void_p(p))
    return text, n.value


def ref_vcf_header(fasta_path, command, samples, switches=None):
    """Genotyper::get_vcf_header of the reference (contigs through its FastaReader on the library's reader); switches: the
    reference's output switches as a mask of the library's LTR_VCF_* bits."""
    lib = C.CDLL(_FASTA_SO)
    lib.ltr_ref_vcf_header.restype = C.c_void_p
    lib.ltr_ref_vcf_header.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
    lib.ltr_ref_vcf_header_switches.restype = C.c_void_p
    lib.ltr_ref_vcf_header_switches.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint]
    if switches is None:
        p = lib.ltr_ref_vcf_header(fasta_path.encode(), command.encode(), "\n".join(samples).encode())
    else:
        p = lib.ltr_ref_vcf_header_switches(fasta_path.encode(), command.encode(), "\n".join(samples).encode(), int(switches))
    text = C.string_at(p).decode()
    C.CDLL(None).free(C.c_void_p(p))
    return text


def ref_poa(seqs):
    """HaplotypeGenerator::poa (reference, compiled in place) on top of the restated spoa; fewer than 30 sequences."""
    import numpy as np
    assert len(seqs) < 30
    lib = C.CDLL(_HAPGEN_POA_SO)
    lib.ltr_ref_poa.restype = C.c_void_p
    lib.ltr_ref_poa.argtypes = [C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)]
    off = np.zeros(len(seqs) + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(x) for x in seqs])
    data = np.frombuffer(("".join(seqs) + "\0").encode(), dtype=np.uint8).copy()
    p = lib.ltr_ref_poa(len(seqs), off.ctypes.data_as(C.POINTER(C.c_uint32)), data.ctypes.data_as(C.POINTER(C.c_uint8)))
    text = C.string_at(p).decode()
    C.CDLL(None).free(C.c_void_p(p))
    return text


def ref_candidate_alleles(reads, n_samples, region_start, region_stop, motif, chrom_seq, indel_flank_len=5, assemble=False):
    """reads: dicts as longtr_b200.abi.region_collect returns them.  -> dict(status) or dict(block, lstart, lflank, rflank,
    alleles, inexact) from HaplotypeGenerator::add_haplotype_block + fuse_haplotype_blocks.  assemble: the build whose spoa
    names are served by the restatement (otherwise a region that reaches the consensus answers "needs assembly")."""
    import re

    import numpy as np
    lib = C.CDLL(_HAPGEN_POA_SO if assemble else _HAPGEN_SO)
    u32p, i32p, u8p = C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
    lib.ltr_ref_candidate_alleles.restype = C.c_void_p
    lib.ltr_ref_candidate_alleles.argtypes = [C.c_uint32, C.c_uint32, i32p, i32p, i32p, u32p, u8p, u32p, u32p, u8p, u8p,
                                              C.c_int32, C.c_int32, C.c_char_p, C.c_char_p, C.c_int32]
    n = len(reads)
    sample = np.array([r["sample"] for r in reads] + [0], dtype=np.int32)
    start = np.array([r["start"] for r in reads] + [0], dtype=np.int32)
    stop = np.array([r["stop"] for r in reads] + [0], dtype=np.int32)
    roff = np.zeros(n + 1, dtype=np.uint32)
    roff[1:] = np.cumsum([len(r["seq"]) for r in reads])
    rbytes = np.frombuffer(("".join(r["seq"] for r in reads) + "\0").encode(), dtype=np.uint8).copy()
    ops, coff = [], [0]
    for r in reads:
        for num, op in re.findall(r"(\d+)([MIDNSHP=X])", r["cigar"]):
            ops.append((int(num) << 4) | "MIDNSHP=X".index(op))
        coff.append(len(ops))
    ops = np.array(ops + [0], dtype=np.uint32)
    coff = np.array(coff, dtype=np.uint32)
    ok = np.array([r["hap_gen_ok"] for r in reads] + [0], dtype=np.uint8)
    dele = np.array([r["deleted"] for r in reads] + [0], dtype=np.uint8)
    P = lambda a, t: a.ctypes.data_as(t)
    p = lib.ltr_ref_candidate_alleles(n_samples, n, P(sample, i32p), P(start, i32p), P(stop, i32p), P(roff, u32p),
                                      P(rbytes, u8p), P(coff, u32p), P(ops, u32p), P(ok, u8p), P(dele, u8p), region_start,
                                      region_stop, motif.encode(), chrom_seq.encode(), indel_flank_len)
    text = C.string_at(p).decode()
    C.CDLL(None).free(C.c_void_p(p))
    if text.startswith("status="):
        return dict(status=text[7:])
    m = re.match(r"ok block=(-?\d+),(-?\d+) lstart=(-?\d+) lflank=(\S*) rflank=(\S*) alleles=(\S*) inexact=(\d*)$", text)
    return dict(status="ok", block_start=int(m.group(1)), block_end=int(m.group(2)), lflank_start=int(m.group(3)),
                lflank=m.group(4), rflank=m.group(5), alleles=re.findall(r"\[([^\]]*)\]", m.group(6)),
                inexact=[int(ch) for ch in m.group(7)])
