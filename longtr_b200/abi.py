"""ctypes binding of the C ABI in include/longtr_b200.h (liblongtr_b200.so).

This is the Python host side of the drop-in boundary: plain pointers and sizes only.
The library is built in-tree by ``longtr_b200.build`` (nvcc, sm_100a); importing this
module never falls back to a CPU implementation -- if the shared library or a CUDA
device is missing, the calls raise.
"""
import ctypes as C
import os

import numpy as np

from .flat import FlatLocus

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LONGTR_B200_LIB") or os.path.join(_HERE, "csrc", "liblongtr_b200.so")

LTR_OK = 0

_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)
_dp = C.POINTER(C.c_double)


class Params(C.Structure):
    """ltr_params: AlignmentModel of the reference (HapAligner.h:12-37)."""
    _fields_ = [("ins_ins", C.c_float), ("ins_match", C.c_float), ("del_del", C.c_float),
                ("del_match", C.c_float), ("match_match", C.c_float), ("match_ins", C.c_float),
                ("match_del", C.c_float), ("indel_flank_len", C.c_int32)]


class ViterbiBatch(C.Structure):
    _fields_ = [("n_loci", C.c_uint32), ("locus_hap_begin", _u32p), ("locus_read_begin", _u32p),
                ("hap_off", _u32p), ("hap_bytes", _u8p), ("read_off", _u32p), ("read_bytes", _u8p)]


class PosteriorBatch(C.Structure):
    _fields_ = [("locus_sread_begin", _u32p), ("pool_index", _u32p), ("sample_label", _i32p),
                ("log_p1", _dp), ("log_p2", _dp), ("locus_n_samples", _u32p), ("locus_haploid", _u8p),
                ("second_mate", _u8p), ("read_aligned", _u8p), ("prune_uncalled", C.c_int32)]


class JobOutputs(C.Structure):
    _fields_ = [("ll", _dp), ("post", _dp), ("totals", _dp), ("kept_mask", _u8p)]


class JobStats(C.Structure):
    _fields_ = [("n_pairs", C.c_uint64), ("n_cells", C.c_uint64), ("n_fallback", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("n_launches", C.c_uint32),
                ("kernel_ms", C.c_float), ("viterbi_ms", C.c_float), ("n_pairs_computed", C.c_uint64),
                ("n_cells_computed", C.c_uint64), ("n_band_pairs", C.c_uint64), ("n_band_uncertified", C.c_uint64),
                ("plan_ms", C.c_float), ("n_band_retried", C.c_uint64)]


class BamReads(C.Structure):
    """ltr_bam_reads (include/longtr_b200.h)."""
    _fields_ = [("n", C.c_uint32), ("tid", _i32p), ("pos", _i32p), ("end", _i32p), ("flag", C.POINTER(C.c_uint16)),
                ("mapq", _u8p), ("mate_tid", _i32p), ("mate_pos", _i32p), ("name_off", _u32p), ("names", C.c_void_p),
                ("seq_off", _u32p), ("seq", _u8p), ("qual", _u8p), ("cigar_off", _u32p), ("cigar_ops", _u32p),
                ("hp", _i32p), ("raw_off", _u32p), ("raw", _u8p), ("owner", C.c_void_p)]


class RegionParams(C.Structure):
    _fields_ = [("max_mate_dist", C.c_int32), ("min_mean_qual", C.c_double), ("min_mapq", C.c_double),
                ("require_spanning", C.c_int32), ("min_flank", C.c_int32), ("flank_size", C.c_int32),
                ("phased_bam", C.c_int32), ("check_hard_clips", C.c_int32)]


class RegionReads(C.Structure):
    _fields_ = [("n_samples", C.c_uint32), ("sample_file", _u32p), ("sample_read_begin", _u32p), ("n_reads", C.c_uint32),
                ("read_start", _i32p), ("read_stop", _i32p), ("read_off", _u32p), ("read_bytes", _u8p), ("qual_bytes", _u8p),
                ("cigar_off", _u32p), ("cigar_ops", _u32p), ("read_sample", _i32p), ("log_p1", _dp), ("log_p2", _dp),
                ("hap_gen_ok", _u8p), ("deleted", _u8p), ("name_off", _u32p), ("names", C.c_void_p),
                ("n_overlapping", C.c_uint32), ("n_hard_clipped", C.c_uint32), ("n_has_n", C.c_uint32),
                ("n_low_qual", C.c_uint32), ("n_low_mapq", C.c_uint32), ("n_not_spanning", C.c_uint32),
                ("n_not_unique", C.c_uint32), ("n_passed", C.c_uint32), ("n_trim_failed", C.c_uint32), ("read_hp", _i32p),
                ("owner", C.c_void_p)]


class Candidates(C.Structure):
    _fields_ = [("status", C.c_int32), ("block_start", C.c_int32), ("block_end", C.c_int32), ("n_alleles", C.c_int32),
                ("allele_off", _u32p), ("allele_bytes", _u8p), ("lflank_start", C.c_int32), ("lflank", C.c_char_p),
                ("rflank", C.c_char_p), ("n_cluster_samples", C.c_uint32), ("cluster_sample_begin", _u32p),
                ("cluster_off", _u32p), ("cluster_bytes", _u8p), ("cluster_count", _i32p), ("allele_inexact", _u8p),
                ("n_consensus", C.c_uint32), ("assembly_threshold", C.c_int32), ("owner", C.c_void_p)]


def _candidates_dict(c):
    ab = C.string_at(c.allele_bytes, c.allele_off[c.n_alleles]) if c.n_alleles else b""
    ncs = c.cluster_sample_begin[c.n_cluster_samples] if c.n_cluster_samples else 0
    cb = C.string_at(c.cluster_bytes, c.cluster_off[ncs]) if ncs else b""
    return dict(
        status=c.status, block_start=c.block_start, block_end=c.block_end, lflank_start=c.lflank_start,
        alleles=[ab[c.allele_off[k]:c.allele_off[k + 1]].decode() for k in range(c.n_alleles)],
        lflank=(c.lflank or b"").decode(), rflank=(c.rflank or b"").decode(),
        inexact=[int(c.allele_inexact[k]) for k in range(c.n_alleles)], n_consensus=c.n_consensus,
        assembly_threshold=c.assembly_threshold,
        cluster_sets=[[(cb[c.cluster_off[k]:c.cluster_off[k + 1]].decode(), c.cluster_count[k])
                       for k in range(c.cluster_sample_begin[s], c.cluster_sample_begin[s + 1])]
                      for s in range(c.n_cluster_samples)])


def candidate_alleles_from_reads(reads, n_samples, region_start, region_stop, period, ref_seq, ref_seq_start=0,
                                 indel_flank_len=5, flags=0):
    """ltr_candidate_alleles_flags on reads the caller holds (dicts with start, stop, seq, cigar, sample; optional hap_gen_ok,
    deleted; sample-major) instead of reads ltr_region_collect prepared.  Host only."""
    import re
    lib = load()
    n = len(reads)
    start = np.array([r["start"] for r in reads] + [0], dtype=np.int32)
    stop = np.array([r["stop"] for r in reads] + [0], dtype=np.int32)
    sample = np.array([r["sample"] for r in reads] + [0], dtype=np.int32)
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
    ok = np.array([r.get("hap_gen_ok", 1) for r in reads] + [0], dtype=np.uint8)
    dele = np.array([r.get("deleted", 0) for r in reads] + [0], dtype=np.uint8)
    R = RegionReads()
    R.n_samples = n_samples
    R.n_reads = n
    R.read_start, R.read_stop, R.read_sample = ptr(start, _i32p), ptr(stop, _i32p), ptr(sample, _i32p)
    R.read_off, R.read_bytes = ptr(roff, _u32p), ptr(rbytes, _u8p)
    R.cigar_off, R.cigar_ops = ptr(coff, _u32p), ptr(ops, _u32p)
    R.hap_gen_ok, R.deleted = ptr(ok, _u8p), ptr(dele, _u8p)
    ref = np.frombuffer(ref_seq.encode() if isinstance(ref_seq, str) else bytes(ref_seq), dtype=np.uint8)
    cp = C.POINTER(Candidates)()
    rc = lib.ltr_candidate_alleles_flags(C.byref(R), region_start, region_stop, period, ptr(ref, _u8p), ref_seq_start, len(ref),
                                         indel_flank_len, flags, C.byref(cp))
    if rc != 0:
        raise RuntimeError("ltr_candidate_alleles failed: %d" % rc)
    out = _candidates_dict(cp.contents)
    lib.ltr_candidates_free(cp)
    return out


def region_collect(bams, chrom, start, stop, ref_seq, ref_seq_start=0, candidates=None, **overrides):
    """ltr_region_collect -> dict(samples=[file index], reads=[dict per read, sample-major], counters).
    candidates=dict(period=.., indel_flank_len=5, flags=0): also ltr_candidate_alleles_flags on the same reads ->
    res["candidates"] (flags=1: stop before the assembly)."""
    lib = load()
    prm = RegionParams()
    lib.ltr_region_params_default(C.byref(prm))
    for k, v in overrides.items():
        setattr(prm, k, v)
    handles = (C.c_void_p * len(bams))(*[b.h for b in bams])
    ref = np.frombuffer(ref_seq.encode() if isinstance(ref_seq, str) else bytes(ref_seq), dtype=np.uint8)
    out = C.POINTER(RegionReads)()
    rc = lib.ltr_region_collect(handles, len(bams), chrom.encode(), start, stop, ptr(ref, _u8p), ref_seq_start, len(ref),
                                C.byref(prm), C.byref(out))
    if rc != 0:
        raise RuntimeError("ltr_region_collect failed: %d" % rc)
    r = out.contents
    n = r.n_reads
    seq = C.string_at(r.read_bytes, r.read_off[n]) if n else b""
    qual = C.string_at(r.qual_bytes, r.read_off[n]) if n else b""
    names = C.string_at(r.names, r.name_off[n]) if n else b""
    reads = []
    for i in range(n):
        cig = "".join("%d%s" % (r.cigar_ops[k] >> 4, "MIDNSHP=X"[r.cigar_ops[k] & 15])
                      for k in range(r.cigar_off[i], r.cigar_off[i + 1]))
        reads.append(dict(name=names[r.name_off[i]:r.name_off[i + 1] - 1].decode(), start=r.read_start[i], stop=r.read_stop[i],
                          seq=seq[r.read_off[i]:r.read_off[i + 1]].decode(), qual=qual[r.read_off[i]:r.read_off[i + 1]].decode(),
                          cigar=cig, sample=r.read_sample[i], log_p1=r.log_p1[i], log_p2=r.log_p2[i],
                          hap_gen_ok=int(r.hap_gen_ok[i]), deleted=int(r.deleted[i]), hp=int(r.read_hp[i])))
    res = dict(samples=[r.sample_file[s] for s in range(r.n_samples)], reads=reads,
               counters={k: getattr(r, k) for k in ("n_overlapping", "n_hard_clipped", "n_has_n", "n_low_qual", "n_low_mapq",
                                                    "n_not_spanning", "n_not_unique", "n_passed", "n_trim_failed")})
    if candidates is not None:
        cp = C.POINTER(Candidates)()
        rc = lib.ltr_candidate_alleles_flags(out, start, stop, candidates["period"], ptr(ref, _u8p), ref_seq_start, len(ref),
                                             candidates.get("indel_flank_len", 5), candidates.get("flags", 0), C.byref(cp))
        if rc != 0:
            lib.ltr_region_reads_free(out)
            raise RuntimeError("ltr_candidate_alleles failed: %d" % rc)
        res["candidates"] = _candidates_dict(cp.contents)
        lib.ltr_candidates_free(cp)
    lib.ltr_region_reads_free(out)
    return res


class Region(C.Structure):
    _fields_ = [("start", C.c_int32), ("stop", C.c_int32), ("period", C.c_int32)]


class Bed(C.Structure):
    _fields_ = [("n_regions", C.c_uint32), ("regions", C.POINTER(Region)), ("region_chrom", _i32p),
                ("names", C.POINTER(C.c_char_p)), ("motifs", C.POINTER(C.c_char_p)), ("n_chroms", C.c_uint32),
                ("chroms", C.POINTER(C.c_char_p)), ("owner", C.c_void_p)]


class FastaFile:
    """ltr_fasta_*: indexed FASTA access (one file or a directory of *.fa files; host only)."""

    def __init__(self, path):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.ltr_fasta_open(path.encode(), C.byref(h))
        if rc != 0:
            raise RuntimeError("ltr_fasta_open(%s) failed: %d" % (path, rc))
        self.h = h
        self.names = [self.lib.ltr_fasta_seq_name(h, i).decode() for i in range(self.lib.ltr_fasta_n_seqs(h))]

    def length(self, name):
        return self.lib.ltr_fasta_seq_len(self.h, name.encode())

    def fetch(self, name, start=0, end=None):
        if end is None:
            end = self.length(name)
        buf = np.zeros(max(1, end - start), dtype=np.uint8)
        rc = self.lib.ltr_fasta_fetch(self.h, name.encode(), start, end, ptr(buf, _u8p))
        if rc != 0:
            raise RuntimeError("ltr_fasta_fetch failed: %d" % rc)
        return buf[:max(0, end - start)].tobytes().decode()

    def close(self):
        if self.h:
            self.lib.ltr_fasta_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def bed_read(path, max_regions=0, chrom_limit=None, keep_handle=False):
    """ltr_bed_read -> dict(chroms, regions=[(chrom index, start, stop, period, name, motif)]) (+ handle with keep_handle;
    free it with load().ltr_bed_free)."""
    lib = load()
    bp = C.POINTER(Bed)()
    rc = lib.ltr_bed_read(path.encode(), max_regions, chrom_limit.encode() if chrom_limit else None, C.byref(bp))
    if rc != 0:
        raise RuntimeError("ltr_bed_read failed: %d" % rc)
    b = bp.contents
    res = dict(chroms=[b.chroms[i].decode() for i in range(b.n_chroms)],
               regions=[(b.region_chrom[i], b.regions[i].start, b.regions[i].stop, b.regions[i].period, b.names[i].decode(),
                         b.motifs[i].decode()) for i in range(b.n_regions)])
    if keep_handle:
        res["handle"] = bp
    else:
        lib.ltr_bed_free(bp)
    return res


class BamFile:
    """ltr_bam_*: BGZF / BAM / BAI reader of the library (host only)."""

    def __init__(self, path, index_path=None):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.ltr_bam_open(path.encode(), index_path.encode() if index_path else None, C.byref(h))
        if rc != 0:
            raise RuntimeError("ltr_bam_open(%s) failed: %d" % (path, rc))
        self.h = h
        self.refs = [(self.lib.ltr_bam_ref_name(h, t).decode(), self.lib.ltr_bam_ref_len(h, t))
                     for t in range(self.lib.ltr_bam_n_refs(h))]
        self.has_index = bool(self.lib.ltr_bam_has_index(h))

    def build_index(self):
        rc = self.lib.ltr_bam_build_index(self.h)
        if rc != 0:
            raise RuntimeError("ltr_bam_build_index failed: %d" % rc)
        self.has_index = True

    def fetch(self, tid=-1, beg=0, end=1 << 29, keep_raw=False):
        """Records overlapping [beg, end) of reference tid (tid < 0: all) as a list of dicts."""
        p = C.POINTER(BamReads)()
        rc = self.lib.ltr_bam_fetch(self.h, tid, beg, end, 1 if keep_raw else 0, C.byref(p))
        if rc != 0:
            raise RuntimeError("ltr_bam_fetch failed: %d" % rc)
        r = p.contents
        n = r.n
        out = []
        names = C.string_at(r.names, r.name_off[n]) if n else b""
        seq = C.string_at(r.seq, r.seq_off[n]) if n else b""
        qual = C.string_at(r.qual, r.seq_off[n]) if n else b""
        raw = C.string_at(r.raw, r.raw_off[n]) if n and r.raw_off[n] else b""
        for i in range(n):
            cig = [("MIDNSHP=X"[r.cigar_ops[k] & 15], r.cigar_ops[k] >> 4) for k in range(r.cigar_off[i], r.cigar_off[i + 1])]
            out.append(dict(name=names[r.name_off[i]:r.name_off[i + 1] - 1].decode(), flag=r.flag[i], tid=r.tid[i],
                            pos=r.pos[i], end=r.end[i], mapq=r.mapq[i], cigar=cig, mate_tid=r.mate_tid[i],
                            mate_pos=r.mate_pos[i], seq=seq[r.seq_off[i]:r.seq_off[i + 1]].decode(),
                            qual=qual[r.seq_off[i]:r.seq_off[i + 1]].decode(), hp=r.hp[i],
                            raw=raw[r.raw_off[i]:r.raw_off[i + 1]]))
        self.lib.ltr_bam_reads_free(p)
        return out

    def close(self):
        if self.h:
            self.lib.ltr_bam_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LocusCalls(C.Structure):
    """ltr_locus_calls (include/longtr_b200.h)."""
    _fields_ = [("best_gts", _i32p), ("log_phased_posteriors", _dp), ("log_unphased_posteriors", _dp),
                ("hap_log_phased_posteriors", _dp), ("hap_log_unphased_posteriors", _dp), ("gls", _dp),
                ("pls", _i32p), ("phased_gls", _dp), ("gl_diffs", _dp), ("sample_total_lls", _dp),
                ("log_sample_posteriors", _dp), ("total_ll", C.c_double)]


def make_locus_calls(S, H, haploid):
    """Allocates every output array of ltr_locus_calls; returns (struct, dict of numpy arrays)."""
    n_gl = H if haploid else H * (H + 1) // 2
    n_pgl = H if haploid else H * H
    a = dict(best_gts=np.zeros((S, 2), np.int32), log_phased_posteriors=np.zeros(S), log_unphased_posteriors=np.zeros(S),
             hap_log_phased_posteriors=np.zeros(S), hap_log_unphased_posteriors=np.zeros(S), gls=np.zeros((S, n_gl)),
             pls=np.zeros((S, n_gl), np.int32), phased_gls=np.zeros((S, n_pgl)), gl_diffs=np.zeros(S),
             sample_total_lls=np.zeros(S), log_sample_posteriors=np.zeros((S, H, H)))
    c = LocusCalls()
    for k, v in a.items():
        setattr(c, k, v.ctypes.data_as(_i32p if v.dtype == np.int32 else _dp))
    return c, a


DEFAULT_ALN_PARAMS = (-1.0, -0.458675, -1.0, -0.458675, -0.00005800168, -10.448214728, -10.448214728)


def make_params(aln_params=None, indel_flank_len=5):
    p = Params()
    vals = DEFAULT_ALN_PARAMS if aln_params is None else aln_params
    (p.ins_ins, p.ins_match, p.del_del, p.del_match, p.match_match, p.match_ins, p.match_del) = \
        [float(x) for x in vals]
    p.indel_flank_len = indel_flank_len
    return p


def ptr(a, t):
    return a.ctypes.data_as(t)


def make_viterbi_batch(batch):
    """dict of numpy arrays -> (ViterbiBatch, keepalive list)."""
    keep = dict(
        lhb=np.ascontiguousarray(batch["locus_hap_begin"], dtype=np.uint32),
        lrb=np.ascontiguousarray(batch["locus_read_begin"], dtype=np.uint32),
        hoff=np.ascontiguousarray(batch["hap_off"], dtype=np.uint32),
        roff=np.ascontiguousarray(batch["read_off"], dtype=np.uint32),
        hb=np.ascontiguousarray(batch["hap_bytes"], dtype=np.uint8),
        rb=np.ascontiguousarray(batch["read_bytes"], dtype=np.uint8))
    b = ViterbiBatch()
    b.n_loci = len(keep["lhb"]) - 1
    b.locus_hap_begin = ptr(keep["lhb"], _u32p)
    b.locus_read_begin = ptr(keep["lrb"], _u32p)
    b.hap_off = ptr(keep["hoff"], _u32p)
    b.hap_bytes = ptr(keep["hb"], _u8p)
    b.read_off = ptr(keep["roff"], _u32p)
    b.read_bytes = ptr(keep["rb"], _u8p)
    return b, keep


def ll_size(batch):
    lhb = np.asarray(batch["locus_hap_begin"], dtype=np.int64)
    lrb = np.asarray(batch["locus_read_begin"], dtype=np.int64)
    return int(np.sum((lhb[1:] - lhb[:-1]) * (lrb[1:] - lrb[:-1])))


def make_posterior_batch(post):
    keep = dict(
        lsb=np.ascontiguousarray(post["locus_sread_begin"], dtype=np.uint32),
        pool=np.ascontiguousarray(post["pool_index"], dtype=np.uint32),
        lab=np.ascontiguousarray(post["sample_label"], dtype=np.int32),
        p1=np.ascontiguousarray(post["log_p1"], dtype=np.float64),
        p2=np.ascontiguousarray(post["log_p2"], dtype=np.float64),
        ns=np.ascontiguousarray(post["locus_n_samples"], dtype=np.uint32))
    b = PosteriorBatch()
    b.locus_sread_begin = ptr(keep["lsb"], _u32p)
    b.pool_index = ptr(keep["pool"], _u32p)
    b.sample_label = ptr(keep["lab"], _i32p)
    b.log_p1 = ptr(keep["p1"], _dp)
    b.log_p2 = ptr(keep["p2"], _dp)
    b.locus_n_samples = ptr(keep["ns"], _u32p)
    if post.get("locus_haploid") is not None:
        keep["hap"] = np.ascontiguousarray(post["locus_haploid"], dtype=np.uint8)
        b.locus_haploid = ptr(keep["hap"], _u8p)
    for key in ("second_mate", "read_aligned"):
        if post.get(key) is not None:
            keep[key] = np.ascontiguousarray(post[key], dtype=np.uint8)
            setattr(b, key, ptr(keep[key], _u8p))
    b.prune_uncalled = 1 if post.get("prune_uncalled") else 0
    return b, keep


_lib = None


def load():
    """Load liblongtr_b200.so and declare every symbol of include/longtr_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("liblongtr_b200.so is not built (run `python -m longtr_b200.build`); "
                           "there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    lib.ltr_params_default.argtypes = [C.POINTER(Params)]
    lib.ltr_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.ltr_ctx_create.restype = C.c_int
    lib.ltr_ctx_destroy.argtypes = [vp]
    lib.ltr_ctx_set_band.argtypes = [vp, C.c_int32]
    lib.ltr_ctx_set_band.restype = C.c_int
    lib.ltr_strerror.argtypes = [C.c_int]
    lib.ltr_strerror.restype = C.c_char_p
    lib.ltr_last_error.argtypes = [vp]
    lib.ltr_last_error.restype = C.c_char_p
    lib.ltr_version.restype = C.c_char_p
    lib.ltr_viterbi_ll.argtypes = [vp, C.POINTER(Params), C.POINTER(ViterbiBatch), _dp, C.POINTER(JobStats)]
    lib.ltr_viterbi_ll.restype = C.c_int
    lib.ltr_posteriors.argtypes = [vp, C.c_int, C.c_int32, C.c_int32, C.c_int32, _dp, _dp, _dp, _i32p,
                                   _dp, _dp, _dp]
    lib.ltr_posteriors.restype = C.c_int
    lib.ltr_job_create.argtypes = [vp, C.POINTER(Params), C.POINTER(ViterbiBatch), C.POINTER(PosteriorBatch),
                                   C.POINTER(vp)]
    lib.ltr_job_create.restype = C.c_int
    lib.ltr_job_run.argtypes = [vp, vp]
    lib.ltr_job_run.restype = C.c_int
    lib.ltr_job_sizes.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.ltr_job_download.argtypes = [vp, vp, _dp, _dp, _dp]
    lib.ltr_job_download.restype = C.c_int
    lib.ltr_job_get_stats.argtypes = [vp, C.POINTER(JobStats)]
    lib.ltr_job_destroy.argtypes = [vp, vp]
    lib.ltr_ctx_set_plan.argtypes = [vp, C.c_int32]
    lib.ltr_ctx_set_plan.restype = C.c_int
    lib.ltr_ctx_set_read_encoding.argtypes = [vp, C.c_int32]
    lib.ltr_ctx_set_read_encoding.restype = C.c_int
    lib.ltr_job_submit.argtypes = [vp, C.POINTER(Params), C.POINTER(ViterbiBatch), C.POINTER(PosteriorBatch), _dp, _dp, _dp,
                                   C.POINTER(vp)]
    lib.ltr_job_submit.restype = C.c_int
    lib.ltr_job_submit_outputs.argtypes = [vp, C.POINTER(Params), C.POINTER(ViterbiBatch), C.POINTER(PosteriorBatch),
                                           C.POINTER(JobOutputs), C.POINTER(vp)]
    lib.ltr_job_submit_outputs.restype = C.c_int
    lib.ltr_job_download_kept.argtypes = [vp, vp, _u8p]
    lib.ltr_job_download_kept.restype = C.c_int
    lib.ltr_posteriors_batch.argtypes = [vp, C.c_uint32, _u32p, _u32p, _dp, C.POINTER(PosteriorBatch), _dp, _dp, _u8p]
    lib.ltr_posteriors_batch.restype = C.c_int
    lib.ltr_job_wait.argtypes = [vp, vp]
    lib.ltr_job_wait.restype = C.c_int
    lib.ltr_job_poll.argtypes = [vp, vp]
    lib.ltr_job_poll.restype = C.c_int
    lib.ltr_process_reads_flat.argtypes = [vp, C.POINTER(FlatLocus), _dp, _i32p]
    lib.ltr_process_reads_flat.restype = C.c_int
    lib.ltr_process_reads_flat_batch.argtypes = [vp, C.c_int32, C.POINTER(FlatLocus), C.POINTER(_dp), C.POINTER(_i32p)]
    lib.ltr_process_reads_flat_batch.restype = C.c_int
    lib.ltr_pipeline_create.argtypes = [C.c_int, C.c_int32, C.c_int32, C.POINTER(vp)]
    lib.ltr_pipeline_create.restype = C.c_int
    lib.ltr_pipeline_submit.argtypes = [vp, C.POINTER(FlatLocus), C.c_uint64, _i32p, C.c_double]
    lib.ltr_pipeline_submit.restype = C.c_int
    lib.ltr_pipeline_flush.argtypes = [vp]
    lib.ltr_pipeline_flush.restype = C.c_int
    lib.ltr_pipeline_next.argtypes = [vp, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                      C.POINTER(_dp), C.POINTER(_i32p), C.POINTER(C.c_int)]
    lib.ltr_pipeline_next.restype = C.c_int
    lib.ltr_pipeline_destroy.argtypes = [vp]
    lib.ltr_pipeline_destroy.restype = None
    lib.ltr_genotype_locus.argtypes = [vp, C.c_int, C.c_int32, _i32p, C.c_int32, _dp, _dp, _dp, C.POINTER(LocusCalls)]
    lib.ltr_genotype_locus.restype = C.c_int
    lib.ltr_genotype_locus_pruned.argtypes = [vp, C.c_int, C.c_int32, _i32p, C.c_int32, _dp, _dp, _dp, _i32p, _i32p, _i32p,
                                              C.POINTER(LocusCalls)]
    lib.ltr_genotype_locus_pruned.restype = C.c_int
    lib.ltr_extract_calls.argtypes = [C.c_int, C.c_int32, C.c_int32, _dp, _dp, C.POINTER(LocusCalls)]
    lib.ltr_extract_calls.restype = C.c_int
    lib.ltr_trim_read_flat.argtypes = [C.POINTER(FlatLocus), C.c_int32, C.c_char_p, C.c_int32]
    lib.ltr_trim_read_flat.restype = C.c_int32
    lib.ltr_seed_base_flat.argtypes = [C.POINTER(FlatLocus), C.c_int32]
    lib.ltr_seed_base_flat.restype = C.c_int32
    lib.ltr_stutter_ll.argtypes = [vp, C.POINTER(Params), C.c_void_p, _dp, C.POINTER(JobStats)]
    lib.ltr_stutter_ll.restype = C.c_int
    lib.ltr_stutter_ll_status.argtypes = [vp, C.POINTER(Params), C.c_void_p, _dp, _i32p, C.POINTER(JobStats)]
    lib.ltr_stutter_ll_status.restype = C.c_int
    lib.ltr_fp64_issue_rate.argtypes = [C.c_int, C.c_int, _dp, _dp]
    lib.ltr_fp64_issue_rate.restype = C.c_int
    lib.ltr_bam_open.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(vp)]
    lib.ltr_bam_open.restype = C.c_int
    lib.ltr_bam_close.argtypes = [vp]
    lib.ltr_bam_close.restype = None
    lib.ltr_bam_n_refs.argtypes = [vp]
    lib.ltr_bam_n_refs.restype = C.c_int32
    lib.ltr_bam_ref_name.argtypes = [vp, C.c_int32]
    lib.ltr_bam_ref_name.restype = C.c_char_p
    lib.ltr_bam_ref_len.argtypes = [vp, C.c_int32]
    lib.ltr_bam_ref_len.restype = C.c_int64
    lib.ltr_bam_ref_id.argtypes = [vp, C.c_char_p]
    lib.ltr_bam_ref_id.restype = C.c_int32
    lib.ltr_bam_header_text.argtypes = [vp]
    lib.ltr_bam_header_text.restype = C.c_char_p
    lib.ltr_bam_has_index.argtypes = [vp]
    lib.ltr_bam_has_index.restype = C.c_int
    lib.ltr_bam_build_index.argtypes = [vp]
    lib.ltr_bam_build_index.restype = C.c_int
    lib.ltr_bam_fetch.argtypes = [vp, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.POINTER(C.POINTER(BamReads))]
    lib.ltr_bam_fetch.restype = C.c_int
    lib.ltr_bam_reads_free.argtypes = [C.POINTER(BamReads)]
    lib.ltr_bam_reads_free.restype = None
    lib.ltr_fasta_open.argtypes = [C.c_char_p, C.POINTER(vp)]
    lib.ltr_fasta_open.restype = C.c_int
    lib.ltr_fasta_close.argtypes = [vp]
    lib.ltr_fasta_close.restype = None
    lib.ltr_fasta_n_seqs.argtypes = [vp]
    lib.ltr_fasta_n_seqs.restype = C.c_int32
    lib.ltr_fasta_seq_name.argtypes = [vp, C.c_int32]
    lib.ltr_fasta_seq_name.restype = C.c_char_p
    lib.ltr_fasta_seq_len.argtypes = [vp, C.c_char_p]
    lib.ltr_fasta_seq_len.restype = C.c_int64
    lib.ltr_fasta_fetch.argtypes = [vp, C.c_char_p, C.c_int64, C.c_int64, _u8p]
    lib.ltr_fasta_fetch.restype = C.c_int
    lib.ltr_bed_read.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.POINTER(C.POINTER(Bed))]
    lib.ltr_bed_read.restype = C.c_int
    lib.ltr_bed_free.argtypes = [C.POINTER(Bed)]
    lib.ltr_bed_free.restype = None
    lib.ltr_region_params_default.argtypes = [C.POINTER(RegionParams)]
    lib.ltr_region_params_default.restype = None
    lib.ltr_region_collect.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.c_char_p, C.c_int32, C.c_int32, _u8p, C.c_int64,
                                       C.c_int64, C.POINTER(RegionParams), C.POINTER(C.POINTER(RegionReads))]
    lib.ltr_region_collect.restype = C.c_int
    lib.ltr_region_reads_free.argtypes = [C.POINTER(RegionReads)]
    lib.ltr_region_reads_free.restype = None
    lib.ltr_candidate_alleles.argtypes = [C.POINTER(RegionReads), C.c_int32, C.c_int32, C.c_int32, _u8p, C.c_int64, C.c_int64,
                                          C.c_int32, C.POINTER(C.POINTER(Candidates))]
    lib.ltr_candidate_alleles.restype = C.c_int
    lib.ltr_candidate_alleles_flags.argtypes = [C.POINTER(RegionReads), C.c_int32, C.c_int32, C.c_int32, _u8p, C.c_int64,
                                                C.c_int64, C.c_int32, C.c_uint32, C.POINTER(C.POINTER(Candidates))]
    lib.ltr_candidate_alleles_flags.restype = C.c_int
    lib.ltr_poa_consensus.argtypes = [_u8p, _u32p, C.c_uint32, _u8p, C.c_uint32, _u32p]
    lib.ltr_poa_consensus.restype = C.c_int
    lib.ltr_vcf_header.argtypes = [vp, C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_uint32, C.c_char_p, C.c_uint32, _u32p]
    lib.ltr_vcf_header.restype = C.c_int
    lib.ltr_vcf_record.argtypes = [C.POINTER(VcfLocus), C.c_char_p, C.c_uint32, _u32p]
    lib.ltr_vcf_record.restype = C.c_int
    lib.ltr_vcf_record_ex.argtypes = [C.POINTER(VcfLocus), C.POINTER(VcfExtras), C.c_char_p, C.c_uint32, _u32p]
    lib.ltr_vcf_record_ex.restype = C.c_int
    lib.ltr_vcf_header_ex.argtypes = [vp, C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_uint32, C.c_uint32, C.c_char_p,
                                      C.c_uint32, _u32p]
    lib.ltr_vcf_header_ex.restype = C.c_int
    lib.ltr_extract_cigar_bp_diff.argtypes = [_u32p, C.c_uint32, C.c_int32, C.c_int32, C.c_int32, _i32p]
    lib.ltr_extract_cigar_bp_diff.restype = C.c_int
    lib.ltr_em_opts_default.argtypes = [C.POINTER(EmOpts)]
    lib.ltr_em_opts_default.restype = None
    lib.ltr_em_stutter_train.argtypes = [vp, C.POINTER(EmBatch), C.POINTER(EmOpts), _dp, _i32p, _i32p, _dp, _dp, C.c_uint32]
    lib.ltr_em_stutter_train.restype = C.c_int
    lib.ltr_candidates_free.argtypes = [C.POINTER(Candidates)]
    lib.ltr_candidates_free.restype = None
    lib.ltr_edit_distances.argtypes = [vp, _u8p, _u32p, C.c_uint32, _u32p, _u32p, _i32p, C.c_uint32, _i32p,
                                       C.POINTER(JobStats)]
    lib.ltr_edit_distances.restype = C.c_int
    lib.ltr_cluster_greedy.argtypes = [vp, _u8p, _u32p, C.c_uint32, _u32p, _u32p, _i32p, C.c_uint32, _i32p, _i32p, _u8p,
                                       C.POINTER(JobStats)]
    lib.ltr_cluster_greedy.restype = C.c_int
    _lib = lib
    return lib


EXPORTED_SYMBOLS = [
    "ltr_params_default", "ltr_ctx_create", "ltr_ctx_destroy", "ltr_ctx_set_band", "ltr_strerror", "ltr_last_error",
    "ltr_version", "ltr_viterbi_ll", "ltr_posteriors", "ltr_job_create", "ltr_job_run", "ltr_job_sizes",
    "ltr_job_download", "ltr_job_get_stats", "ltr_job_destroy", "ltr_process_reads_flat",
    "ltr_process_reads_flat_batch", "ltr_pipeline_create", "ltr_pipeline_submit", "ltr_pipeline_flush", "ltr_pipeline_next",
    "ltr_pipeline_destroy", "ltr_flatten_loci", "ltr_flat_batch_free",
    "ltr_fp64_issue_rate", "ltr_genotype_locus", "ltr_extract_calls", "ltr_trim_read_flat", "ltr_seed_base_flat",
    "ltr_stutter_ll", "ltr_genotype_locus_pruned", "ltr_ctx_set_plan", "ltr_ctx_set_read_encoding", "ltr_job_submit", "ltr_job_submit_outputs", "ltr_job_wait", "ltr_job_poll", "ltr_job_download_kept", "ltr_posteriors_batch", "ltr_genotyper_create", "ltr_genotyper_destroy",
    "ltr_genotyper_run", "ltr_batch_calls_free", "ltr_locus_batch_trim_read", "ltr_stutter_ll_status", "ltr_pool_reads",
    "ltr_edit_distances", "ltr_cluster_greedy",
    "ltr_bam_open", "ltr_bam_close", "ltr_bam_n_refs", "ltr_bam_ref_name", "ltr_bam_ref_len", "ltr_bam_ref_id",
    "ltr_bam_header_text", "ltr_bam_has_index", "ltr_bam_build_index", "ltr_bam_fetch", "ltr_bam_reads_free",
    "ltr_region_params_default", "ltr_region_collect", "ltr_region_reads_free",
    "ltr_candidate_alleles", "ltr_candidate_alleles_flags", "ltr_poa_consensus", "ltr_candidates_free", "ltr_regions_opts_default", "ltr_regions_run",
    "ltr_regions_result_free", "ltr_fasta_open", "ltr_fasta_close", "ltr_fasta_n_seqs", "ltr_fasta_seq_name",
    "ltr_fasta_seq_len", "ltr_fasta_fetch", "ltr_bed_read", "ltr_bed_free", "ltr_run_bed", "ltr_bed_run_result_free",
    "ltr_em_opts_default", "ltr_em_stutter_train", "ltr_vcf_record", "ltr_vcf_header", "ltr_extract_cigar_bp_diff", "ltr_genotyper_set_read_alleles",
    "ltr_vcf_record_ex", "ltr_vcf_header_ex", "ltr_genotyper_set_phased_gls", "ltr_run_bed_stream",
]


class VcfLocus(C.Structure):
    _fields_ = [("chrom", C.c_char_p), ("name", C.c_char_p), ("motif", C.c_char_p), ("region_start", C.c_int32),
                ("region_stop", C.c_int32), ("chrom_seq", _u8p), ("chrom_seq_start", C.c_int64), ("chrom_seq_len", C.c_int64),
                ("block_start", C.c_int32), ("block_end", C.c_int32), ("n_alleles", C.c_int32), ("allele_off", _u32p),
                ("allele_bytes", _u8p), ("allele_inexact", _u8p), ("kept_mask", _u8p), ("haploid", C.c_int32),
                ("n_samples", C.c_int32), ("gts", _i32p), ("log_unphased_posteriors", _dp), ("log_phased_posteriors", _dp),
                ("gl_diffs", _dp), ("n_p1", _i32p), ("n_p2", _i32p), ("n_reads", C.c_int32), ("read_sample", _i32p),
                ("log_p1", _dp), ("log_p2", _dp), ("read_bp_diff", _i32p), ("read_allele", _i32p), ("n_columns", C.c_int32),
                ("column_sample", _i32p)]


_u64p = C.POINTER(C.c_uint64)


class VcfExtras(C.Structure):
    _fields_ = [("switches", C.c_uint32), ("gl_begin", _u64p), ("gls", _dp), ("pls", _i32p), ("pgl_begin", _u64p),
                ("phased_gls", _dp)]


VCF_ALLREADS, VCF_MALLREADS, VCF_GLS, VCF_PLS, VCF_PHASED_GLS, VCF_FILTERS = 1, 2, 4, 8, 16, 32
VCF_DEFAULT = VCF_ALLREADS | VCF_MALLREADS
INT32_MIN = -2147483648


def extract_cigar_bp_diff(cigar, cigar_start, region_start, region_end):
    """ltr_extract_cigar_bp_diff on a CIGAR string -> bp difference or None."""
    import re
    lib = load()
    ops = np.array([(int(n) << 4) | "MIDNSHP=X".index(o) for n, o in re.findall(r"(\d+)([MIDNSHP=X])", cigar)] + [0],
                   dtype=np.uint32)
    d = C.c_int32(0)
    ok = lib.ltr_extract_cigar_bp_diff(ptr(ops, _u32p), len(ops) - 1, cigar_start, region_start, region_end, C.byref(d))
    return d.value if ok else None


def vcf_record(chrom, name, motif, region_start, region_stop, chrom_seq, chrom_seq_start, block_start, block_end, alleles,
               inexact, kept_mask, gts, log_unphased, log_phased, gl_diffs, n_p1, n_p2, read_sample, log_p1, log_p2,
               read_bp_diff, read_allele, column_sample, haploid=False, switches=None, gl_begin=None, gls=None, pls=None,
               pgl_begin=None, phased_gls=None):
    """ltr_vcf_record -> str.  read_bp_diff entries may be None; read_allele may be None.  switches (VCF_* mask) with the
    per-sample slices gl_begin / gls / pls / pgl_begin / phased_gls (ltr_batch_calls layout): ltr_vcf_record_ex."""
    lib = load()
    seq = np.frombuffer(chrom_seq.encode() if isinstance(chrom_seq, str) else bytes(chrom_seq), dtype=np.uint8)
    ab, aoff = pack_seqs(alleles)
    inex = np.array(list(inexact) + [0], dtype=np.uint8)
    kept = np.array(list(kept_mask) + [0], dtype=np.uint8)
    S = len(gts) // 2 if not hasattr(gts, "shape") else int(np.asarray(gts).size // 2)
    g = np.array(list(np.asarray(gts).ravel()) + [0], dtype=np.int32)
    lu = np.array(list(log_unphased) + [0.0], dtype=np.float64)
    lp = np.array(list(log_phased) + [0.0], dtype=np.float64)
    gd = np.array(list(gl_diffs) + [0.0], dtype=np.float64)
    p1c = np.array(list(n_p1) + [0], dtype=np.int32)
    p2c = np.array(list(n_p2) + [0], dtype=np.int32)
    rs = np.array(list(read_sample) + [0], dtype=np.int32)
    l1 = np.array(list(log_p1) + [0.0], dtype=np.float64)
    l2 = np.array(list(log_p2) + [0.0], dtype=np.float64)
    bd = np.array([INT32_MIN if x is None else x for x in read_bp_diff] + [0], dtype=np.int64).astype(np.int32)
    ra = None if read_allele is None else np.array(list(read_allele) + [0], dtype=np.int32)
    cs = np.array(list(column_sample) + [0], dtype=np.int32)
    L = VcfLocus(chrom.encode(), (name or "").encode(), motif.encode(), region_start, region_stop, ptr(seq, _u8p),
                 chrom_seq_start, len(seq), block_start, block_end, len(alleles), ptr(aoff, _u32p), ptr(ab, _u8p),
                 ptr(inex, _u8p), ptr(kept, _u8p), int(haploid), S, ptr(g, _i32p), ptr(lu, _dp), ptr(lp, _dp), ptr(gd, _dp),
                 ptr(p1c, _i32p), ptr(p2c, _i32p), len(read_sample), ptr(rs, _i32p), ptr(l1, _dp), ptr(l2, _dp), ptr(bd, _i32p),
                 None if ra is None else ptr(ra, _i32p), len(column_sample), ptr(cs, _i32p))
    cap = 1 << 16
    for _ in range(2):
        buf = C.create_string_buffer(cap)
        n = C.c_uint32(0)
        if switches is None:
            rc = lib.ltr_vcf_record(C.byref(L), buf, cap, C.byref(n))
        else:
            def arr(x, dt):
                return None if x is None else np.ascontiguousarray(list(x) + [0], dtype=dt)
            xa = [arr(gl_begin, np.uint64), arr(gls, np.float64), arr(pls, np.int32), arr(pgl_begin, np.uint64),
                  arr(phased_gls, np.float64)]
            pt = [_u64p, _dp, _i32p, _u64p, _dp]
            X = VcfExtras(int(switches), *[None if a is None else ptr(a, t) for a, t in zip(xa, pt)])
            rc = lib.ltr_vcf_record_ex(C.byref(L), C.byref(X), buf, cap, C.byref(n))
        if rc == 0:
            return buf.value.decode()
        if n.value + 1 > cap:
            cap = n.value + 16
            continue
        break
    raise RuntimeError("ltr_vcf_record failed: %d" % rc)


def vcf_header(fasta, fasta_path, command, sample_names, switches=None):
    """ltr_vcf_header (switches: ltr_vcf_header_ex) -> str (fasta: FastaFile)."""
    lib = load()
    names = (C.c_char_p * max(1, len(sample_names)))(*[x.encode() for x in sample_names])
    cap = 1 << 16
    for _ in range(2):
        buf = C.create_string_buffer(cap)
        n = C.c_uint32(0)
        if switches is None:
            rc = lib.ltr_vcf_header(fasta.h, fasta_path.encode(), command.encode(), names, len(sample_names), buf, cap, C.byref(n))
        else:
            rc = lib.ltr_vcf_header_ex(fasta.h, fasta_path.encode(), command.encode(), names, len(sample_names), int(switches),
                                       buf, cap, C.byref(n))
        if rc == 0:
            return buf.value.decode()
        cap = n.value + 16
    raise RuntimeError("ltr_vcf_header failed: %d" % rc)


class EmBatch(C.Structure):
    _fields_ = [("n_loci", C.c_uint32), ("locus_sample_begin", _u32p), ("sample_read_begin", _u32p), ("read_bp_diff", _i32p),
                ("log_p1", _dp), ("log_p2", _dp), ("locus_motif_len", _i32p), ("locus_haploid", _u8p)]


class EmOpts(C.Structure):
    _fields_ = [("max_iter", C.c_int32), ("abs_ll_converge", C.c_double), ("frac_ll_converge", C.c_double)]


def em_pack(loci):
    """[dict(reads_per_sample, bp_diff, log_p1, log_p2, motif_len, haploid=False)] (reads sample-major) -> the arrays of an
    ltr_em_batch (kept alive by the returned dict)."""
    n = len(loci)
    lsb = np.zeros(n + 1, dtype=np.uint32)
    srb, bd, p1, p2 = [0], [], [], []
    for k, L in enumerate(loci):
        for c in L["reads_per_sample"]:
            srb.append(srb[-1] + int(c))
        lsb[k + 1] = len(srb) - 1
        bd.extend(L["bp_diff"])
        p1.extend(L["log_p1"])
        p2.extend(L["log_p2"])
    P = dict(n=n, lsb=lsb, srb=np.array(srb, dtype=np.uint32), bd=np.array(bd + [0], dtype=np.int32),
             p1=np.array(p1 + [0.0], dtype=np.float64), p2=np.array(p2 + [0.0], dtype=np.float64),
             ml=np.array([L["motif_len"] for L in loci] + [1], dtype=np.int32),
             hp=np.array([1 if L.get("haploid") else 0 for L in loci] + [0], dtype=np.uint8))
    P["struct"] = EmBatch(n, ptr(P["lsb"], _u32p), ptr(P["srb"], _u32p), ptr(P["bd"], _i32p), ptr(P["p1"], _dp),
                          ptr(P["p2"], _dp), ptr(P["ml"], _i32p), ptr(P["hp"], _u8p))
    return P


def em_stutter_train(ctx, loci, max_iter=None, abs_conv=None, frac_conv=None, prior_stride=32):
    """ltr_em_stutter_train.  loci: a list as em_pack takes it, or what em_pack returned.
    -> dict(params [n, 6], trained, n_iter, ll, log_gt_priors [n, prior_stride])."""
    lib = load()
    P = loci if isinstance(loci, dict) else em_pack(loci)
    n = P["n"]
    O = EmOpts()
    lib.ltr_em_opts_default(C.byref(O))
    if max_iter is not None:
        O.max_iter = max_iter
    if abs_conv is not None:
        O.abs_ll_converge = abs_conv
    if frac_conv is not None:
        O.frac_ll_converge = frac_conv
    params = np.zeros((max(n, 1), 6))
    trained = np.zeros(max(n, 1), dtype=np.int32)
    n_iter = np.zeros(max(n, 1), dtype=np.int32)
    ll = np.zeros(max(n, 1))
    pri = np.zeros((max(n, 1), prior_stride))
    rc = lib.ltr_em_stutter_train(ctx, C.byref(P["struct"]), C.byref(O), ptr(params, _dp), ptr(trained, _i32p),
                                  ptr(n_iter, _i32p), ptr(ll, _dp), ptr(pri, _dp), prior_stride)
    if rc != 0:
        raise RuntimeError("ltr_em_stutter_train failed: %d" % rc)
    return dict(params=params[:n], trained=trained[:n], n_iter=n_iter[:n], ll=ll[:n], log_gt_priors=pri[:n])


def poa_consensus(seqs):
    """ltr_poa_consensus: partial-order consensus of the sequences, added in the given order."""
    lib = load()
    data, off = pack_seqs(seqs)
    cap = 2 * max([len(x) for x in seqs] + [1]) + 16
    out = np.zeros(cap, dtype=np.uint8)
    n = C.c_uint32(0)
    rc = lib.ltr_poa_consensus(ptr(data, _u8p), ptr(off, _u32p), len(seqs), ptr(out, _u8p), cap, C.byref(n))
    if rc != 0 and n.value > cap:
        cap = n.value
        out = np.zeros(cap, dtype=np.uint8)
        rc = lib.ltr_poa_consensus(ptr(data, _u8p), ptr(off, _u32p), len(seqs), ptr(out, _u8p), cap, C.byref(n))
    if rc != 0:
        raise RuntimeError("ltr_poa_consensus failed: %d" % rc)
    return out[:n.value].tobytes().decode()


_NIBBLE = np.full(256, 15, dtype=np.uint8)   # BAM's 4-bit codes; anything else becomes N
for _k, _c in enumerate("=ACMGRSVTWYHKDBN"):
    _NIBBLE[ord(_c)] = _k


def pack_reads_4bit(read_bytes, n_bases=None):
    """Bytes of a batch's reads (uint8 array, one byte per base) -> the 4-bit stream ltr_ctx_set_read_encoding(1) expects:
    base b in byte b // 2, even b in the high nibble."""
    a = np.asarray(read_bytes, dtype=np.uint8)
    if n_bases is not None:
        a = a[:n_bases]
    codes = _NIBBLE[a]
    if len(codes) % 2:
        codes = np.concatenate([codes, np.zeros(1, dtype=np.uint8)])
    return ((codes[0::2] << 4) | codes[1::2]).astype(np.uint8)


def pack_seqs(seqs):
    """Strings / bytes -> (seq_bytes uint8[], seq_off uint32[n+1])."""
    raw = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    off = np.zeros(len(raw) + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(r) for r in raw], dtype=np.uint64)
    data = np.frombuffer(b"".join(raw) + b"\0", dtype=np.uint8).copy()
    return data, off


def extract_calls(post, totals, haploid=False):
    """Host-only genotype extraction (Genotyper::extract_genotypes_and_likelihoods) on given posteriors."""
    lib = load()
    post = np.ascontiguousarray(post, dtype=np.float64)
    totals = np.ascontiguousarray(totals, dtype=np.float64)
    S, H = post.shape[0], post.shape[1]
    c, arrays = make_locus_calls(S, H, haploid)
    rc = lib.ltr_extract_calls(int(haploid), S, H, ptr(post, _dp), ptr(totals, _dp), C.byref(c))
    if rc != LTR_OK:
        raise RuntimeError("ltr_extract_calls failed: %d" % rc)
    arrays["total_ll"] = c.total_ll
    return arrays


def pool_reads(seqs, quals, lib=None, fn="ltr_pool_reads"):
    """ReadPooler on plain strings; returns (pool_index list, [median quality string per pool]).  ``lib`` / ``fn`` let the
    tests call the same signature in oracle/_ref (the reference's own ReadPooler)."""
    lib = lib or load()
    f = getattr(lib, fn)
    n = len(seqs)
    sp = (C.c_char_p * max(1, n))(*[s.encode() for s in seqs])
    qp = (C.c_char_p * max(1, n))(*[q.encode() for q in quals])
    pool = np.zeros(max(1, n), np.int32)
    npools = C.c_int32(0)
    cap = sum(len(s) for s in seqs) + 16
    buf = C.create_string_buffer(cap)
    f.argtypes = [C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), _i32p, C.POINTER(C.c_int32), C.c_char_p, C.c_int64]
    f.restype = C.c_int64
    got = f(n, sp, qp, ptr(pool, _i32p), C.byref(npools), buf, cap)
    if got < 0:
        raise RuntimeError("%s failed: %d" % (fn, got))
    raw = buf.raw[:got].decode()
    meds, off = [], 0
    first = {}
    for r in range(n):
        first.setdefault(int(pool[r]), r)
    for k in range(npools.value):
        ln = len(seqs[first[k]])
        meds.append(raw[off:off + ln])
        off += ln
    return [int(x) for x in pool[:n]], meds


def trim_read(locus, read_index):
    lib = load()
    n = len(locus.reads[read_index].seq)
    buf = C.create_string_buffer(n + 16)
    k = lib.ltr_trim_read_flat(C.byref(locus), read_index, buf, n + 16)
    if k < 0:
        raise RuntimeError("ltr_trim_read_flat failed: %d" % k)
    return buf.value[:k]


def seed_base(locus, read_index):
    return load().ltr_seed_base_flat(C.byref(locus), read_index)


class FlatBatch(C.Structure):
    """ltr_flat_batch (include/longtr_b200.h)."""
    _fields_ = [("vit", ViterbiBatch), ("hap_col", _i32p), ("read_row", _i32p), ("n_haps", C.c_uint32),
                ("n_reads", C.c_uint32)]


def flatten_loci(loci):
    """ltr_flatten_loci (host only): returns (batch dict as for make_viterbi_batch, hap_col, read_row, aln params)."""
    lib = load()
    lib.ltr_flatten_loci.argtypes = [C.c_int32, C.POINTER(FlatLocus), C.POINTER(Params), C.POINTER(C.POINTER(FlatBatch))]
    lib.ltr_flatten_loci.restype = C.c_int
    lib.ltr_flat_batch_free.argtypes = [C.POINTER(FlatBatch)]
    lib.ltr_flat_batch_free.restype = None
    n = len(loci)
    arr = (FlatLocus * max(1, n))(*loci)
    p = Params()
    h = C.POINTER(FlatBatch)()
    rc = lib.ltr_flatten_loci(n, arr, C.byref(p), C.byref(h))
    if rc != LTR_OK:
        raise RuntimeError("ltr_flatten_loci failed: %d" % rc)
    fb = h.contents
    as_arr = np.ctypeslib.as_array

    def take(ptr, count, dtype):
        return as_arr(ptr, (max(1, count),))[:count].astype(dtype).copy()
    nh, nr = fb.n_haps, fb.n_reads
    hap_off = take(fb.vit.hap_off, nh + 1, np.uint32)
    read_off = take(fb.vit.read_off, nr + 1, np.uint32)
    batch = dict(locus_hap_begin=take(fb.vit.locus_hap_begin, n + 1, np.uint32),
                 locus_read_begin=take(fb.vit.locus_read_begin, n + 1, np.uint32), hap_off=hap_off, read_off=read_off,
                 hap_bytes=take(fb.vit.hap_bytes, int(hap_off[-1]), np.uint8),
                 read_bytes=take(fb.vit.read_bytes, int(read_off[-1]), np.uint8))
    hap_col, read_row = take(fb.hap_col, nh, np.int32), take(fb.read_row, nr, np.int32)
    params = (p.ins_ins, p.ins_match, p.del_del, p.del_match, p.match_match, p.match_ins, p.match_del)
    flank = p.indel_flank_len
    lib.ltr_flat_batch_free(h)
    return batch, hap_col, read_row, params, flank


class StutterBatch(C.Structure):
    """ltr_stutter_batch (include/longtr_b200.h)."""
    _fields_ = [("n_loci", C.c_uint32), ("locus_allele_begin", _u32p), ("locus_read_begin", _u32p),
                ("lflank_off", _u32p), ("lflank_bytes", _u8p), ("rflank_off", _u32p), ("rflank_bytes", _u8p),
                ("allele_off", _u32p), ("allele_bytes", _u8p), ("stutter", _dp), ("motif_len", _i32p),
                ("read_off", _u32p), ("read_bytes", _u8p), ("qual_bytes", _u8p), ("read_seed", _i32p),
                ("realign_allele", _u8p), ("realign_read", _u8p)]


STUTTER_FIELDS = [("locus_allele_begin", np.uint32), ("locus_read_begin", np.uint32), ("lflank_off", np.uint32),
                  ("lflank_bytes", np.uint8), ("rflank_off", np.uint32), ("rflank_bytes", np.uint8),
                  ("allele_off", np.uint32), ("allele_bytes", np.uint8), ("stutter", np.float64),
                  ("motif_len", np.int32), ("read_off", np.uint32), ("read_bytes", np.uint8), ("qual_bytes", np.uint8),
                  ("read_seed", np.int32)]


def make_stutter_batch(batch):
    """dict of numpy arrays (keys of STUTTER_FIELDS) -> (StutterBatch, keepalive)."""
    keep = {k: np.ascontiguousarray(batch[k], dtype=dt) for k, dt in STUTTER_FIELDS}
    b = StutterBatch()
    b.n_loci = len(keep["locus_allele_begin"]) - 1
    for k, dt in STUTTER_FIELDS:
        t = {np.uint32: _u32p, np.uint8: _u8p, np.float64: _dp, np.int32: _i32p}[dt]
        setattr(b, k, ptr(keep[k], t))
    return b, keep


def stutter_ll_size(batch):
    a = np.asarray(batch["locus_allele_begin"], dtype=np.int64)
    r = np.asarray(batch["locus_read_begin"], dtype=np.int64)
    return int(np.sum((a[1:] - a[:-1]) * (r[1:] - r[:-1])))
