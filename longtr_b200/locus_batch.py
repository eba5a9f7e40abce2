"""Python side of ltr_genotyper_* (include/longtr_b200.h): many raw loci -> genotype calls.

``build_locus_batch`` turns a list of plain-Python loci into the structure-of-arrays ``ltr_locus_batch``;
``Genotyper`` owns an ``ltr_genotyper`` (one context per device + host worker threads).  ctypes only: nothing here
computes on the CPU beyond packing arrays, and there is no fallback without a GPU."""
import ctypes as C
import re

import numpy as np

from . import abi

_u32p, _u8p, _i32p, _dp = abi._u32p, abi._u8p, abi._i32p, abi._dp
_u64p = C.POINTER(C.c_uint64)


class LocusBatch(C.Structure):
    _fields_ = [("n_loci", C.c_uint32), ("lflank_off", _u32p), ("lflank_bytes", _u8p), ("rflank_off", _u32p),
                ("rflank_bytes", _u8p), ("locus_allele_begin", _u32p), ("allele_off", _u32p), ("allele_bytes", _u8p),
                ("repeat_start", _i32p), ("repeat_end", _i32p), ("locus_read_begin", _u32p), ("read_start", _i32p),
                ("read_stop", _i32p), ("read_off", _u32p), ("read_bytes", _u8p), ("cigar_off", _u32p),
                ("cigar_ops", _u32p), ("read_sample", _i32p), ("log_p1", _dp), ("log_p2", _dp), ("second_mate", _u8p),
                ("locus_n_samples", _u32p), ("locus_haploid", _u8p)]


class BatchCalls(C.Structure):
    _fields_ = [("n_loci", C.c_uint32), ("status", _i32p), ("locus_sample_begin", _u32p), ("locus_allele_begin", _u32p),
                ("kept_mask", _u8p), ("n_kept", _i32p), ("n_pools", _i32p), ("gts", _i32p),
                ("log_phased_posteriors", _dp), ("log_unphased_posteriors", _dp), ("gl_diffs", _dp),
                ("sample_total_lls", _dp), ("n_reads", _i32p), ("gl_begin", _u64p), ("gls", _dp), ("pls", _i32p),
                ("prep_ms", C.c_double), ("gpu_wait_ms", C.c_double), ("post_ms", C.c_double), ("total_ms", C.c_double),
                ("submit_ms", C.c_double), ("n_chunks", C.c_uint32), ("read_allele", _i32p), ("pgl_begin", _u64p),
                ("phased_gls", _dp)]


BAM_OPS = "MIDNSHP=X"
_CIGAR_RE = re.compile(r"(\d+)(.)")


class Region(C.Structure):
    _fields_ = [("start", C.c_int32), ("stop", C.c_int32), ("period", C.c_int32)]


class RegionsOpts(C.Structure):
    _fields_ = [("host_threads", C.c_int32), ("max_tr_len", C.c_int32), ("min_total_reads", C.c_int32),
                ("no_assembly", C.c_int32), ("vcf_records", C.c_int32), ("region_names", C.POINTER(C.c_char_p)),
                ("region_motifs", C.POINTER(C.c_char_p)), ("haploid", C.c_int32), ("vcf_switches", C.c_uint32)]


class RegionsResult(C.Structure):
    _fields_ = [("n_regions", C.c_uint32), ("status", _i32p), ("locus_index", _i32p), ("n_loci", C.c_uint32),
                ("calls", C.POINTER(BatchCalls)), ("block_start", _i32p), ("block_end", _i32p),
                ("region_allele_begin", _u32p), ("allele_off", _u32p), ("allele_bytes", _u8p),
                ("region_sample_begin", _u32p), ("sample_file", _u32p), ("allele_inexact", _u8p),
                ("n_assembled", C.c_uint32), ("record_off", _u32p), ("records", C.c_void_p), ("owner", C.c_void_p),
                ("prepare_ms", C.c_double), ("layout_ms", C.c_double), ("genotype_ms", C.c_double), ("records_ms", C.c_double)]


REGIONS_SINK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(RegionsResult))   # ltr_regions_sink


class BedRunResult(C.Structure):
    _fields_ = [("n_chroms", C.c_uint32), ("per_chrom", C.POINTER(C.POINTER(RegionsResult))), ("chrom_region_begin", _u32p),
                ("owner", C.c_void_p)]


def pack_cigar(cigar):
    """'110=4I90=' -> BAM-encoded uint32 list (length << 4 | op).  Unknown operations get code 15 (rejected by the library
    like the reference rejects them)."""
    out = []
    for n, op in _CIGAR_RE.findall(cigar):
        k = BAM_OPS.find(op)
        out.append((int(n) << 4) | (k if k >= 0 else 15))
    return out


def build_locus_batch(loci):
    """loci: list of dicts with lflank, rflank, alleles (list of str), repeat_start, repeat_end, n_samples, haploid (opt),
    reads: list of dicts start, stop, seq, cigar (str), sample, log_p1, log_p2, second_mate (opt) -- sample-major.
    Returns a dict of numpy arrays (the fields of ltr_locus_batch)."""
    def offs(strs):
        return np.concatenate([[0], np.cumsum([len(s) for s in strs])]).astype(np.uint32)

    def cat(strs):
        return np.frombuffer("".join(strs).encode(), dtype=np.uint8).copy()
    lf, rf = [l["lflank"] for l in loci], [l["rflank"] for l in loci]
    alleles = [a for l in loci for a in l["alleles"]]
    reads = [r for l in loci for r in l["reads"]]
    cig = [pack_cigar(r["cigar"]) for r in reads]
    any_mate = any(r.get("second_mate") for r in reads)
    b = dict(
        lflank_off=offs(lf), lflank_bytes=cat(lf), rflank_off=offs(rf), rflank_bytes=cat(rf),
        locus_allele_begin=np.concatenate([[0], np.cumsum([len(l["alleles"]) for l in loci])]).astype(np.uint32),
        allele_off=offs(alleles), allele_bytes=cat(alleles),
        repeat_start=np.array([l["repeat_start"] for l in loci], np.int32),
        repeat_end=np.array([l["repeat_end"] for l in loci], np.int32),
        locus_read_begin=np.concatenate([[0], np.cumsum([len(l["reads"]) for l in loci])]).astype(np.uint32),
        read_start=np.array([r["start"] for r in reads], np.int32), read_stop=np.array([r["stop"] for r in reads], np.int32),
        read_off=offs([r["seq"] for r in reads]), read_bytes=cat([r["seq"] for r in reads]),
        cigar_off=np.concatenate([[0], np.cumsum([len(c) for c in cig])]).astype(np.uint32),
        cigar_ops=np.array([x for c in cig for x in c], np.uint32),
        read_sample=np.array([r["sample"] for r in reads], np.int32),
        log_p1=np.array([r["log_p1"] for r in reads], np.float64), log_p2=np.array([r["log_p2"] for r in reads], np.float64),
        second_mate=(np.array([1 if r.get("second_mate") else 0 for r in reads], np.uint8) if any_mate else None),
        locus_n_samples=np.array([l["n_samples"] for l in loci], np.uint32),
        locus_haploid=np.array([1 if l.get("haploid") else 0 for l in loci], np.uint8))
    return b


_PTR = dict(lflank_off=_u32p, lflank_bytes=_u8p, rflank_off=_u32p, rflank_bytes=_u8p, locus_allele_begin=_u32p,
            allele_off=_u32p, allele_bytes=_u8p, repeat_start=_i32p, repeat_end=_i32p, locus_read_begin=_u32p,
            read_start=_i32p, read_stop=_i32p, read_off=_u32p, read_bytes=_u8p, cigar_off=_u32p, cigar_ops=_u32p,
            read_sample=_i32p, log_p1=_dp, log_p2=_dp, second_mate=_u8p, locus_n_samples=_u32p, locus_haploid=_u8p)
_DT = {_u32p: np.uint32, _u8p: np.uint8, _i32p: np.int32, _dp: np.float64}


def make_locus_batch(b):
    """dict of numpy arrays -> (LocusBatch, keepalive)."""
    s = LocusBatch()
    keep = {}
    s.n_loci = len(b["locus_read_begin"]) - 1
    for k, t in _PTR.items():
        v = b.get(k)
        if v is None:
            continue
        a = np.ascontiguousarray(v, dtype=_DT[t])
        if a.size == 0:
            a = np.zeros(1, dtype=_DT[t])
        keep[k] = a
        setattr(s, k, a.ctypes.data_as(t))
    return s, keep


def _declare(lib):
    vp = C.c_void_p
    lib.ltr_genotyper_create.argtypes = [_i32p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(vp)]
    lib.ltr_genotyper_create.restype = C.c_int
    lib.ltr_genotyper_destroy.argtypes = [vp]
    lib.ltr_genotyper_destroy.restype = None
    lib.ltr_genotyper_set_read_alleles.argtypes = [vp, C.c_int32]
    lib.ltr_genotyper_set_read_alleles.restype = C.c_int
    lib.ltr_genotyper_set_phased_gls.argtypes = [vp, C.c_int32]
    lib.ltr_genotyper_set_phased_gls.restype = C.c_int
    lib.ltr_genotyper_run.argtypes = [vp, C.POINTER(abi.Params), C.POINTER(LocusBatch), C.POINTER(C.POINTER(BatchCalls))]
    lib.ltr_genotyper_run.restype = C.c_int
    lib.ltr_batch_calls_free.argtypes = [C.POINTER(BatchCalls)]
    lib.ltr_batch_calls_free.restype = None
    lib.ltr_locus_batch_trim_read.argtypes = [C.POINTER(LocusBatch), C.POINTER(abi.Params), C.c_uint32, C.c_uint32, _u8p,
                                              C.c_int32]
    lib.ltr_locus_batch_trim_read.restype = C.c_int32
    lib.ltr_regions_run.argtypes = [vp, C.POINTER(abi.Params), C.POINTER(C.c_void_p), C.c_int32, C.c_char_p,
                                    C.POINTER(Region), C.c_uint32, _u8p, C.c_int64, C.c_int64, C.POINTER(abi.RegionParams),
                                    C.POINTER(RegionsOpts), C.POINTER(C.POINTER(RegionsResult))]
    lib.ltr_regions_run.restype = C.c_int
    lib.ltr_regions_result_free.argtypes = [C.POINTER(RegionsResult)]
    lib.ltr_regions_result_free.restype = None
    lib.ltr_run_bed.argtypes = [vp, C.POINTER(abi.Params), C.POINTER(C.c_void_p), C.c_int32, vp, C.POINTER(abi.Bed),
                                C.POINTER(abi.RegionParams), C.POINTER(RegionsOpts), C.POINTER(C.POINTER(BedRunResult))]
    lib.ltr_run_bed.restype = C.c_int
    lib.ltr_run_bed_stream.argtypes = [vp, C.POINTER(abi.Params), C.POINTER(C.c_void_p), C.c_int32, vp, C.POINTER(abi.Bed),
                                       C.POINTER(abi.RegionParams), C.POINTER(RegionsOpts), C.c_int32, REGIONS_SINK, vp]
    lib.ltr_run_bed_stream.restype = C.c_int
    lib.ltr_bed_run_result_free.argtypes = [C.POINTER(BedRunResult)]
    lib.ltr_bed_run_result_free.restype = None


def trim_read(batch_struct, locus, read, aln_params=None, indel_flank_len=5, cap=1 << 16):
    """ltr_locus_batch_trim_read (host only)."""
    lib = abi.load()
    _declare(lib)
    p = abi.make_params(aln_params, indel_flank_len)
    buf = np.zeros(cap, np.uint8)
    n = lib.ltr_locus_batch_trim_read(C.byref(batch_struct), C.byref(p), locus, read, buf.ctypes.data_as(_u8p), cap)
    if n < 0:
        raise RuntimeError("ltr_locus_batch_trim_read failed: %d" % n)
    return bytes(buf[:n])


class Genotyper:
    """``ltr_genotyper``: contexts on the given devices + host worker threads."""

    def __init__(self, devices=(0,), host_threads=0, chunk_loci=0):
        from .engine import LongTRError
        self.lib = abi.load()
        _declare(self.lib)
        self.h = C.c_void_p()
        dev = np.array(list(devices), np.int32)
        rc = self.lib.ltr_genotyper_create(dev.ctypes.data_as(_i32p), len(dev), host_threads, chunk_loci, C.byref(self.h))
        if rc != abi.LTR_OK:
            self.h = None
            raise LongTRError("ltr_genotyper_create: %s" % self.lib.ltr_strerror(rc).decode())

    def run_struct(self, batch_struct, aln_params=None, indel_flank_len=5):
        """Runs on a prepared LocusBatch; returns the raw BatchCalls pointer (free with ``free``)."""
        from .engine import LongTRError
        p = abi.make_params(aln_params, indel_flank_len)
        out = C.POINTER(BatchCalls)()
        rc = self.lib.ltr_genotyper_run(self.h, C.byref(p), C.byref(batch_struct), C.byref(out))
        if rc != abi.LTR_OK:
            raise LongTRError("ltr_genotyper_run: %s" % self.lib.ltr_strerror(rc).decode())
        return out

    def free(self, calls):
        self.lib.ltr_batch_calls_free(calls)

    def run(self, batch, aln_params=None, indel_flank_len=5):
        """batch: dict of numpy arrays (build_locus_batch).  Returns a dict of numpy copies of ltr_batch_calls."""
        s, keep = make_locus_batch(batch)
        calls = self.run_struct(s, aln_params, indel_flank_len)
        out = self._calls_dict(calls, int(np.asarray(batch["locus_read_begin"])[-1]))
        self.free(calls)
        return out

    def set_read_alleles(self, on=True):
        """ltr_genotyper_set_read_alleles: later runs also return read_allele (what MALLREADS counts)."""
        self.lib.ltr_genotyper_set_read_alleles(self.h, 1 if on else 0)

    def set_phased_gls(self, on=True):
        """ltr_genotyper_set_phased_gls: later runs also return pgl_begin / phased_gls (PHASEDGL)."""
        self.lib.ltr_genotyper_set_phased_gls(self.h, 1 if on else 0)

    @staticmethod
    def _calls_dict(calls, nr=0):
        c = calls.contents
        n = c.n_loci
        arr = np.ctypeslib.as_array

        def take(ptr, count):
            return arr(ptr, (max(1, count),))[:count].copy()
        lsb = take(c.locus_sample_begin, n + 1)
        lab = take(c.locus_allele_begin, n + 1)
        ns, na = int(lsb[-1]), int(lab[-1])
        glb = take(c.gl_begin, ns + 1)
        out = dict(status=take(c.status, n), locus_sample_begin=lsb, locus_allele_begin=lab, kept_mask=take(c.kept_mask, na),
                   n_kept=take(c.n_kept, n), n_pools=take(c.n_pools, n), gts=take(c.gts, 2 * ns).reshape(ns, 2),
                   log_phased_posteriors=take(c.log_phased_posteriors, ns),
                   log_unphased_posteriors=take(c.log_unphased_posteriors, ns), gl_diffs=take(c.gl_diffs, ns),
                   sample_total_lls=take(c.sample_total_lls, ns), n_reads=take(c.n_reads, ns), gl_begin=glb,
                   gls=take(c.gls, int(glb[-1])), pls=take(c.pls, int(glb[-1])),
                   read_allele=(None if not c.read_allele or not nr else take(c.read_allele, nr)),
                   timing=dict(prep_ms=c.prep_ms, gpu_wait_ms=c.gpu_wait_ms, post_ms=c.post_ms, total_ms=c.total_ms, submit_ms=c.submit_ms,
                               n_chunks=c.n_chunks))
        if c.pgl_begin:
            out["pgl_begin"] = take(c.pgl_begin, ns + 1)
            out["phased_gls"] = take(c.phased_gls, int(out["pgl_begin"][-1]))
        return out

    def _regions_dict(self, r):
        n = r.n_regions
        ab = C.string_at(r.allele_bytes, r.allele_off[r.region_allele_begin[n]]) if n and r.region_allele_begin[n] else b""
        return dict(status=[r.status[i] for i in range(n)], locus_index=[r.locus_index[i] for i in range(n)],
                   block=[(r.block_start[i], r.block_end[i]) for i in range(n)],
                   alleles=[[ab[r.allele_off[a]:r.allele_off[a + 1]].decode()
                             for a in range(r.region_allele_begin[i], r.region_allele_begin[i + 1])] for i in range(n)],
                   samples=[[r.sample_file[k] for k in range(r.region_sample_begin[i], r.region_sample_begin[i + 1])]
                            for i in range(n)],
                   inexact=[[int(r.allele_inexact[a]) for a in range(r.region_allele_begin[i], r.region_allele_begin[i + 1])]
                            for i in range(n)],
                   n_assembled=r.n_assembled,
                   records=(None if not r.record_off else
                            [C.string_at(r.records + r.record_off[i], r.record_off[i + 1] - r.record_off[i]).decode()
                             for i in range(n)]),
                   calls=self._calls_dict(r.calls) if r.n_loci else None,
                   host_ms=dict(prepare_ms=r.prepare_ms, layout_ms=r.layout_ms, genotype_ms=r.genotype_ms, records_ms=r.records_ms))

    def run_regions(self, bams, chrom, regions, ref_seq, ref_seq_start=0, aln_params=None, indel_flank_len=5,
                    host_threads=0, max_tr_len=1000, min_total_reads=10, no_assembly=0, motifs=None, names=None,
                    haploid=False, vcf_switches=None, **region_overrides):
        """ltr_regions_run: bams = [abi.BamFile], regions = [(start, stop, period)] on `chrom`.  Returns dict(status,
        locus_index, alleles [per region], block [(start, end)], samples [per region: file indices], calls (as ``run``)).
        motifs = [str per region] (+ names): also the VCF record of every genotyped region (``records``)."""
        from .engine import LongTRError
        lib = self.lib
        prm = abi.make_params(aln_params, indel_flank_len)
        rp = abi.RegionParams()
        lib.ltr_region_params_default(C.byref(rp))
        for k, v in region_overrides.items():
            setattr(rp, k, v)
        opts = RegionsOpts(host_threads, max_tr_len, min_total_reads, no_assembly)
        opts.haploid = 1 if haploid else 0
        opts.vcf_switches = abi.VCF_DEFAULT if vcf_switches is None else int(vcf_switches)   # ltr_regions_opts_default
        if motifs is not None:
            m_arr = (C.c_char_p * max(1, len(regions)))(*[m.encode() for m in motifs])
            n_arr = (C.c_char_p * max(1, len(regions)))(*[(x or "").encode() for x in (names or [""] * len(regions))])
            opts.vcf_records = 1
            opts.region_motifs = m_arr
            opts.region_names = n_arr
        regs = (Region * max(1, len(regions)))(*[Region(*r) for r in regions])
        handles = (C.c_void_p * len(bams))(*[b.h for b in bams])
        ref = np.frombuffer(ref_seq.encode() if isinstance(ref_seq, str) else bytes(ref_seq), dtype=np.uint8)
        out = C.POINTER(RegionsResult)()
        import time
        t0 = time.perf_counter()
        rc = lib.ltr_regions_run(self.h, C.byref(prm), handles, len(bams), chrom.encode(), regs, len(regions),
                                 ref.ctypes.data_as(_u8p), ref_seq_start, len(ref), C.byref(rp), C.byref(opts), C.byref(out))
        call_ms = (time.perf_counter() - t0) * 1e3
        if rc != abi.LTR_OK:
            raise LongTRError("ltr_regions_run: %s" % lib.ltr_strerror(rc).decode())
        res = self._regions_dict(out.contents)
        res["call_ms"] = call_ms   # the ltr_regions_run call alone (the conversion into Python objects is harness time)
        lib.ltr_regions_result_free(out)
        return res

    def run_bed(self, bams, fasta, bed_path, aln_params=None, indel_flank_len=5, host_threads=0, max_tr_len=1000,
                min_total_reads=10, no_assembly=0, chrom_limit=None, vcf_records=False, vcf_switches=None, **region_overrides):
        """ltr_run_bed: bams = [abi.BamFile], fasta = abi.FastaFile, bed_path = region file (CHROM START STOP MOTIF [NAME]).
        Returns dict(chroms, bed=[(chrom index, start, stop, period, name, motif)], per_chrom=[as run_regions])."""
        from .engine import LongTRError
        lib = self.lib
        bed = abi.bed_read(bed_path, 0, chrom_limit, keep_handle=True)
        prm = abi.make_params(aln_params, indel_flank_len)
        rp = abi.RegionParams()
        lib.ltr_region_params_default(C.byref(rp))
        for k, v in region_overrides.items():
            setattr(rp, k, v)
        opts = RegionsOpts(host_threads, max_tr_len, min_total_reads, no_assembly)
        opts.vcf_records = 1 if vcf_records else 0
        opts.vcf_switches = abi.VCF_DEFAULT if vcf_switches is None else int(vcf_switches)
        handles = (C.c_void_p * len(bams))(*[b.h for b in bams])
        out = C.POINTER(BedRunResult)()
        rc = lib.ltr_run_bed(self.h, C.byref(prm), handles, len(bams), fasta.h, bed["handle"], C.byref(rp), C.byref(opts),
                             C.byref(out))
        if rc != abi.LTR_OK:
            lib.ltr_bed_free(bed["handle"])
            raise LongTRError("ltr_run_bed: %s" % lib.ltr_strerror(rc).decode())
        r = out.contents
        res = dict(chroms=bed["chroms"], bed=bed["regions"],
                   chrom_region_begin=[r.chrom_region_begin[c] for c in range(r.n_chroms + 1)],
                   per_chrom=[self._regions_dict(r.per_chrom[c].contents) for c in range(r.n_chroms)])
        lib.ltr_bed_run_result_free(out)
        lib.ltr_bed_free(bed["handle"])
        return res

    def run_bed_stream(self, bams, fasta, bed_path, on_chunk, chunk_regions=0, aln_params=None, indel_flank_len=5, host_threads=0,
                       max_tr_len=1000, min_total_reads=10, no_assembly=0, chrom_limit=None, vcf_records=False, vcf_switches=None,
                       **region_overrides):
        """ltr_run_bed_stream: as run_bed, but the regions go through in chunks and on_chunk(chrom index, first region,
        result dict as run_regions') is called per chunk (a true return value stops the run)."""
        from .engine import LongTRError
        lib = self.lib
        bed = abi.bed_read(bed_path, 0, chrom_limit, keep_handle=True)
        prm = abi.make_params(aln_params, indel_flank_len)
        rp = abi.RegionParams()
        lib.ltr_region_params_default(C.byref(rp))
        for k, v in region_overrides.items():
            setattr(rp, k, v)
        opts = RegionsOpts(host_threads, max_tr_len, min_total_reads, no_assembly)
        opts.vcf_records = 1 if vcf_records else 0
        opts.vcf_switches = abi.VCF_DEFAULT if vcf_switches is None else int(vcf_switches)
        handles = (C.c_void_p * len(bams))(*[b.h for b in bams])

        def sink(_user, chrom, first, res):
            return 1 if on_chunk(chrom, first, self._regions_dict(res.contents)) else 0
        cb = REGIONS_SINK(sink)
        rc = lib.ltr_run_bed_stream(self.h, C.byref(prm), handles, len(bams), fasta.h, bed["handle"], C.byref(rp), C.byref(opts),
                                    int(chunk_regions), cb, None)
        lib.ltr_bed_free(bed["handle"])
        if rc != abi.LTR_OK:
            raise LongTRError("ltr_run_bed_stream: %s" % lib.ltr_strerror(rc).decode())
        return dict(chroms=bed["chroms"], bed=bed["regions"])

    def close(self):
        if self.h:
            self.lib.ltr_genotyper_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
