"""Synthetic workloads of BASELINE.json (configs 3-5).  Configs 3 and 4 come from the generator
library synth/libltr_synth.so (plain C++, no product code: the reference arm of bench.py maps only this), config 5 from
the product library (its pooled reads and seeds go through the host mirror).  Returns numpy views over library-owned
memory."""
import ctypes as C
import os

import numpy as np

from . import abi

BASE_SEEDS = {3: 20260103, 4: 20260104, 5: 20260105}


class SynthBatch(C.Structure):
    _fields_ = [("vit", abi.ViterbiBatch), ("post", abi.PosteriorBatch), ("n_haps", C.c_uint32),
                ("n_reads", C.c_uint32), ("n_sreads", C.c_uint32), ("hap_nbytes", C.c_uint64),
                ("read_nbytes", C.c_uint64)]


class Workload:
    """A generated batch: ``.batch`` / ``.post`` dicts (see abi.make_viterbi_batch)."""

    def __init__(self, lib, handle, config, first_locus):
        self._lib, self._h = lib, handle
        s = handle.contents
        n = s.vit.n_loci
        as_arr = np.ctypeslib.as_array
        self.config, self.first_locus, self.n_loci = config, first_locus, n
        self.batch = dict(
            locus_hap_begin=as_arr(s.vit.locus_hap_begin, (n + 1,)),
            locus_read_begin=as_arr(s.vit.locus_read_begin, (n + 1,)),
            hap_off=as_arr(s.vit.hap_off, (s.n_haps + 1,)),
            read_off=as_arr(s.vit.read_off, (s.n_reads + 1,)),
            hap_bytes=as_arr(s.vit.hap_bytes, (max(1, s.hap_nbytes),))[:s.hap_nbytes],
            read_bytes=as_arr(s.vit.read_bytes, (max(1, s.read_nbytes),))[:s.read_nbytes])
        ns = s.n_sreads
        self.post = dict(
            locus_sread_begin=as_arr(s.post.locus_sread_begin, (n + 1,)),
            pool_index=as_arr(s.post.pool_index, (max(1, ns),))[:ns],
            sample_label=as_arr(s.post.sample_label, (max(1, ns),))[:ns],
            log_p1=as_arr(s.post.log_p1, (max(1, ns),))[:ns],
            log_p2=as_arr(s.post.log_p2, (max(1, ns),))[:ns],
            locus_n_samples=as_arr(s.post.locus_n_samples, (max(1, n),))[:n],
            locus_haploid=None)
        p = abi.Params()
        lib.ltr_synth_params(config, C.byref(p))
        self.aln_params = (p.ins_ins, p.ins_match, p.del_del, p.del_match, p.match_match, p.match_ins,
                           p.match_del)
        self.input_bytes = int(sum(v.nbytes for v in self.batch.values()) +
                               sum(v.nbytes for v in self.post.values() if v is not None))

    def subset(self, n_loci):
        """First n_loci loci as independent (copied) dicts -- for bounded CPU-baseline samples."""
        b, p = self.batch, self.post
        nh, nr, ns = int(b["locus_hap_begin"][n_loci]), int(b["locus_read_begin"][n_loci]), \
            int(p["locus_sread_begin"][n_loci])
        sb = dict(locus_hap_begin=b["locus_hap_begin"][:n_loci + 1].copy(),
                  locus_read_begin=b["locus_read_begin"][:n_loci + 1].copy(),
                  hap_off=b["hap_off"][:nh + 1].copy(), read_off=b["read_off"][:nr + 1].copy(),
                  hap_bytes=b["hap_bytes"][:int(b["hap_off"][nh])].copy(),
                  read_bytes=b["read_bytes"][:int(b["read_off"][nr])].copy())
        sp = dict(locus_sread_begin=p["locus_sread_begin"][:n_loci + 1].copy(),
                  pool_index=p["pool_index"][:ns].copy(), sample_label=p["sample_label"][:ns].copy(),
                  log_p1=p["log_p1"][:ns].copy(), log_p2=p["log_p2"][:ns].copy(),
                  locus_n_samples=p["locus_n_samples"][:n_loci].copy(), locus_haploid=None)
        return sb, sp

    def close(self):
        if self._h is not None:
            self._lib.ltr_synth_free(self._h)
            self._h = None
            self.batch = self.post = None


_synth_lib = None


def synth_lib():
    global _synth_lib
    if _synth_lib is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "synth", "libltr_synth.so")
        if not os.path.exists(path):
            raise RuntimeError("libltr_synth.so is not built (run `python -m longtr_b200.build`)")
        _synth_lib = C.CDLL(path)
    return _synth_lib


def generate(config, n_loci, first_locus=0, base_seed=None, n_threads=None):
    lib = synth_lib()
    lib.ltr_synth_generate.argtypes = [C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int,
                                       C.POINTER(C.POINTER(SynthBatch))]
    lib.ltr_synth_generate.restype = C.c_int
    lib.ltr_synth_free.argtypes = [C.POINTER(SynthBatch)]
    lib.ltr_synth_params.argtypes = [C.c_int, C.POINTER(abi.Params)]
    if base_seed is None:
        base_seed = BASE_SEEDS[config]
    if n_threads is None:
        n_threads = min(32, os.cpu_count() or 1)
    h = C.POINTER(SynthBatch)()
    rc = lib.ltr_synth_generate(config, base_seed, first_locus, n_loci, n_threads, C.byref(h))
    if rc != 0:
        raise RuntimeError("ltr_synth_generate failed: %d" % rc)
    return Workload(lib, h, config, first_locus)


_SynthLoci = None


def _synth_loci_type():
    """ltr_synth_loci (synth/longtr_synth.h): the raw form of configs 3 / 4 for ltr_genotyper_run."""
    global _SynthLoci
    if _SynthLoci is None:
        from . import locus_batch

        class SynthLoci(C.Structure):
            _fields_ = [("batch", locus_batch.LocusBatch), ("n_alleles", C.c_uint32), ("n_reads", C.c_uint32),
                        ("n_cigar_ops", C.c_uint32), ("allele_nbytes", C.c_uint64), ("read_nbytes", C.c_uint64)]
        _SynthLoci = SynthLoci
    return _SynthLoci


class RawWorkload:
    """Raw loci (whole reads with CIGARs, flank blocks, candidate alleles) owned by the generator library; ``.struct`` is
    the ltr_locus_batch to hand to Genotyper.run_struct."""

    def __init__(self, lib, handle, config, n_loci):
        self._lib, self._h = lib, handle
        self.config, self.n_loci = config, n_loci
        s = handle.contents
        self.struct = s.batch
        self.n_reads, self.read_nbytes, self.n_alleles = s.n_reads, s.read_nbytes, s.n_alleles
        self.input_bytes = int(s.read_nbytes + s.allele_nbytes + 70 * n_loci + 4 * s.n_cigar_ops + 36 * s.n_reads)
        p = abi.Params()
        lib.ltr_synth_params(config, C.byref(p))
        self.aln_params = (p.ins_ins, p.ins_match, p.del_del, p.del_match, p.match_match, p.match_ins, p.match_del)

    def close(self):
        if self._h is not None:
            self._lib.ltr_synth_loci_free(self._h)
            self._h = None
            self.struct = None


def generate_loci(config, n_loci, first_locus=0, base_seed=None, n_threads=None):
    """ltr_synth_generate_loci: same seeds and read bases as ``generate``, in raw form."""
    lib = synth_lib()
    SynthLoci = _synth_loci_type()
    lib.ltr_synth_generate_loci.argtypes = [C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int,
                                            C.POINTER(C.POINTER(SynthLoci))]
    lib.ltr_synth_generate_loci.restype = C.c_int
    lib.ltr_synth_loci_free.argtypes = [C.POINTER(SynthLoci)]
    lib.ltr_synth_params.argtypes = [C.c_int, C.POINTER(abi.Params)]
    if base_seed is None:
        base_seed = BASE_SEEDS[config]
    if n_threads is None:
        n_threads = min(32, os.cpu_count() or 1)
    h = C.POINTER(SynthLoci)()
    rc = lib.ltr_synth_generate_loci(config, base_seed, first_locus, n_loci, n_threads, C.byref(h))
    if rc != 0:
        raise RuntimeError("ltr_synth_generate_loci failed: %d" % rc)
    return RawWorkload(lib, h, config, n_loci)


class SynthStutterBatch(C.Structure):
    _fields_ = [("batch", abi.StutterBatch), ("post", abi.PosteriorBatch), ("read_start", abi._i32p),
                ("read_stop", abi._i32p), ("cigar_off", abi._u32p), ("cigar_bytes", abi._u8p),
                ("repeat_start", abi._i32p), ("repeat_end", abi._i32p), ("n_alleles", C.c_uint32),
                ("n_reads", C.c_uint32), ("n_sreads", C.c_uint32)]


class StutterWorkload:
    """Config 5 (homopolymers, --stutter-align-len path): ``.batch`` is a dict for Engine.stutter_ll;
    ``flat_locus(l)`` rebuilds locus l as a FlatLocus (what HapAligner::process_reads is handed)."""

    def __init__(self, lib, handle, first_locus):
        self._lib, self._h = lib, handle
        s = handle.contents
        n, na, nr, ns = s.batch.n_loci, s.n_alleles, s.n_reads, s.n_sreads
        arr = np.ctypeslib.as_array
        self.config, self.first_locus, self.n_loci = 5, first_locus, n
        b = s.batch
        lfo, rfo = arr(b.lflank_off, (n + 1,)), arr(b.rflank_off, (n + 1,))
        alo, rdo = arr(b.allele_off, (na + 1,)), arr(b.read_off, (nr + 1,))
        self.batch = dict(
            locus_allele_begin=arr(b.locus_allele_begin, (n + 1,)), locus_read_begin=arr(b.locus_read_begin, (n + 1,)),
            lflank_off=lfo, lflank_bytes=arr(b.lflank_bytes, (int(lfo[-1]),)),
            rflank_off=rfo, rflank_bytes=arr(b.rflank_bytes, (int(rfo[-1]),)),
            allele_off=alo, allele_bytes=arr(b.allele_bytes, (int(alo[-1]),)),
            stutter=arr(b.stutter, (6 * n,)), motif_len=arr(b.motif_len, (n,)),
            read_off=rdo, read_bytes=arr(b.read_bytes, (int(rdo[-1]),)), qual_bytes=arr(b.qual_bytes, (int(rdo[-1]),)),
            read_seed=arr(b.read_seed, (nr,)))
        cgo = arr(s.cigar_off, (nr + 1,))
        self.extra = dict(read_start=arr(s.read_start, (nr,)), read_stop=arr(s.read_stop, (nr,)), cigar_off=cgo,
                          cigar_bytes=arr(s.cigar_bytes, (int(cgo[-1]),)), repeat_start=arr(s.repeat_start, (n,)),
                          repeat_end=arr(s.repeat_end, (n,)))
        self.post = dict(
            locus_sread_begin=arr(s.post.locus_sread_begin, (n + 1,)), pool_index=arr(s.post.pool_index, (ns,)),
            sample_label=arr(s.post.sample_label, (ns,)), log_p1=arr(s.post.log_p1, (ns,)),
            log_p2=arr(s.post.log_p2, (ns,)), locus_n_samples=arr(s.post.locus_n_samples, (n,)), locus_haploid=None)
        self.aln_params = abi.DEFAULT_ALN_PARAMS
        self.input_bytes = int(sum(v.nbytes for v in self.batch.values()))

    def flat_locus(self, l):
        from .flat import make_flat_locus
        b, x = self.batch, self.extra
        s = lambda buf, off, i: bytes(buf[int(off[i]):int(off[i + 1])])
        a0, a1 = int(b["locus_allele_begin"][l]), int(b["locus_allele_begin"][l + 1])
        r0, r1 = int(b["locus_read_begin"][l]), int(b["locus_read_begin"][l + 1])
        alleles = [s(b["allele_bytes"], b["allele_off"], a) for a in range(a0, a1)]
        reads = [(int(x["read_start"][r]), int(x["read_stop"][r]), s(b["read_bytes"], b["read_off"], r),
                  s(b["qual_bytes"], b["read_off"], r), s(x["cigar_bytes"], x["cigar_off"], r)) for r in range(r0, r1)]
        motif = alleles[0][5:6]
        return make_flat_locus(s(b["lflank_bytes"], b["lflank_off"], l), alleles, s(b["rflank_bytes"], b["rflank_off"], l),
                               int(x["repeat_start"][l]), int(x["repeat_end"][l]), 1, reads, motif=motif,
                               switch_old_align_len=20), (r1 - r0, a1 - a0)

    def cells(self, n_loci=None):
        """Cell-equivalents of the first n_loci loci."""
        b = self.batch
        n = self.n_loci if n_loci is None else n_loci
        tot = 0
        for l in range(n):
            a0, a1 = int(b["locus_allele_begin"][l]), int(b["locus_allele_begin"][l + 1])
            r0, r1 = int(b["locus_read_begin"][l]), int(b["locus_read_begin"][l + 1])
            B = np.diff(b["allele_off"][a0:a1 + 1]).astype(np.int64)
            N = np.diff(b["read_off"][r0:r1 + 1]).astype(np.int64)
            ok = b["read_seed"][r0:r1] >= 0
            cols = np.maximum(N[ok] - 1, 0)
            nf = int(b["lflank_off"][l + 1] - b["lflank_off"][l] + b["rflank_off"][l + 1] - b["rflank_off"][l])
            tot += int(np.sum(cols[:, None] * (nf + 13 * B[None, :])))
        return tot

    def close(self):
        if self._h is not None:
            self._lib.ltr_synth_stutter_free(self._h)
            self._h = None
            self.batch = self.post = self.extra = None


def generate_stutter(n_loci, first_locus=0, base_seed=None, n_threads=None):
    lib = abi.load()
    lib.ltr_synth_stutter_generate.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_int,
                                               C.POINTER(C.POINTER(SynthStutterBatch))]
    lib.ltr_synth_stutter_generate.restype = C.c_int
    lib.ltr_synth_stutter_free.argtypes = [C.POINTER(SynthStutterBatch)]
    if base_seed is None:
        base_seed = BASE_SEEDS[5]
    if n_threads is None:
        n_threads = min(32, os.cpu_count() or 1)
    h = C.POINTER(SynthStutterBatch)()
    rc = lib.ltr_synth_stutter_generate(base_seed, first_locus, n_loci, n_threads, C.byref(h))
    if rc != 0:
        raise RuntimeError("ltr_synth_stutter_generate failed: %d" % rc)
    return StutterWorkload(lib, h, first_locus)


# ---- N2: sets of distinct allele-like strings for the clustering kernels (numpy, seeded) --------------------------------
def generate_cluster_sets(n_sets, seed=20260106, seqs_per_set=40, lo=500, hi=1000):
    """The skipped sequences of n_sets (locus, sample) pairs, config-4-like: 2-4 repeat alleles of lo..hi bases that differ
    by whole motifs, each seen as distinct noisy copies (ONT-like 1 % substitutions, 1.5 % indels), ordered like the
    reference orders them before greedy_clustering (first string of the std::map in front, the rest by length and
    sequence: HaplotypeGenerator.cpp:399-401).  Returns (seq_bytes, seq_off, set_begin) with one copy of every string."""
    rng = np.random.default_rng(seed)
    code = np.frombuffer(b"ACGT", dtype=np.uint8)
    chunks, set_begin = [], [0]
    n_total = 0
    for _ in range(n_sets):
        motif = code[rng.integers(0, 4, size=int(rng.integers(10, 61)))]
        copies = int(rng.integers(lo, hi + 1)) // len(motif)
        strings = set()
        n_alleles = int(rng.integers(2, 5))
        for a in range(n_alleles):
            allele = np.tile(motif, max(2, copies + int(rng.integers(-4, 5))))
            for _ in range(max(2, seqs_per_set // n_alleles)):
                s = allele.copy()
                sub = rng.random(len(s)) < 0.01
                s[sub] = code[rng.integers(0, 4, size=int(sub.sum()))]
                keep = rng.random(len(s)) >= 0.0075
                s = s[keep]
                ins = np.flatnonzero(rng.random(len(s)) < 0.0075)
                s = np.insert(s, ins, code[rng.integers(0, 4, size=len(ins))])
                strings.add(s.tobytes())
        strings = sorted(strings)
        ordered = [strings[0]] + sorted(strings[1:], key=lambda x: (len(x), x))
        chunks += ordered
        n_total += len(ordered)
        set_begin.append(n_total)
    seq_off = np.zeros(n_total + 1, dtype=np.uint32)
    seq_off[1:] = np.cumsum([len(c) for c in chunks], dtype=np.uint64)
    seq_bytes = np.frombuffer(b"".join(chunks) + b"\0", dtype=np.uint8).copy()
    return seq_bytes, seq_off, np.array(set_begin, dtype=np.uint32)
