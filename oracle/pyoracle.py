"""TEST INFRASTRUCTURE ONLY -- ctypes access to the two CPU checkers.

* ``liblongtr_oracle.so``  : plain-C restatement (oracle/longtr_oracle.c)
* ``_ref/libltr_ref.so``   : the unmodified reference sources compiled in place
                             (oracle/build_ref.sh), when available.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from longtr_b200.flat import FlatLocus

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "liblongtr_oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libltr_ref.so")


class OracleParams(C.Structure):
    _fields_ = [("ins_ins", C.c_float), ("ins_match", C.c_float), ("del_del", C.c_float),
                ("del_match", C.c_float), ("match_match", C.c_float), ("match_ins", C.c_float),
                ("match_del", C.c_float), ("indel_flank_len", C.c_int32)]


def build(force=False):
    """Compile the restatement (and oracle/_ref when /root/reference exists)."""
    if force or not os.path.exists(_ORACLE_SO) or \
            os.path.getmtime(_ORACLE_SO) < max(os.path.getmtime(os.path.join(_HERE, f)) for f in
                                              ("longtr_oracle.c", "longtr_oracle_short.c", "longtr_oracle_edit.c", "longtr_oracle.h")):
        subprocess.check_call(["make", "-C", _HERE, "liblongtr_oracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir(os.environ.get("LONGTR_REFERENCE", "/root/reference") + "/src"):
        outs = [os.path.join(_HERE, "_ref", f) for f in
                ("libltr_ref.so", "libltr_ref_io.so", "libltr_ref_hapgen.so", "libltr_ref_hapgen_poa.so", "libltr_ref_em.so", "libltr_ref_fasta.so",
                 "ltr_ref_full", "ltr_ref_trace", "ltr_ref_lazy")]
        srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp", ".sh"))]
        srcs += [os.path.join(_HERE, "shim", "spoa", "spoa.hpp"), os.path.join(_HERE, "shim", "hts_stubs.cpp")]
        newest = max(os.path.getmtime(f) for f in srcs if os.path.exists(f))
        if force or any(not os.path.exists(o) or os.path.getmtime(o) < newest for o in outs):
            subprocess.check_call([os.path.join(_HERE, "build_ref.sh")], stdout=subprocess.DEVNULL)


_oracle = None
_ref = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)


def oracle_lib():
    global _oracle
    if _oracle is None:
        build()
        lib = C.CDLL(_ORACLE_SO)
        lib.ltr_oracle_default_params.argtypes = [C.POINTER(OracleParams)]
        lib.ltr_oracle_viterbi_pair.restype = C.c_double
        lib.ltr_oracle_viterbi_pair.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32,
                                                C.POINTER(OracleParams)]
        lib.ltr_oracle_viterbi_pair_cells.restype = C.c_double
        lib.ltr_oracle_viterbi_pair_cells.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32,
                                                      C.POINTER(OracleParams), C.POINTER(C.c_int64)]
        lib.ltr_oracle_trim_read.restype = C.c_int32
        lib.ltr_oracle_trim_read.argtypes = [C.POINTER(FlatLocus), C.c_int32, C.c_char_p]
        lib.ltr_oracle_process_reads.restype = C.c_int
        lib.ltr_oracle_process_reads.argtypes = [C.POINTER(FlatLocus), _dp, _ip]
        lib.ltr_oracle_viterbi_batch.restype = C.c_int
        lib.ltr_oracle_viterbi_batch.argtypes = [C.c_uint32, _u32p, _u32p, _u32p, _u8p, _u32p, _u8p,
                                                 C.POINTER(OracleParams), _dp, C.POINTER(C.c_int64),
                                                 C.c_int]
        lib.ltr_oracle_log_sample_posteriors.restype = C.c_double
        lib.ltr_oracle_log_sample_posteriors.argtypes = [C.c_int, C.c_int32, C.c_int32, C.c_int32, _dp,
                                                         _dp, _dp, _ip, _dp, _dp]
        lib.ltr_oracle_optimal_haplotypes.argtypes = [C.c_int32, C.c_int32, _dp, _ip]
        lib.ltr_oracle_edit_score.restype = C.c_int32
        lib.ltr_oracle_edit_score.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.c_int32]
        lib.ltr_oracle_greedy_cluster.restype = C.c_int32
        lib.ltr_oracle_greedy_cluster.argtypes = [_u8p, _u32p, _u32p, C.c_int32, C.c_int32, _ip, _ip]
        _oracle = lib
    return _oracle


def ref_available():
    build()
    return os.path.exists(_REF_SO)


def ref_lib():
    global _ref
    if _ref is None:
        build()
        lib = C.CDLL(_REF_SO)
        lib.ltr_ref_process_reads.restype = C.c_int
        lib.ltr_ref_process_reads.argtypes = [C.POINTER(FlatLocus), _dp, _ip, _dp]
        lib.ltr_ref_log_sample_posteriors.restype = C.c_double
        lib.ltr_ref_log_sample_posteriors.argtypes = [C.c_int, C.c_int, _ip, C.c_int, _dp, _dp, _dp,
                                                      _dp, _dp, _dp, _ip]
        lib.ltr_ref_viterbi_batch.restype = C.c_int
        lib.ltr_ref_viterbi_batch.argtypes = [C.c_uint32, _u32p, _u32p, _u32p, _u8p, _u32p, _u8p, C.c_int,
                                              C.POINTER(C.c_float), C.c_int, _dp, _dp,
                                              _u32p, _u32p, _dp, _dp, _dp]
        from longtr_b200.abi import LocusCalls
        lib.ltr_ref_genotype_locus.restype = C.c_int
        lib.ltr_ref_genotype_locus.argtypes = [C.c_int, C.c_int, _ip, C.c_int, _dp, _dp, _dp, _dp, C.POINTER(LocusCalls)]
        lib.ltr_ref_seed_bases.restype = C.c_int
        lib.ltr_ref_seed_bases.argtypes = [C.POINTER(FlatLocus), _ip]
        lib.ltr_ref_edit_score.restype = C.c_int32
        lib.ltr_ref_edit_score.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.c_int32]
        lib.ltr_ref_greedy_cluster.restype = C.c_int32
        lib.ltr_ref_greedy_cluster.argtypes = [_u8p, _u32p, _u32p, C.c_int32, C.c_int32, _ip, _ip]
        _ref = lib
    return _ref


def ref_genotype_locus(ll, log_p1, log_p2, reads_per_sample, haploid=False):
    """Reference calc_log_sample_posteriors + extract_genotypes_and_likelihoods (hap_to_allele = identity)."""
    from longtr_b200.abi import make_locus_calls
    ll = np.ascontiguousarray(ll, dtype=np.float64)
    R, H = ll.shape
    rps = np.ascontiguousarray(reads_per_sample, dtype=np.int32)
    p1 = np.ascontiguousarray(log_p1, dtype=np.float64)
    p2 = np.ascontiguousarray(log_p2, dtype=np.float64)
    out_ll = np.zeros_like(ll)
    c, arrays = make_locus_calls(len(rps), H, haploid)
    rc = ref_lib().ltr_ref_genotype_locus(int(haploid), len(rps), _ptr(rps, _ip), H, _ptr(ll, _dp), _ptr(p1, _dp),
                                          _ptr(p2, _dp), _ptr(out_ll, _dp), C.byref(c))
    if rc != 0:
        raise RuntimeError("ltr_ref_genotype_locus failed rc=%d" % rc)
    arrays["total_ll"] = c.total_ll
    arrays["ll_clamped"] = out_ll
    return arrays


def ref_seed_bases(locus, n_reads):
    seeds = np.zeros(n_reads, dtype=np.int32)
    rc = ref_lib().ltr_ref_seed_bases(C.byref(locus), _ptr(seeds, _ip))
    if rc != 0:
        raise RuntimeError("ltr_ref_seed_bases failed")
    return seeds


def make_params(aln_params=None, indel_flank_len=5):
    p = OracleParams()
    oracle_lib().ltr_oracle_default_params(C.byref(p))
    if aln_params is not None:
        (p.ins_ins, p.ins_match, p.del_del, p.del_match, p.match_match, p.match_ins,
         p.match_del) = [float(x) for x in aln_params]
    p.indel_flank_len = indel_flank_len
    return p


def _ptr(a, t):
    return a.ctypes.data_as(t)


def viterbi_pair(full_hap, read, aln_params=None, indel_flank_len=5):
    p = make_params(aln_params, indel_flank_len)
    cells = C.c_int64(0)
    h = full_hap if isinstance(full_hap, bytes) else full_hap.encode()
    r = read if isinstance(read, bytes) else read.encode()
    v = oracle_lib().ltr_oracle_viterbi_pair_cells(h, len(h), r, len(r), C.byref(p), C.byref(cells))
    return v, cells.value


def process_reads(locus, n_reads, n_alleles, fill=0.0, which="oracle"):
    """Run HapAligner::process_reads on a FlatLocus. Returns (ll[P,H], seeds[P], seconds)."""
    ll = np.full((n_reads, n_alleles), fill, dtype=np.float64)
    seeds = np.full(n_reads, -12345, dtype=np.int32)
    sec = C.c_double(0.0)
    if which == "oracle":
        rc = oracle_lib().ltr_oracle_process_reads(C.byref(locus), _ptr(ll, _dp), _ptr(seeds, _ip))
    else:
        rc = ref_lib().ltr_ref_process_reads(C.byref(locus), _ptr(ll, _dp), _ptr(seeds, _ip), C.byref(sec))
    if rc != 0:
        raise RuntimeError("process_reads(%s) failed rc=%d" % (which, rc))
    return ll, seeds, sec.value


def trim_read(locus, read_index):
    n = len(locus.reads[read_index].seq)
    buf = C.create_string_buffer(n + 16)
    k = oracle_lib().ltr_oracle_trim_read(C.byref(locus), read_index, buf)
    if k < 0:
        raise RuntimeError("trim failed")
    return buf.value[:k]


def viterbi_batch(batch, aln_params=None, indel_flank_len=5, n_threads=1):
    """batch: dict with numpy arrays locus_hap_begin, locus_read_begin, hap_off, hap_bytes,
    read_off, read_bytes (see include/longtr_b200.h ltr_viterbi_batch). Returns (ll, cells)."""
    p = make_params(aln_params, indel_flank_len)
    lhb = np.ascontiguousarray(batch["locus_hap_begin"], dtype=np.uint32)
    lrb = np.ascontiguousarray(batch["locus_read_begin"], dtype=np.uint32)
    n_loci = len(lhb) - 1
    nout = int(np.sum((lhb[1:] - lhb[:-1]).astype(np.int64) * (lrb[1:] - lrb[:-1]).astype(np.int64)))
    out = np.zeros(nout, dtype=np.float64)
    cells = C.c_int64(0)
    hoff = np.ascontiguousarray(batch["hap_off"], dtype=np.uint32)
    roff = np.ascontiguousarray(batch["read_off"], dtype=np.uint32)
    hb = np.ascontiguousarray(batch["hap_bytes"], dtype=np.uint8)
    rb = np.ascontiguousarray(batch["read_bytes"], dtype=np.uint8)
    rc = oracle_lib().ltr_oracle_viterbi_batch(n_loci, _ptr(lhb, _u32p), _ptr(lrb, _u32p),
                                               _ptr(hoff, _u32p), _ptr(hb, _u8p), _ptr(roff, _u32p),
                                               _ptr(rb, _u8p), C.byref(p), _ptr(out, _dp),
                                               C.byref(cells), n_threads)
    if rc != 0:
        raise RuntimeError("oracle batch failed")
    return out, cells.value


def ref_viterbi_batch(batch, aln_params=None, n_threads=1, post=None):
    """The reference's own HapAligner::process_reads over a flattened batch (haplotypes must carry
    35 bp flanks, reads >= 11 bp). Returns (ll, seconds inside process_reads, max over threads)."""
    lhb = np.ascontiguousarray(batch["locus_hap_begin"], dtype=np.uint32)
    lrb = np.ascontiguousarray(batch["locus_read_begin"], dtype=np.uint32)
    nout = int(np.sum((lhb[1:] - lhb[:-1]).astype(np.int64) * (lrb[1:] - lrb[:-1]).astype(np.int64)))
    out = np.zeros(nout, dtype=np.float64)
    hoff = np.ascontiguousarray(batch["hap_off"], dtype=np.uint32)
    roff = np.ascontiguousarray(batch["read_off"], dtype=np.uint32)
    hb = np.ascontiguousarray(batch["hap_bytes"], dtype=np.uint8)
    rb = np.ascontiguousarray(batch["read_bytes"], dtype=np.uint8)
    sec = C.c_double(0.0)
    if aln_params is None:
        npar, par = 0, (C.c_float * 7)()
    else:
        npar, par = 7, (C.c_float * 7)(*[float(x) for x in aln_params])
    pargs = [None, None, None, None, None]
    out_post = None
    if post is not None:
        H = (lhb[1:] - lhb[:-1]).astype(np.int64)
        out_post = np.zeros(int(np.sum(H * H)), dtype=np.float64)
        keep = [np.ascontiguousarray(post["locus_sread_begin"], dtype=np.uint32),
                np.ascontiguousarray(post["pool_index"], dtype=np.uint32),
                np.ascontiguousarray(post["log_p1"], dtype=np.float64),
                np.ascontiguousarray(post["log_p2"], dtype=np.float64)]
        pargs = [_ptr(keep[0], _u32p), _ptr(keep[1], _u32p), _ptr(keep[2], _dp), _ptr(keep[3], _dp),
                 _ptr(out_post, _dp)]
    rc = ref_lib().ltr_ref_viterbi_batch(len(lhb) - 1, _ptr(lhb, _u32p), _ptr(lrb, _u32p), _ptr(hoff, _u32p),
                                         _ptr(hb, _u8p), _ptr(roff, _u32p), _ptr(rb, _u8p), npar, par,
                                         n_threads, _ptr(out, _dp), C.byref(sec), *pargs)
    if rc != 0:
        raise RuntimeError("ltr_ref_viterbi_batch failed rc=%d" % rc)
    if post is not None:
        return out, sec.value, out_post
    return out, sec.value


def log_sample_posteriors(ll, log_p1, log_p2, sample_label, n_samples, haploid=False, which="oracle"):
    """Returns (ll_clamped, post[S,H,H], totals[S], total, best[S,2])."""
    ll = np.array(ll, dtype=np.float64, order="C", copy=True)
    R, H = ll.shape
    p1 = np.ascontiguousarray(log_p1, dtype=np.float64)
    p2 = np.ascontiguousarray(log_p2, dtype=np.float64)
    lab = np.ascontiguousarray(sample_label, dtype=np.int32)
    post = np.zeros((n_samples, H, H), dtype=np.float64)
    tot = np.zeros(n_samples, dtype=np.float64)
    best = np.zeros((n_samples, 2), dtype=np.int32)
    if which == "oracle":
        total = oracle_lib().ltr_oracle_log_sample_posteriors(int(haploid), n_samples, R, H, _ptr(ll, _dp),
                                                              _ptr(p1, _dp), _ptr(p2, _dp), _ptr(lab, _ip),
                                                              _ptr(post, _dp), _ptr(tot, _dp))
        oracle_lib().ltr_oracle_optimal_haplotypes(n_samples, H, _ptr(post, _dp), _ptr(best, _ip))
    else:
        assert np.all(np.diff(lab) >= 0), "reference needs sample-major reads"
        rps = np.bincount(lab, minlength=n_samples).astype(np.int32)
        out_ll = np.zeros_like(ll)
        total = ref_lib().ltr_ref_log_sample_posteriors(int(haploid), n_samples, _ptr(rps, _ip), H,
                                                        _ptr(ll, _dp), _ptr(p1, _dp), _ptr(p2, _dp),
                                                        _ptr(out_ll, _dp), _ptr(post, _dp), _ptr(tot, _dp),
                                                        _ptr(best, _ip))
        ll = out_ll
    return ll, post, tot, total, best


# ---- IO-less per-locus genotyper of the reference (full_driver.cpp): VCF record text ---------------------------
def full_locus_traces(cases):
    """ltr_ref_trace: per case (VCF record text, [calls]) where every call of Genotyper::calc_log_sample_posteriors inside
    SeqStutterGenotyper::genotype is a dict H, R, S, alleles, seeds, labels, ll, p1, p2, post, totals, gts (doubles as
    hex-float strings, exactly as the reference held them)."""
    build()
    exe = os.path.join(_HERE, "_ref", "ltr_ref_trace")
    p = subprocess.run([exe], input="".join(_case_text(c) for c in cases).encode(), stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, timeout=600)
    if p.returncode != 0:
        raise RuntimeError("ltr_ref_trace failed rc=%d: %s" % (p.returncode, p.stderr.decode()[-400:]))
    out, res, pos = p.stdout.decode(), [], 0
    while pos < len(out):
        assert out.startswith("RECORD ", pos)
        nl = out.index("\n", pos)
        n = int(out[pos + 7:nl])
        rec = out[nl + 1:nl + 1 + n].rstrip("\n")
        pos = nl + 1 + n + 1
        assert out.startswith("TRACE ", pos)
        nl = out.index("\n", pos)
        n = int(out[pos + 6:nl])
        body = out[nl + 1:nl + 1 + n]
        pos = nl + 1 + n
        calls, cur = [], None
        for line in body.split("\n"):
            f = line.split()
            if not f:
                continue
            if f[0] == "CALL":
                cur = dict(H=int(f[1]), R=int(f[2]), S=int(f[3]))
                calls.append(cur)
            elif f[0] == "ALLELES":
                cur["alleles"] = ["" if a == "-" else a for a in f[1:]]
            elif f[0] == "BLOCKS":
                cur["repeat_start"], cur["repeat_end"], cur["lflank"], cur["rflank"] = int(f[1]), int(f[2]), f[3], f[4]
            elif f[0] in ("SEEDS", "LABELS", "GTS"):
                cur[f[0].lower()] = [int(x) for x in f[1:]]
            else:
                assert int(f[1]) == len(f) - 2
                cur[f[0].lower()] = f[2:]
        res.append((rec, calls))
    assert len(res) == len(cases)
    return res


def full_available(which):
    """which = 'full' (all-CPU reference) or 'gpu' (reference + integration/reference_binding.cpp -> GPU)."""
    return os.path.exists(os.path.join(_HERE, "_ref", "ltr_ref_%s" % which))


def _case_text(case):
    st = case.get("stutter", (0.95, 0.05, 0.05, 0.95, 0.01, 0.01))
    params = case.get("aln_params")
    t = [case["chrom_name"], case["chrom_seq"], case["region_start"], case["region_stop"], case["motif"],
         case["region_name"], len(case["samples"])] + list(case["samples"]) + list(case["n_p1s"]) + list(case["n_p2s"])
    t += ["%.17g" % x for x in st] + [case.get("stutter_motif", "A"), case.get("stutter_period", len(case["motif"])),
                                     int(case.get("haploid", False)), case.get("indel_flank_len", 5), case.get("switch", 0)]
    t += [0] if params is None else [7] + ["%.9g" % x for x in params]
    t.append(len(case["reads"]))
    for r in case["reads"]:
        t += [r["start"], r["stop"], int(r["rev"]), r["sample"], r["name"], r["seq"], r["qual"], r["aln"], r["cigar"],
              "%.17g" % r["log_p1"], "%.17g" % r["log_p2"]]
    return " ".join(str(x) for x in t) + "\n"


def full_locus_records(cases, which="full", switches=None):
    """SeqStutterGenotyper ctor -> genotype -> write_vcf_record for every case; returns the VCF record texts
    ('' where genotype() failed).  switches: the reference's output switches as a mask of the library's LTR_VCF_* bits
    (None = the reference's defaults)."""
    build()
    exe = os.path.join(_HERE, "_ref", "ltr_ref_%s" % which)
    env = dict(os.environ)
    env.pop("LTR_REF_OUTPUT_SWITCHES", None)
    if switches is not None:
        env["LTR_REF_OUTPUT_SWITCHES"] = str(int(switches))
    p = subprocess.run([exe], input="".join(_case_text(c) for c in cases).encode(), stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, timeout=600, env=env)
    if p.returncode != 0:
        raise RuntimeError("ltr_ref_%s failed rc=%d: %s" % (which, p.returncode, p.stderr.decode()[-400:]))
    out, recs, pos = p.stdout.decode(), [], 0
    while pos < len(out):
        assert out.startswith("RECORD ", pos)
        nl = out.index("\n", pos)
        n = int(out[pos + 7:nl])
        recs.append(out[nl + 1:nl + 1 + n].rstrip("\n"))
        pos = nl + 1 + n + 1
    assert len(recs) == len(cases)
    return recs


# ---- candidate-haplotype clustering (HaplotypeGenerator::needleman_wunsch / greedy_clustering) ------------------------
def edit_score(cent, read, T, which="oracle"):
    """Score HaplotypeGenerator::needleman_wunsch(cent, read, score, T) leaves behind."""
    a = cent.encode() if isinstance(cent, str) else bytes(cent)
    b = read.encode() if isinstance(read, str) else bytes(read)
    if which == "ref":
        return int(ref_lib().ltr_ref_edit_score(a, len(a), b, len(b), T))
    return int(oracle_lib().ltr_oracle_edit_score(a, len(a), b, len(b), T))


def greedy_cluster(seq_bytes, seq_off, items, T, which="oracle"):
    """greedy_clustering over the sequences items[] -> (ok, centroid_of[len(items)], n_centroids)."""
    items = np.ascontiguousarray(items, dtype=np.uint32)
    cent = np.full(max(1, len(items)), -1, dtype=np.int32)
    n = C.c_int32(0)
    fn = ref_lib().ltr_ref_greedy_cluster if which == "ref" else oracle_lib().ltr_oracle_greedy_cluster
    ok = fn(_ptr(seq_bytes, _u8p), _ptr(seq_off, _u32p), _ptr(items, _u32p), len(items), T, _ptr(cent, _ip), C.byref(n))
    return int(ok), cent[:len(items)], n.value


# ---- the reference's EMStutterGenotyper compiled in place (oracle/em_driver.cpp -> oracle/_ref/libltr_ref_em.so) ----------
_EM_SO = os.path.join(_HERE, "_ref", "libltr_ref_em.so")


def ref_em_available():
    return os.path.exists(_EM_SO)


def ref_em_train(reads_per_sample, bp_diff, log_p1, log_p2, motif_len, haploid=False, max_iter=100, abs_conv=0.01,
                 frac_conv=0.001):
    """EMStutterGenotyper(...).train(...) of the reference on one locus (reads sample-major).
    -> dict(trained, params[6] = in_geom, in_up, in_down, out_geom, out_up, out_down, n_iter, lls[n_iter], log_gt_priors)."""
    import ctypes as C

    import numpy as np
    lib = C.CDLL(_EM_SO)
    i32p, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    lib.ltr_ref_em_train.restype = C.c_int
    lib.ltr_ref_em_train.argtypes = [C.c_uint32, i32p, i32p, dp, dp, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double,
                                     dp, i32p, dp, dp, i32p]
    rps = np.ascontiguousarray(reads_per_sample, dtype=np.int32)
    bd = np.ascontiguousarray(bp_diff, dtype=np.int32)
    p1 = np.ascontiguousarray(log_p1, dtype=np.float64)
    p2 = np.ascontiguousarray(log_p2, dtype=np.float64)
    params = np.zeros(6)
    lls = np.zeros(max_iter + 2)
    pri = np.zeros(64)
    n_iter, n_all = C.c_int32(0), C.c_int32(0)
    P = lambda a, t: a.ctypes.data_as(t)
    trained = lib.ltr_ref_em_train(len(rps), P(rps, i32p), P(bd, i32p), P(p1, dp), P(p2, dp), motif_len, int(haploid), max_iter,
                                   abs_conv, frac_conv, P(params, dp), C.byref(n_iter), P(lls, dp), P(pri, dp), C.byref(n_all))
    return dict(trained=bool(trained), params=params, n_iter=n_iter.value, lls=lls[:n_iter.value],
                log_gt_priors=pri[:min(64, n_all.value)], n_alleles=n_all.value)
