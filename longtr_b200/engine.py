"""Python host wrapper over the C ABI (include/longtr_b200.h).

``Engine`` owns one ``ltr_ctx`` (one per process and GPU); ``Job`` is a batch of loci kept
resident in HBM (``ltr_job_*``).  Everything below calls the CUDA library through ctypes;
nothing here computes on the CPU and nothing falls back to the oracle.
"""
import ctypes as C
import weakref

import numpy as np

from . import abi


class LongTRError(RuntimeError):
    pass


def _check(lib, ctx, rc, what):
    if rc != abi.LTR_OK:
        detail = lib.ltr_last_error(ctx).decode() if ctx else ""
        raise LongTRError("%s failed: %s %s" % (what, lib.ltr_strerror(rc).decode(), detail))


class Job:
    """A flattened batch of loci resident on the GPU (``ltr_job``)."""

    def __init__(self, engine, handle, keep):
        self._e = engine
        self._h = handle
        self._keep = keep
        engine._jobs.add(self)  # the engine closes its live jobs before destroying the context
        n_ll, n_post, n_tot = C.c_uint64(), C.c_uint64(), C.c_uint64()
        engine.lib.ltr_job_sizes(handle, C.byref(n_ll), C.byref(n_post), C.byref(n_tot))
        self.n_ll, self.n_post, self.n_totals = n_ll.value, n_post.value, n_tot.value

    def run(self):
        _check(self._e.lib, self._e.ctx, self._e.lib.ltr_job_run(self._e.ctx, self._h), "ltr_job_run")
        return self.stats()

    def download(self, want_post=True, out_ll=None, out_post=None, out_totals=None):
        ll = out_ll if out_ll is not None else np.empty(self.n_ll, dtype=np.float64)
        post = tot = None
        pp = tp = None
        if want_post and self.n_post:
            post = out_post if out_post is not None else np.empty(self.n_post, dtype=np.float64)
            tot = out_totals if out_totals is not None else np.empty(self.n_totals, dtype=np.float64)
            pp, tp = abi.ptr(post, abi._dp), abi.ptr(tot, abi._dp)
        rc = self._e.lib.ltr_job_download(self._e.ctx, self._h, abi.ptr(ll, abi._dp), pp, tp)
        _check(self._e.lib, self._e.ctx, rc, "ltr_job_download")
        return ll, post, tot

    def stats(self):
        s = abi.JobStats()
        self._e.lib.ltr_job_get_stats(self._h, C.byref(s))
        return s

    def wait(self):
        """ltr_job_wait of a job started with Engine.submit_job: blocks until the results are in the caller's arrays."""
        _check(self._e.lib, self._e.ctx, self._e.lib.ltr_job_wait(self._e.ctx, self._h), "ltr_job_wait")
        return self.stats()

    def poll(self):
        rc = self._e.lib.ltr_job_poll(self._e.ctx, self._h)
        if rc < 0:
            _check(self._e.lib, self._e.ctx, rc, "ltr_job_poll")
        return rc == 1

    def close(self):
        if self._h is not None:
            if self._e.ctx:
                self._e.lib.ltr_job_destroy(self._e.ctx, self._h)
            self._h = None
            self._e._jobs.discard(self)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Pipeline:
    """``ltr_pipeline``: loci handed over one at a time, processed in batches on worker threads (each with its own
    context); results come back by tag."""

    def __init__(self, device=0, batch_loci=256, slots=2):
        self.lib = abi.load()
        self.h = C.c_void_p()
        rc = self.lib.ltr_pipeline_create(device, batch_loci, slots, C.byref(self.h))
        if rc != abi.LTR_OK:
            self.h = None
            raise LongTRError("ltr_pipeline_create: %s" % self.lib.ltr_strerror(rc).decode())

    def submit(self, locus, tag, fill=0.0):
        rc = self.lib.ltr_pipeline_submit(self.h, C.byref(locus), tag, None, fill)
        if rc != abi.LTR_OK:
            raise LongTRError("ltr_pipeline_submit: %s" % self.lib.ltr_strerror(rc).decode())

    def flush(self):
        self.lib.ltr_pipeline_flush(self.h)

    def next(self, wait=True):
        """(tag, ll[n_reads, n_alleles], seeds, status) of the next finished locus, or None."""
        tag, nr, na, st = C.c_uint64(), C.c_int32(), C.c_int32(), C.c_int()
        ll, seeds = abi._dp(), abi._i32p()
        got = self.lib.ltr_pipeline_next(self.h, 1 if wait else 0, C.byref(tag), C.byref(nr), C.byref(na), C.byref(ll),
                                         C.byref(seeds), C.byref(st))
        if got <= 0:
            return None
        n = nr.value * na.value
        a = np.ctypeslib.as_array(ll, (max(1, n),))[:n].reshape(nr.value, na.value).copy()
        s = np.ctypeslib.as_array(seeds, (max(1, nr.value),))[:nr.value].copy()
        return tag.value, a, s, st.value

    def close(self):
        if self.h:
            self.lib.ltr_pipeline_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """One GPU context. Raises ``LongTRError`` when no CUDA device is usable (no CPU fallback)."""

    def __init__(self, device=0):
        self.lib = abi.load()
        self.ctx = C.c_void_p()
        self.device = device
        self._jobs = weakref.WeakSet()
        rc = self.lib.ltr_ctx_create(device, C.byref(self.ctx))
        if rc != abi.LTR_OK:
            self.ctx = None
            raise LongTRError("ltr_ctx_create(%d): %s" % (device, self.lib.ltr_strerror(rc).decode()))

    def close(self):
        if self.ctx:
            for job in list(self._jobs):
                job.close()
            self.lib.ltr_ctx_destroy(self.ctx)
            self.ctx = None

    def set_band(self, half_width):
        """ltr_ctx_set_band: < 0 disables the banded kernel, 0 = automatic margin, > 0 = margin in diagonals.
        Results never depend on it (uncertified pairs are re-run over the full matrix)."""
        _check(self.lib, self.ctx, self.lib.ltr_ctx_set_band(self.ctx, int(half_width)), "ltr_ctx_set_band")

    def set_read_encoding(self, encoding):
        """ltr_ctx_set_read_encoding: 0 = one byte per base, 1 = one 4-bit stream over all reads (abi.pack_reads_4bit)."""
        _check(self.lib, self.ctx, self.lib.ltr_ctx_set_read_encoding(self.ctx, int(encoding)), "ltr_ctx_set_read_encoding")

    def set_plan(self, mode):
        """ltr_ctx_set_plan: 0 automatic, 1 plan on the host, 2 plan on the device.  Results never depend on it."""
        _check(self.lib, self.ctx, self.lib.ltr_ctx_set_plan(self.ctx, int(mode)), "ltr_ctx_set_plan")

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- HapAligner::process_reads, long path, for a flattened batch of loci ------------------
    def viterbi_ll(self, batch, aln_params=None, indel_flank_len=5):
        vb, keep = abi.make_viterbi_batch(batch)
        p = abi.make_params(aln_params, indel_flank_len)
        out = np.empty(abi.ll_size(batch), dtype=np.float64)
        st = abi.JobStats()
        rc = self.lib.ltr_viterbi_ll(self.ctx, C.byref(p), C.byref(vb), abi.ptr(out, abi._dp), C.byref(st))
        _check(self.lib, self.ctx, rc, "ltr_viterbi_ll")
        return out, st

    # -- Genotyper::calc_log_sample_posteriors for one locus -----------------------------------
    def posteriors(self, ll, log_p1, log_p2, sample_label, n_samples, haploid=False):
        ll = np.array(ll, dtype=np.float64, order="C", copy=True)
        R, H = ll.shape
        p1 = np.ascontiguousarray(log_p1, dtype=np.float64)
        p2 = np.ascontiguousarray(log_p2, dtype=np.float64)
        lab = np.ascontiguousarray(sample_label, dtype=np.int32)
        post = np.empty((n_samples, H, H), dtype=np.float64)
        tot = np.empty(n_samples, dtype=np.float64)
        total = C.c_double(0.0)
        rc = self.lib.ltr_posteriors(self.ctx, int(haploid), n_samples, R, H, abi.ptr(ll, abi._dp),
                                     abi.ptr(p1, abi._dp), abi.ptr(p2, abi._dp), abi.ptr(lab, abi._i32p),
                                     abi.ptr(post, abi._dp), abi.ptr(tot, abi._dp), C.byref(total))
        _check(self.lib, self.ctx, rc, "ltr_posteriors")
        return ll, post, tot, total.value

    # -- HapAligner::process_reads, homopolymer / --stutter-align-len path, for a batch of loci ---------
    def stutter_ll(self, batch, aln_params=None, out=None, per_locus_status=False):
        """ltr_stutter_ll; with per_locus_status=True ltr_stutter_ll_status: returns (out, stats, status[n_loci]) and a
        locus that cannot be processed fails alone."""
        sb, keep = abi.make_stutter_batch(batch)
        p = abi.make_params(aln_params, 5)
        if out is None:
            out = np.zeros(abi.stutter_ll_size(batch), dtype=np.float64)
        st = abi.JobStats()
        if per_locus_status:
            status = np.zeros(len(batch["locus_read_begin"]) - 1, dtype=np.int32)
            rc = self.lib.ltr_stutter_ll_status(self.ctx, C.byref(p), C.byref(sb), abi.ptr(out, abi._dp),
                                                abi.ptr(status, abi._i32p), C.byref(st))
            _check(self.lib, self.ctx, rc, "ltr_stutter_ll_status")
            return out, st, status
        rc = self.lib.ltr_stutter_ll(self.ctx, C.byref(p), C.byref(sb), abi.ptr(out, abi._dp), C.byref(st))
        _check(self.lib, self.ctx, rc, "ltr_stutter_ll")
        return out, st

    # -- posteriors (GPU) + Genotyper::extract_genotypes_and_likelihoods (host) for one locus ----------
    def genotype_locus(self, ll, log_p1, log_p2, reads_per_sample, haploid=False):
        ll = np.array(ll, dtype=np.float64, order="C", copy=True)
        R, H = ll.shape
        rps = np.ascontiguousarray(reads_per_sample, dtype=np.int32)
        p1 = np.ascontiguousarray(log_p1, dtype=np.float64)
        p2 = np.ascontiguousarray(log_p2, dtype=np.float64)
        c, arrays = abi.make_locus_calls(len(rps), H, haploid)
        rc = self.lib.ltr_genotype_locus(self.ctx, int(haploid), len(rps), abi.ptr(rps, abi._i32p), H,
                                         abi.ptr(ll, abi._dp), abi.ptr(p1, abi._dp), abi.ptr(p2, abi._dp), C.byref(c))
        _check(self.lib, self.ctx, rc, "ltr_genotype_locus")
        arrays["total_ll"] = c.total_ll
        arrays["ll_clamped"] = ll
        return arrays

    def genotype_locus_pruned(self, ll, log_p1, log_p2, reads_per_sample, haploid=False, seeds=None):
        """genotype_locus + the reference's removal of uncalled alleles and second posterior pass."""
        ll = np.array(ll, dtype=np.float64, order="C", copy=True)
        R, H = ll.shape
        rps = np.ascontiguousarray(reads_per_sample, dtype=np.int32)
        p1 = np.ascontiguousarray(log_p1, dtype=np.float64)
        p2 = np.ascontiguousarray(log_p2, dtype=np.float64)
        sd = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.int32)
        kept = np.zeros(H, dtype=np.int32)
        nk = C.c_int32(0)
        c, arrays = abi.make_locus_calls(len(rps), H, haploid)
        rc = self.lib.ltr_genotype_locus_pruned(self.ctx, int(haploid), len(rps), abi.ptr(rps, abi._i32p), H,
                                                abi.ptr(ll, abi._dp), abi.ptr(p1, abi._dp), abi.ptr(p2, abi._dp),
                                                None if sd is None else abi.ptr(sd, abi._i32p), abi.ptr(kept, abi._i32p),
                                                C.byref(nk), C.byref(c))
        _check(self.lib, self.ctx, rc, "ltr_genotype_locus_pruned")
        arrays["total_ll"] = c.total_ll
        arrays["kept"] = kept[:nk.value].copy()
        return arrays

    def create_job(self, batch, post=None, aln_params=None, indel_flank_len=5):
        vb, keep = abi.make_viterbi_batch(batch)
        p = abi.make_params(aln_params, indel_flank_len)
        pb_ref = None
        keep2 = None
        if post is not None:
            pb, keep2 = abi.make_posterior_batch(post)
            pb_ref = C.byref(pb)
        h = C.c_void_p()
        rc = self.lib.ltr_job_create(self.ctx, C.byref(p), C.byref(vb), pb_ref, C.byref(h))
        _check(self.lib, self.ctx, rc, "ltr_job_create")
        return Job(self, h, (keep, keep2))

    def submit_job(self, batch, post=None, aln_params=None, indel_flank_len=5, out_ll=None, out_post=None, out_totals=None,
                   prepared=None):
        """ltr_job_submit: enqueue upload + plan + kernels + posteriors + download and return at once; the arrays of
        ``batch`` / ``post`` and the output arrays must stay alive and untouched until ``Job.wait()``.
        ``prepared`` = (ViterbiBatch, PosteriorBatch or None, keepalive) skips the per-call struct building."""
        if prepared is None:
            vb, keep = abi.make_viterbi_batch(batch)
            pb, keep2 = (abi.make_posterior_batch(post) if post is not None else (None, None))
        else:
            vb, pb, keep = prepared
            keep2 = None
        p = abi.make_params(aln_params, indel_flank_len)
        h = C.c_void_p()
        rc = self.lib.ltr_job_submit(self.ctx, C.byref(p), C.byref(vb), C.byref(pb) if pb is not None else None,
                                     None if out_ll is None else abi.ptr(out_ll, abi._dp),
                                     None if out_post is None else abi.ptr(out_post, abi._dp),
                                     None if out_totals is None else abi.ptr(out_totals, abi._dp), C.byref(h))
        _check(self.lib, self.ctx, rc, "ltr_job_submit")
        return Job(self, h, (keep, keep2, vb, pb, out_ll, out_post, out_totals))

    def posteriors_batch(self, locus_hap_begin, locus_read_begin, ll, post):
        """ltr_posteriors_batch: posteriors (optionally mate pairs / removal of uncalled alleles) for many loci from LL
        matrices the caller holds.  Returns (post flat, totals flat, kept_mask[n_haps])."""
        lhb = np.ascontiguousarray(locus_hap_begin, dtype=np.uint32)
        lrb = np.ascontiguousarray(locus_read_begin, dtype=np.uint32)
        ll = np.ascontiguousarray(ll, dtype=np.float64)
        pb, keep = abi.make_posterior_batch(post)
        H = np.diff(lhb).astype(np.int64)
        S = np.asarray(post["locus_n_samples"], dtype=np.int64)
        out_post = np.zeros(int(np.sum(S * H * H)), dtype=np.float64)
        out_tot = np.zeros(int(np.sum(S)), dtype=np.float64)
        kept = np.zeros(int(lhb[-1]), dtype=np.uint8)
        rc = self.lib.ltr_posteriors_batch(self.ctx, len(lhb) - 1, abi.ptr(lhb, abi._u32p), abi.ptr(lrb, abi._u32p),
                                           abi.ptr(ll, abi._dp), C.byref(pb), abi.ptr(out_post, abi._dp),
                                           abi.ptr(out_tot, abi._dp), abi.ptr(kept, abi._u8p))
        _check(self.lib, self.ctx, rc, "ltr_posteriors_batch")
        return out_post, out_tot, kept

    def process_reads_flat(self, locus, n_reads, n_alleles, fill=0.0):
        ll = np.full((n_reads, n_alleles), fill, dtype=np.float64)
        seeds = np.full(n_reads, -12345, dtype=np.int32)
        rc = self.lib.ltr_process_reads_flat(self.ctx, C.byref(locus), abi.ptr(ll, abi._dp),
                                             abi.ptr(seeds, abi._i32p))
        _check(self.lib, self.ctx, rc, "ltr_process_reads_flat")
        return ll, seeds

    def process_reads_flat_batch(self, loci, shapes, fill=0.0):
        """ltr_process_reads_flat_batch: ``loci`` = list of FlatLocus, ``shapes`` = [(n_reads, n_alleles)];
        returns ([ll], [seeds]) per locus.  All long-path loci run as ONE GPU job."""
        n = len(loci)
        arr = (abi.FlatLocus * max(1, n))(*loci)
        fills = fill if isinstance(fill, (list, tuple)) else [fill] * n
        lls = [np.full(sh, f, dtype=np.float64) for sh, f in zip(shapes, fills)]
        seeds = [np.full(sh[0], -12345, dtype=np.int32) for sh in shapes]
        ll_ptrs = (abi._dp * max(1, n))(*[abi.ptr(a, abi._dp) for a in lls])
        seed_ptrs = (abi._i32p * max(1, n))(*[abi.ptr(a, abi._i32p) for a in seeds])
        rc = self.lib.ltr_process_reads_flat_batch(self.ctx, n, arr, ll_ptrs, seed_ptrs)
        _check(self.lib, self.ctx, rc, "ltr_process_reads_flat_batch")
        return lls, seeds

    # -- HaplotypeGenerator::needleman_wunsch / greedy_clustering for many pairs / sets ------------------
    def edit_distances(self, seq_bytes, seq_off, pair_a, pair_b, pair_T):
        """ltr_edit_distances on packed sequences (abi.pack_seqs); returns (scores int32[n_pairs], stats)."""
        pair_a = np.ascontiguousarray(pair_a, dtype=np.uint32)
        pair_b = np.ascontiguousarray(pair_b, dtype=np.uint32)
        pair_T = np.ascontiguousarray(pair_T, dtype=np.int32)
        out = np.full(len(pair_a), -1, dtype=np.int32)
        st = abi.JobStats()
        rc = self.lib.ltr_edit_distances(self.ctx, abi.ptr(seq_bytes, abi._u8p), abi.ptr(seq_off, abi._u32p),
                                         len(seq_off) - 1, abi.ptr(pair_a, abi._u32p), abi.ptr(pair_b, abi._u32p),
                                         abi.ptr(pair_T, abi._i32p), len(pair_a), abi.ptr(out, abi._i32p), C.byref(st))
        _check(self.lib, self.ctx, rc, "ltr_edit_distances")
        return out, st

    def cluster_greedy(self, seq_bytes, seq_off, set_begin, set_items, set_T):
        """ltr_cluster_greedy; returns (centroid_of int32[n_items], n_centroids int32[n_sets], ok uint8[n_sets], stats)."""
        set_begin = np.ascontiguousarray(set_begin, dtype=np.uint32)
        set_items = np.ascontiguousarray(set_items, dtype=np.uint32)
        set_T = np.ascontiguousarray(set_T, dtype=np.int32)
        n_sets = len(set_begin) - 1
        cent = np.full(max(1, len(set_items)), -2, dtype=np.int32)
        ncent = np.full(max(1, n_sets), -2, dtype=np.int32)
        ok = np.full(max(1, n_sets), 255, dtype=np.uint8)
        st = abi.JobStats()
        rc = self.lib.ltr_cluster_greedy(self.ctx, abi.ptr(seq_bytes, abi._u8p), abi.ptr(seq_off, abi._u32p),
                                         len(seq_off) - 1, abi.ptr(set_begin, abi._u32p), abi.ptr(set_items, abi._u32p),
                                         abi.ptr(set_T, abi._i32p), n_sets, abi.ptr(cent, abi._i32p),
                                         abi.ptr(ncent, abi._i32p), abi.ptr(ok, abi._u8p), C.byref(st))
        _check(self.lib, self.ctx, rc, "ltr_cluster_greedy")
        return cent[:len(set_items)], ncent[:n_sets], ok[:n_sets], st

    def fp64_issue_rate(self, kind=0):
        rate, ms = C.c_double(0.0), C.c_double(0.0)
        rc = self.lib.ltr_fp64_issue_rate(self.device, kind, C.byref(rate), C.byref(ms))
        _check(self.lib, self.ctx, rc, "ltr_fp64_issue_rate")
        return rate.value, ms.value
