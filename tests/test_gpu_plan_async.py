"""GPU: the plan built on the device (plan_kernels.cu) against the plan built on the host (make_plan) and the oracle,
and the asynchronous ltr_job_submit / ltr_job_wait pair with several jobs in flight on one context."""
import ctypes as C

import numpy as np
import pytest

import synth
from longtr_b200 import abi
from longtr_b200.engine import LongTRError
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

ONT = (-1.0, -0.458675, -1.0, -0.458675, -0.0202027, -4.60517, -4.60517)
ODD = (-0.7, -0.61, -0.35, -1.3, -0.013, -3.9, -4.4)

CASES = [
    (1, dict(n_loci=60), None),
    (2, dict(n_loci=40, n_lo=20, n_hi=400), None),
    (3, dict(n_loci=40, n_lo=200, n_hi=520, reads_hi=4, haps_hi=3), ONT),
    (4, dict(n_loci=40, n_lo=10, n_hi=150, weird=0.3), ODD),
    (5, dict(n_loci=40, n_lo=1, n_hi=40, weird=0.3), None),
    (6, dict(n_loci=6, n_lo=600, n_hi=1100, reads_hi=3, haps_hi=3, sub=0.02, indel=0.03), ONT),
]


@pytest.fixture()
def plan_engine(engine):
    yield engine
    engine.set_plan(0)
    engine.set_band(0)
    engine.set_read_encoding(0)


def _stats_tuple(st):
    return (st.n_pairs, st.n_cells, st.n_pairs_computed, st.n_band_pairs)


@pytest.mark.parametrize("seed,kw,params", CASES)
@pytest.mark.parametrize("band_w", [-1, 0, 5])
def test_device_plan_bit_exact(plan_engine, seed, kw, params, band_w):
    eng = plan_engine
    b = synth.make_pair_batch(seed, **kw)
    want, _cells = po.viterbi_batch(b, aln_params=params, n_threads=4)
    eng.set_band(band_w)
    eng.set_plan(1)
    host, st_h = eng.viterbi_ll(b, aln_params=params)
    eng.set_plan(2)
    dev, st_d = eng.viterbi_ll(b, aln_params=params)
    assert np.array_equal(host, want)
    bad = np.nonzero(dev != want)[0]
    assert len(bad) == 0, (bad[:10], dev[bad[:10]], want[bad[:10]])
    assert _stats_tuple(st_h) == _stats_tuple(st_d)
    assert st_d.plan_ms > 0.0 and st_h.plan_ms < st_d.plan_ms + 1.0


@pytest.mark.parametrize("seed", range(15))
def test_device_plan_pathological_batches(plan_engine, seed):
    eng = plan_engine
    b = synth.make_pathological_batch(seed)
    want, _ = po.viterbi_batch(b)
    eng.set_plan(2)
    got, st = eng.viterbi_ll(b)
    assert np.array_equal(got, want)


def test_device_plan_config3_with_posteriors(plan_engine):
    """A slice of the benchmark workload (about half of the pooled reads are duplicates after trimming): LL matrices and
    posteriors of the device-planned job equal the host-planned job's bit for bit; so do the plan statistics."""
    from longtr_b200 import workloads
    eng = plan_engine
    w = workloads.generate(3, 600)
    b, p = w.subset(600)
    res = {}
    for mode in (1, 2):
        eng.set_plan(mode)
        job = eng.create_job(b, p, aln_params=w.aln_params)
        st = job.run()
        ll, post, tot = job.download()
        res[mode] = (ll.copy(), post.copy(), tot.copy(), _stats_tuple(st), st.n_cells_computed, st.n_band_uncertified)
        job.close()
    assert np.array_equal(res[1][0], res[2][0])
    assert np.array_equal(res[1][1], res[2][1]) and np.array_equal(res[1][2], res[2][2])
    assert res[1][3:] == res[2][3:]
    assert not np.isnan(res[2][0]).any()
    want, _ = po.viterbi_batch(b, aln_params=w.aln_params, n_threads=4)
    assert np.array_equal(res[2][0], want)
    w.close()


@pytest.mark.parametrize("mode", [1, 2])
def test_async_jobs_in_flight(plan_engine, mode):
    """Five jobs submitted back to back on one context (they share three stream lanes), waited for in submission order:
    every job's results land in its own arrays and equal the oracle."""
    eng = plan_engine
    eng.set_plan(mode)
    jobs = []
    for i in range(5):
        b = synth.make_pair_batch(100 + i, n_loci=30 + 7 * i)
        out = np.full(abi.ll_size(b), 7.0)
        jobs.append((b, out, eng.submit_job(b, out_ll=out)))
    for b, out, job in jobs:
        st = job.wait()
        want, _ = po.viterbi_batch(b)
        assert np.array_equal(out, want)
        assert st.n_pairs == len(want)
        job.close()


def test_async_job_with_posteriors_matches_resident_job(plan_engine):
    from longtr_b200 import workloads
    eng = plan_engine
    w = workloads.generate(3, 300)
    b, p = w.subset(300)
    job = eng.create_job(b, p, aln_params=w.aln_params)
    job.run()
    ll, post, tot = job.download()
    job.close()
    out_ll, out_post, out_tot = np.zeros_like(ll), np.zeros_like(post), np.zeros_like(tot)
    j1 = eng.submit_job(b, p, aln_params=w.aln_params, out_ll=out_ll, out_post=out_post, out_totals=out_tot)
    out2 = np.zeros_like(ll)
    j2 = eng.submit_job(b, p, aln_params=w.aln_params, out_ll=out2)   # second job in flight, LL only
    assert j2.wait().n_pairs == len(ll)
    j1.wait()
    assert np.array_equal(out_ll, ll) and np.array_equal(out_post, post) and np.array_equal(out_tot, tot)
    assert np.array_equal(out2, ll)
    j1.close()
    j2.close()
    w.close()


@pytest.mark.parametrize("mode", [1, 2])
def test_malformed_batches_by_plan_mode(plan_engine, mode):
    """An empty read is found by the host plan at submit time and by the device plan on the device: reported by
    ltr_job_create, and by ltr_job_wait for asynchronous jobs.  The context stays usable."""
    eng = plan_engine
    eng.set_plan(mode)
    good = synth.make_pair_batch(3, n_loci=4)
    bad = dict(good)
    bad["read_off"] = good["read_off"].copy()
    bad["read_off"][2] = bad["read_off"][1]
    with pytest.raises(LongTRError, match="invalid"):
        eng.create_job(bad)
    with pytest.raises(LongTRError, match="invalid"):
        out = np.zeros(abi.ll_size(good))
        eng.submit_job(bad, out_ll=out).wait()
    bad2 = dict(good)
    bad2["read_off"] = good["read_off"].copy()
    bad2["read_off"][3] = bad2["read_off"][-1] + 100000      # points far outside the uploaded bytes
    with pytest.raises(LongTRError, match="invalid"):
        eng.viterbi_ll(bad2)
    got, _ = eng.viterbi_ll(good)
    want, _ = po.viterbi_batch(good)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("mode", [1, 2])
def test_posterior_batch_validation(plan_engine, mode):
    """the posterior half of a job is validated like the Viterbi half -- missing arrays, offsets that go
    backwards, pool indices outside the locus and unknown sample labels are errors, not crashes."""
    from longtr_b200 import workloads
    eng = plan_engine
    eng.set_plan(mode)
    w = workloads.generate(3, 40)
    b, p = w.subset(40)

    def attempt(post):
        job = eng.create_job(b, post, aln_params=w.aln_params)
        try:
            job.run()
        finally:
            job.close()

    q = dict(p)
    q["locus_sread_begin"] = p["locus_sread_begin"].copy()
    q["locus_sread_begin"][5] = q["locus_sread_begin"][7]      # goes backwards at locus 6
    with pytest.raises(LongTRError, match="invalid"):
        attempt(q)
    q = dict(p)
    q["pool_index"] = p["pool_index"].copy()
    q["pool_index"][17] = 10 ** 6
    with pytest.raises(LongTRError, match="invalid"):
        attempt(q)
    q = dict(p)
    q["sample_label"] = p["sample_label"].copy()
    q["sample_label"][3] = 9
    with pytest.raises(LongTRError, match="invalid"):
        attempt(q)
    vb, keep = abi.make_viterbi_batch(b)
    pb, keep2 = abi.make_posterior_batch(p)
    pb.log_p1 = None
    h = C.c_void_p()
    prm = abi.make_params(w.aln_params)
    assert eng.lib.ltr_job_create(eng.ctx, C.byref(prm), C.byref(vb), C.byref(pb), C.byref(h)) == -3
    attempt(p)   # the well-formed batch still runs
    w.close()


@pytest.mark.parametrize("seed,kw,params", [CASES[0], CASES[1], CASES[2], CASES[5]])
@pytest.mark.parametrize("plan", [1, 2])
def test_reads_as_4bit_stream_same_results(plan_engine, seed, kw, params, plan):
    """ltr_ctx_set_read_encoding(1): the reads of the batch travel as one 4-bit stream (half the bytes) and are expanded by the
    first kernel of the device plan (on the host for the small batches of the host plan) -- same bits, same statistics."""
    eng = plan_engine
    b = synth.make_pair_batch(seed, **kw)
    n_bases = int(np.asarray(b["read_off"])[-1])
    assert set(np.unique(np.asarray(b["read_bytes"])[:n_bases]).tolist()) <= set(b"ACGTN")
    eng.set_plan(plan)
    want, st_b = eng.viterbi_ll(b, aln_params=params)
    packed = dict(b, read_bytes=abi.pack_reads_4bit(b["read_bytes"], n_bases))
    assert len(packed["read_bytes"]) == (n_bases + 1) // 2
    eng.set_read_encoding(1)
    got, st_p = eng.viterbi_ll(packed, aln_params=params)
    eng.set_read_encoding(0)
    assert np.array_equal(got, want)
    assert _stats_tuple(st_p) == _stats_tuple(st_b)
    if plan == 2:
        assert st_b.h2d_bytes - st_p.h2d_bytes == n_bases - (n_bases + 1) // 2
    with pytest.raises(LongTRError):
        eng.set_read_encoding(2)


def test_4bit_stream_through_submit_wait(plan_engine):
    """The asynchronous pair with the packed encoding, odd total length, posteriors attached."""
    eng = plan_engine
    b = synth.make_pair_batch(77, n_loci=300)
    n_bases = int(np.asarray(b["read_off"])[-1])
    post = synth.make_posterior_inputs(b, 77) if hasattr(synth, "make_posterior_inputs") else None
    eng.set_plan(2)
    want, _ = eng.viterbi_ll(b)
    packed = dict(b, read_bytes=abi.pack_reads_4bit(b["read_bytes"], n_bases))
    eng.set_read_encoding(1)
    out = np.zeros(abi.ll_size(b))
    j = eng.submit_job(packed, None, out_ll=out)
    j.wait()
    j.close()
    eng.set_read_encoding(0)
    assert np.array_equal(out, want)
