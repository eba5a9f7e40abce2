"""GPU parity: posterior kernel vs the oracle (floating point: exp/log, tolerance stated)."""
import numpy as np
import pytest

import golden_util as gu
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

# CUDA exp/log are within 1 ulp of libm; sums of <= a few hundred terms of magnitude <= 1e3
RTOL, ATOL = 1e-12, 1e-10


def _case(rng, S, H, haploid):
    rps = rng.integers(1, 15, size=S)
    lab = np.repeat(np.arange(S), rps).astype(np.int32)
    R = len(lab)
    ll = -rng.exponential(20, size=(R, H))
    ll[rng.random((R, H)) < 0.05] = -700
    ll[rng.random((R, H)) < 0.02] = -1e9
    hp = rng.integers(0, 3, size=R)
    p1 = np.where(hp == 0, -1e-6, np.where(hp == 1, -1000.0, 0.0))
    p2 = np.where(hp == 0, -1000.0, np.where(hp == 1, -1e-6, 0.0))
    return ll, p1, p2, lab


def test_posteriors_match_oracle(engine):
    rng = np.random.default_rng(11)
    for t in range(60):
        S, H = int(rng.integers(1, 4)), int(rng.integers(1, 9))
        hap = (t % 5 == 0)
        ll, p1, p2, lab = _case(rng, S, H, hap)
        w_ll, w_post, w_tot, w_total, w_best = po.log_sample_posteriors(ll, p1, p2, lab, S, haploid=hap)
        g_ll, g_post, g_tot, g_total = engine.posteriors(ll, p1, p2, lab, S, haploid=hap)
        assert np.array_equal(g_ll, w_ll)  # in-place clamp semantics
        np.testing.assert_allclose(g_post, w_post, rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(g_tot, w_tot, rtol=RTOL, atol=ATOL)
        assert abs(g_total - w_total) <= ATOL + RTOL * abs(w_total)


@pytest.mark.parametrize("S,H,reads", [(1, 24, 40), (3, 48, 30), (2, 120, 25), (1, 300, 12)])
def test_posteriors_many_haplotypes(engine, S, H, reads):
    """Loci far beyond the synthetic configurations (the reference allows up to 1 000 haplotypes,
    src/seq_stutter_genotyper.cpp:622): the kernel's tables no longer fit its shared-memory budget and it falls back to
    fewer tables / the plain loop.  Same values, same tolerance."""
    rng = np.random.default_rng(100 + H)
    lab = np.repeat(np.arange(S), reads).astype(np.int32)
    R = len(lab)
    ll = -rng.exponential(30, size=(R, H))
    ll[rng.random((R, H)) < 0.03] = -700
    hp = rng.integers(0, 3, size=R)
    p1 = np.where(hp == 0, -1e-6, np.where(hp == 1, -1000.0, 0.0))
    p2 = np.where(hp == 0, -1000.0, np.where(hp == 1, -1e-6, 0.0))
    for haploid in (False, True):
        w_ll, w_post, w_tot, w_total, _ = po.log_sample_posteriors(ll, p1, p2, lab, S, haploid=haploid)
        g_ll, g_post, g_tot, g_total = engine.posteriors(ll, p1, p2, lab, S, haploid=haploid)
        assert np.array_equal(g_ll, w_ll)
        np.testing.assert_allclose(g_post, w_post, rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(g_tot, w_tot, rtol=RTOL, atol=ATOL)
        assert abs(g_total - w_total) <= ATOL + RTOL * abs(w_total)


def test_job_posteriors_multi_sample_trio(engine):
    """Resident-job path with three samples per locus (the trio configuration of BASELINE.json configs[1]):
    Viterbi LLs of the pooled reads -> per-sample posteriors, against the oracle locus by locus."""
    import synth
    rng = np.random.default_rng(2026)
    b = synth.make_pair_batch(31, n_loci=40, n_lo=30, n_hi=160, reads_lo=4, reads_hi=9, haps_lo=2, haps_hi=6, weird=0.05)
    lrb = b["locus_read_begin"].astype(np.int64)
    lhb = b["locus_hap_begin"].astype(np.int64)
    lsb, pool, lab, p1, p2, nsamp, hap = [0], [], [], [], [], [], []
    for l in range(40):
        P = int(lrb[l + 1] - lrb[l])
        S = 3
        for s in range(S):                       # sample-major reads, each pointing at a pooled read of the locus
            k = int(rng.integers(2, 12))
            pool += [int(x) for x in rng.integers(0, P, size=k)]
            lab += [s] * k
            hp = rng.integers(0, 3, size=k)
            p1 += [(-1e-6, -1000.0, 0.0)[h] for h in hp]
            p2 += [(-1000.0, -1e-6, 0.0)[h] for h in hp]
        lsb.append(len(pool))
        nsamp.append(S)
        hap.append(1 if l % 7 == 0 else 0)
    post = dict(locus_sread_begin=np.array(lsb, np.uint32), pool_index=np.array(pool, np.uint32),
                sample_label=np.array(lab, np.int32), log_p1=np.array(p1), log_p2=np.array(p2),
                locus_n_samples=np.array(nsamp, np.uint32), locus_haploid=np.array(hap, np.uint8))
    job = engine.create_job(b, post)
    job.run()
    ll, gpost, gtot = job.download()
    job.close()
    want_ll, _ = po.viterbi_batch(b)
    assert np.array_equal(ll, want_ll)
    off = np.concatenate([[0], np.cumsum((lhb[1:] - lhb[:-1]) * (lrb[1:] - lrb[:-1]))])
    po_off, t_off = 0, 0
    for l in range(40):
        H, P = int(lhb[l + 1] - lhb[l]), int(lrb[l + 1] - lrb[l])
        mat = want_ll[off[l]:off[l + 1]].reshape(P, H)
        r0, r1 = lsb[l], lsb[l + 1]
        _cl, wpost, wtot, _total, _best = po.log_sample_posteriors(mat[post["pool_index"][r0:r1]], post["log_p1"][r0:r1],
                                                                   post["log_p2"][r0:r1], post["sample_label"][r0:r1], 3,
                                                                   haploid=bool(hap[l]))
        np.testing.assert_allclose(gpost[po_off:po_off + 3 * H * H], wpost.ravel(), rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(gtot[t_off:t_off + 3], wtot, rtol=RTOL, atol=ATOL)
        po_off += 3 * H * H
        t_off += 3


def _fixture_batch(cases, prune):
    """The pruning fixture's loci as one ltr_posteriors_batch call (pools = reads)."""
    lhb, lrb, lsb, ll, pool, lab, p1, p2, ns, hap = [0], [0], [0], [], [], [], [], [], [], []
    for c in cases:
        S, H, R = c["S"], c["H"], c["R"]
        lhb.append(lhb[-1] + H)
        lrb.append(lrb[-1] + R)
        lsb.append(lsb[-1] + R)
        ll.append(gu.unhex(c["ll"]))
        pool.append(np.arange(R))
        lab.append(np.repeat(np.arange(S), c["reads_per_sample"]))
        p1.append(gu.unhex(c["log_p1"]))
        p2.append(gu.unhex(c["log_p2"]))
        ns.append(S)
        hap.append(1 if c["haploid"] else 0)
    post = dict(locus_sread_begin=np.array(lsb, np.uint32), pool_index=np.concatenate(pool).astype(np.uint32),
                sample_label=np.concatenate(lab).astype(np.int32), log_p1=np.concatenate(p1), log_p2=np.concatenate(p2),
                locus_n_samples=np.array(ns, np.uint32), locus_haploid=np.array(hap, np.uint8), prune_uncalled=prune)
    return np.array(lhb, np.uint32), np.array(lrb, np.uint32), np.concatenate(ll), post


def test_batch_posteriors_with_removal_of_uncalled_alleles_match_reference(engine):
    """ltr_posteriors_batch with prune_uncalled on the LL matrices the reference's SeqStutterGenotyper::genotype held
    (tests/golden/pruning.json, recorded by oracle/_ref/ltr_ref_trace): the device drops the same alleles
    (src/seq_stutter_genotyper.cpp:250-311, 636-645) and its second-pass posteriors equal the reference's (1e-12)."""
    cases = gu.load("pruning")
    lhb, lrb, ll, post = _fixture_batch(cases, True)
    got_post, got_tot, kept = engine.posteriors_batch(lhb, lrb, ll, post)
    po_off, to_off = 0, 0
    for i, c in enumerate(cases):
        S, H = c["S"], c["H"]
        mask = kept[lhb[i]:lhb[i + 1]]
        assert list(np.nonzero(mask)[0]) == c["kept"], c["name"]
        K = len(c["kept"])
        np.testing.assert_allclose(got_post[po_off:po_off + S * K * K], gu.unhex(c["out_post"]), rtol=1e-12, atol=1e-10,
                                   err_msg=c["name"])
        np.testing.assert_allclose(got_tot[to_off:to_off + S], gu.unhex(c["out_totals"]), rtol=1e-12, atol=1e-10)
        po_off += S * H * H
        to_off += S
    # without the removal the same call gives the first pass
    lhb, lrb, ll, post = _fixture_batch(cases, False)
    got_post, got_tot, kept = engine.posteriors_batch(lhb, lrb, ll, post)
    assert kept.all()
    po_off = 0
    for c in cases:
        n = c["S"] * c["H"] * c["H"]
        np.testing.assert_allclose(got_post[po_off:po_off + n], gu.unhex(c["first_post"]), rtol=1e-12, atol=1e-10)
        po_off += n


def test_mate_pairs_sum_their_rows(engine):
    """second_mate (src/seq_stutter_genotyper.cpp:494, 546-559): the LL rows of a read and its mate are replaced by their
    sum, accumulated in read order along runs of flagged reads, before the posteriors are formed."""
    rng = np.random.default_rng(91)
    lhb, lrb, lsb, lls, pools, labs, p1s, p2s, nss, mates, want = [0], [0], [0], [], [], [], [], [], [], [], []
    for t in range(12):
        S, H = int(rng.integers(1, 4)), int(rng.integers(1, 6))
        rps = [int(x) for x in rng.integers(2, 9, size=S)]
        R = sum(rps)
        P = int(rng.integers(1, R + 1))                      # pooled reads
        pool_ll = -rng.exponential(20, size=(P, H)) - 1
        pool = rng.integers(0, P, size=R)
        lab = np.repeat(np.arange(S), rps)
        mate = np.zeros(R, np.uint8)
        for r in range(1, R):
            if lab[r] == lab[r - 1] and rng.random() < 0.35:
                mate[r] = 1
        p1 = np.log(rng.uniform(0.05, 1.0, size=R))
        p2 = np.log(rng.uniform(0.05, 1.0, size=R))
        rows = pool_ll[pool].copy()                          # the reference's per-read rows, then :546-559 literally
        for i in range(R):
            if mate[i]:
                tot = rows[i - 1] + rows[i]
                rows[i - 1] = tot
                rows[i] = tot
        _cl, post, tot, _t, _b = po.log_sample_posteriors(rows, p1, p2, lab.astype(np.int32), S)
        want.append((post.ravel(), tot))
        lhb.append(lhb[-1] + H); lrb.append(lrb[-1] + P); lsb.append(lsb[-1] + R)
        lls.append(pool_ll.ravel()); pools.append(pool); labs.append(lab); p1s.append(p1); p2s.append(p2); nss.append(S)
        mates.append(mate)
    post_b = dict(locus_sread_begin=np.array(lsb, np.uint32), pool_index=np.concatenate(pools).astype(np.uint32),
                  sample_label=np.concatenate(labs).astype(np.int32), log_p1=np.concatenate(p1s), log_p2=np.concatenate(p2s),
                  locus_n_samples=np.array(nss, np.uint32), locus_haploid=None, second_mate=np.concatenate(mates))
    got_post, got_tot, _kept = engine.posteriors_batch(np.array(lhb, np.uint32), np.array(lrb, np.uint32), np.concatenate(lls),
                                                       post_b)
    po_off = to_off = 0
    for (wp, wt) in want:
        np.testing.assert_allclose(got_post[po_off:po_off + len(wp)], wp, rtol=1e-12, atol=1e-10)
        np.testing.assert_allclose(got_tot[to_off:to_off + len(wt)], wt, rtol=1e-12, atol=1e-10)
        po_off += len(wp)
        to_off += len(wt)
