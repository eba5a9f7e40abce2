"""GPU, BASELINE.json full sizes: size-independent properties of the hot path on the complete synthetic workloads
(config 3: 100 000 HiFi STR loci; config 4: 10 000 VNTR loci; config 5: 50 000 homopolymer loci), plus oracle spot
checks on random loci of the full batch.  Complements the bit-exact small-batch parity tests."""
import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def _locus_slices(b):
    H = np.diff(b["locus_hap_begin"]).astype(np.int64)
    P = np.diff(b["locus_read_begin"]).astype(np.int64)
    off = np.concatenate([[0], np.cumsum(H * P)])
    return H, P, off


def _sub_batch(b, l0, l1):
    lhb, lrb = b["locus_hap_begin"].astype(np.int64), b["locus_read_begin"].astype(np.int64)
    h0, h1, r0, r1 = lhb[l0], lhb[l1], lrb[l0], lrb[l1]
    return dict(locus_hap_begin=(lhb[l0:l1 + 1] - h0).astype(np.uint32), locus_read_begin=(lrb[l0:l1 + 1] - r0).astype(np.uint32),
                hap_off=(b["hap_off"][h0:h1 + 1] - b["hap_off"][h0]).astype(np.uint32),
                read_off=(b["read_off"][r0:r1 + 1] - b["read_off"][r0]).astype(np.uint32),
                hap_bytes=b["hap_bytes"][b["hap_off"][h0]:b["hap_off"][h1]],
                read_bytes=b["read_bytes"][b["read_off"][r0]:b["read_off"][r1]])


@pytest.mark.parametrize("config,n_loci,n_check", [(3, 100000, 2400), (4, 10000, 208)])
def test_full_size_properties(engine, config, n_loci, n_check):
    from longtr_b200 import workloads
    work = workloads.generate(config, n_loci)
    b = work.batch
    job = engine.create_job(b, work.post, aln_params=work.aln_params)
    st = job.run()
    ll, post, tot = job.download()
    assert st.n_pairs == len(ll) and st.n_cells_computed <= 1.1 * st.n_cells  # band attempts that fail count twice
    # (1) idempotence: a second run over the resident job reproduces every bit
    job.run()
    ll2, post2, tot2 = job.download()
    assert np.array_equal(ll, ll2) and np.array_equal(post, post2) and np.array_equal(tot, tot2)
    job.close()
    # (2) range of values: log-likelihoods are <= 0; below -700 only the reference's -1e9 sentinel exists (a final
    #     score slightly under -600 is legitimate: the row test adds the band penalty to every cell of the last row)
    assert np.all(ll <= 0.0)
    low = ll[ll < -600.0]
    assert np.all((low == -1e9) | (low >= -700.0))
    # (3) posteriors are normalised per sample: logsumexp over the H*H entries == 0
    H, P, off = _locus_slices(b)
    poff = np.concatenate([[0], np.cumsum(H * H)])  # one sample per locus in the synthetic sets
    rng = np.random.default_rng(config)
    for l in rng.integers(0, n_loci, size=400):
        v = post[poff[l]:poff[l + 1]]
        assert abs(np.log(np.sum(np.exp(v - v.max()))) + v.max()) < 1e-9
    # (4) shard invariance ("checksum of checksums"): the batch cut into 4 contiguous locus shards, each run as its
    #     own job, gives the same bits as the single job -- what locus sharding over GPUs relies on
    cuts = [0, n_loci // 4, n_loci // 2, 3 * n_loci // 4, n_loci]
    for a, c in zip(cuts[:-1], cuts[1:]):
        part, _ = engine.viterbi_ll(_sub_batch(b, a, c), aln_params=work.aln_params)
        assert np.array_equal(part, ll[off[a]:off[c]])
    # (5) identical trimmed reads of a locus carry identical rows (the plan aligns them once)
    for l in rng.integers(0, n_loci, size=200):
        r0, r1 = int(b["locus_read_begin"][l]), int(b["locus_read_begin"][l + 1])
        seqs = [bytes(b["read_bytes"][b["read_off"][r]:b["read_off"][r + 1]]) for r in range(r0, r1)]
        mat = ll[off[l]:off[l + 1]].reshape(P[l], H[l])
        first = {}
        for i, s in enumerate(seqs):
            if s in first:
                assert np.array_equal(mat[i], mat[first[s]])
            first.setdefault(s, i)
    # (6) bit for bit against the CPU checkers on n_check loci of the full batch (random runs of consecutive loci, all
    #     host threads): the reference's own process_reads (oracle/_ref) where it is built, the restatement otherwise --
    #     and the restatement on every other run anyway
    import os
    threads = os.cpu_count() or 1
    run = 8 if config == 4 else 100
    for k, l0 in enumerate(rng.integers(0, n_loci - run, size=n_check // run)):
        sb = _sub_batch(b, int(l0), int(l0) + run)
        if po.ref_available() and k % 2 == 0:
            want, _sec = po.ref_viterbi_batch(sb, work.aln_params, n_threads=threads)
        else:
            want, _ = po.viterbi_batch(sb, aln_params=work.aln_params, n_threads=threads)
        got = ll[off[l0]:off[l0 + run]]
        bad = np.nonzero(got != want)[0]
        assert len(bad) == 0, (config, int(l0), bad[:5], got[bad[:5]], want[bad[:5]])


def test_full_size_homopolymer_path(engine):
    from longtr_b200 import workloads
    work = workloads.generate_stutter(50000)
    out, st = engine.stutter_ll(work.batch)
    out2, _ = engine.stutter_ll(work.batch)
    assert np.array_equal(out, out2)                      # idempotence
    assert np.all(out <= 1e-10) and np.all(np.isfinite(out))  # the reference asserts LL < TOLERANCE (HapAligner.cpp:231)
    rng = np.random.default_rng(5)
    a = work.batch["locus_allele_begin"].astype(np.int64)
    r = work.batch["locus_read_begin"].astype(np.int64)
    off = np.concatenate([[0], np.cumsum((a[1:] - a[:-1]) * (r[1:] - r[:-1]))])
    # bit for bit against the CPU checkers on 512 random loci (reference where built, restatement otherwise / alternating)
    import os
    from concurrent.futures import ThreadPoolExecutor
    loci = [int(l) for l in rng.integers(0, work.n_loci, size=512)]

    def check(args):
        k, l = args
        (L, keep), (P, H) = work.flat_locus(l)
        which = "ref" if (po.ref_available() and k % 2 == 0) else "oracle"
        want, _seeds, _ = po.process_reads(L, P, H, which=which)
        return l, bool(np.array_equal(out[off[l]:off[l + 1]].reshape(P, H), want))
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
        res = list(ex.map(check, enumerate(loci)))
    assert all(ok for _l, ok in res), [l for l, ok in res if not ok][:10]
