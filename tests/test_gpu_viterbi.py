"""GPU parity: CUDA Viterbi path (through the C ABI) vs the CPU oracle, bit for bit."""
import numpy as np
import pytest

import synth
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

ONT = (-1.0, -0.458675, -1.0, -0.458675, -0.0202027, -4.60517, -4.60517)
ODD = (-0.7, -0.61, -0.35, -1.3, -0.013, -3.9, -4.4)


@pytest.mark.parametrize("seed,kw,params", [
    (1, dict(n_loci=60), None),
    (2, dict(n_loci=40, n_lo=20, n_hi=400), None),
    (3, dict(n_loci=40, n_lo=200, n_hi=520, reads_hi=4, haps_hi=3), ONT),
    (4, dict(n_loci=40, n_lo=10, n_hi=150, weird=0.3), ODD),
    (5, dict(n_loci=40, n_lo=1, n_hi=40, weird=0.3), None),
    (6, dict(n_loci=6, n_lo=600, n_hi=1100, reads_hi=3, haps_hi=3, sub=0.02, indel=0.03), ONT),
])
def test_viterbi_bit_exact(engine, seed, kw, params):
    b = synth.make_pair_batch(seed, **kw)
    want, _cells = po.viterbi_batch(b, aln_params=params, n_threads=4)
    got, st = engine.viterbi_ll(b, aln_params=params)
    assert st.n_pairs == len(want)
    bad = np.nonzero(got != want)[0]
    assert len(bad) == 0, (bad[:10], got[bad[:10]], want[bad[:10]])


def test_empty_and_single(engine):
    b = synth.make_pair_batch(7, n_loci=1, reads_lo=1, reads_hi=1, haps_lo=1, haps_hi=1)
    want, _ = po.viterbi_batch(b)
    got, _ = engine.viterbi_ll(b)
    assert np.array_equal(got, want)
    empty = dict(locus_hap_begin=np.zeros(1, np.uint32), locus_read_begin=np.zeros(1, np.uint32),
                 hap_off=np.zeros(1, np.uint32), read_off=np.zeros(1, np.uint32),
                 hap_bytes=np.zeros(0, np.uint8), read_bytes=np.zeros(0, np.uint8))
    got, _ = engine.viterbi_ll(empty)
    assert len(got) == 0


def test_job_rerun_is_idempotent(engine):
    b = synth.make_pair_batch(8, n_loci=50)
    want, _ = po.viterbi_batch(b)
    job = engine.create_job(b)
    for _ in range(3):
        job.run()
        ll, _, _ = job.download()
        assert np.array_equal(ll, want)
    job.close()


@pytest.mark.parametrize("seed", range(15))
def test_pathological_batches(engine, seed):
    """Loci without reads or haplotypes, tiny and long haplotypes side by side, duplicated / one-base / huge reads."""
    b = synth.make_pathological_batch(seed)
    want, _ = po.viterbi_batch(b)
    got, st = engine.viterbi_ll(b)
    assert np.array_equal(got, want)
