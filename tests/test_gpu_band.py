"""GPU parity of the banded anti-diagonal kernel (band_kernel.cu): whatever margin is requested -- off, automatic,
narrow, wide -- the C ABI returns the oracle's bits; the statistics show which route the pairs took."""
import numpy as np
import pytest

import synth
from longtr_b200 import workloads
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

ONT = (-1.0, -0.458675, -1.0, -0.458675, -0.0202027, -4.60517, -4.60517)


@pytest.fixture()
def band_engine(engine):
    yield engine
    engine.set_band(0)  # back to the default for the other tests of the session


@pytest.mark.parametrize("seed,kw,params", [
    (1, dict(n_loci=60), None),
    (2, dict(n_loci=40, n_lo=20, n_hi=400), None),
    (3, dict(n_loci=40, n_lo=200, n_hi=520, reads_hi=4, haps_hi=3), ONT),
    (6, dict(n_loci=6, n_lo=600, n_hi=1100, reads_hi=3, haps_hi=3, sub=0.02, indel=0.03), ONT),
])
@pytest.mark.parametrize("band_w", [-1, 0, 1, 5, 24, 60, 110, 180])
def test_band_bit_exact(band_engine, seed, kw, params, band_w):
    b = synth.make_pair_batch(seed, **kw)
    want, _cells = po.viterbi_batch(b, aln_params=params, n_threads=4)
    band_engine.set_band(band_w)
    got, st = band_engine.viterbi_ll(b, aln_params=params)
    bad = np.nonzero(got != want)[0]
    assert len(bad) == 0, (band_w, bad[:10], got[bad[:10]], want[bad[:10]])
    if band_w < 0:
        assert st.n_band_pairs == 0
    else:
        assert st.n_band_uncertified <= st.n_band_pairs <= st.n_pairs_computed


def test_band_statistics_config3(band_engine):
    """Config-3 loci (HiFi-like reads): nearly every pair is banded and certified, and far fewer cells are evaluated."""
    w = workloads.generate(3, 300)
    b, _ = w.subset(300)
    want, _ = po.viterbi_batch(b, aln_params=w.aln_params, n_threads=8)
    band_engine.set_band(-1)
    full, st_full = band_engine.viterbi_ll(b, aln_params=w.aln_params)
    band_engine.set_band(0)
    got, st = band_engine.viterbi_ll(b, aln_params=w.aln_params)
    assert np.array_equal(full, want) and np.array_equal(got, want)
    assert st.n_band_pairs > 0.9 * st.n_pairs_computed
    assert st.n_band_uncertified < 0.05 * st.n_band_pairs
    assert st.n_cells_computed < 0.5 * st_full.n_cells_computed
    w.close()


def test_band_job_rerun_is_idempotent(band_engine):
    b = synth.make_pair_batch(8, n_loci=50)
    want, _ = po.viterbi_batch(b)
    band_engine.set_band(3)
    job = band_engine.create_job(b)
    for _ in range(3):
        job.run()
        ll, _, _ = job.download()
        assert np.array_equal(ll, want)
    job.close()


def test_band_noisy_reads_abandon(band_engine):
    """ONT-like reads against a narrow band: most pairs cannot be certified, the kernel stops trying (performance
    heuristic) and everything is still exact."""
    w = workloads.generate(4, 24)
    b, _ = w.subset(24)
    want, _ = po.viterbi_batch(b, aln_params=w.aln_params, n_threads=8)
    band_engine.set_band(4)
    got, st = band_engine.viterbi_ll(b, aln_params=w.aln_params)
    assert np.array_equal(got, want)
    w.close()
