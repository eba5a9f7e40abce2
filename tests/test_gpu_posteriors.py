"""GPU parity: posterior kernel vs the oracle (floating point: exp/log, tolerance stated)."""
import numpy as np
import pytest

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
