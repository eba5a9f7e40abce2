"""N4 on the GPU: ltr_em_stutter_train (csrc/em_kernel.cu, one warp per locus, the whole EM loop on the device) against the
reference's EMStutterGenotyper -- recorded (tests/golden/em.json) and live (oracle/_ref/libltr_ref_em.so) -- and against the
restatement oracle/pyem.py.  Tolerance: the kernel's exp / log in double are CUDA's (<= 1 ulp from glibc's), everything else
is the reference's operations in the reference's order, so parameters / log-frequencies / log-likelihoods agree to 1e-9
relative (observed ~1e-13) and the iteration counts are equal."""
import json
import os

import numpy as np
import pytest

import em_cases
import golden_util as gu
from longtr_b200 import abi
from oracle import pyem
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "em.json")
RTOL = 1e-9


def _unhex(xs):
    return np.array([float.fromhex(x) for x in xs])


def _compare(got, k, want_trained, want_iter, want_params, want_ll, want_pri, tag):
    assert bool(got["trained"][k]) == bool(want_trained), tag
    assert got["n_iter"][k] == want_iter, tag
    np.testing.assert_allclose(got["params"][k], want_params, rtol=RTOL, atol=1e-12, err_msg=str(tag))
    np.testing.assert_allclose(got["ll"][k], want_ll, rtol=RTOL, err_msg=str(tag))
    n = min(len(want_pri), got["log_gt_priors"].shape[1])
    np.testing.assert_allclose(got["log_gt_priors"][k][:n], want_pri[:n], rtol=RTOL, atol=1e-10, err_msg=str(tag))


def test_kernel_matches_the_recorded_reference(engine):
    g = json.load(open(GOLD))
    loci = [em_cases.em_locus(c["seed"]) for c in g["cases"]]
    got = abi.em_stutter_train(engine.ctx, loci)
    for k, c in enumerate(g["cases"]):
        _compare(got, k, c["trained"], c["n_iter"], _unhex(c["params"]), float.fromhex(c["ll"]), _unhex(c["log_gt_priors"]),
                 c["seed"])
    short = [em_cases.em_locus(c["seed"]) for c in g["short"]]
    got = abi.em_stutter_train(engine.ctx, short, max_iter=2)
    for k, c in enumerate(g["short"]):
        _compare(got, k, c["trained"], c["n_iter"], _unhex(c["params"]), float.fromhex(c["ll"]), _unhex(c["log_gt_priors"]),
                 c["seed"])
    assert any(not c["trained"] for c in g["short"])


def test_kernel_matches_the_restatement_on_fresh_loci(engine):
    loci = [em_cases.em_locus(7000 + k) for k in range(40)]
    got = abi.em_stutter_train(engine.ctx, loci)
    for k, L in enumerate(loci):
        w = pyem.em_train(L["reads_per_sample"], L["bp_diff"], L["log_p1"], L["log_p2"], L["motif_len"], L["haploid"])
        _compare(got, k, w["trained"], w["n_iter"], w["params"], w["lls"][-1], w["log_gt_priors"], 7000 + k)


@pytest.mark.skipif(not po.ref_em_available(), reason="oracle/_ref/libltr_ref_em.so not built")
def test_kernel_matches_the_reference_on_larger_loci(engine):
    """Hundreds of reads per sample, more allele sizes: the strided loops and the parallel sums of the M step."""
    loci = [em_cases.em_locus(9000 + k, big=True) for k in range(24)]
    got = abi.em_stutter_train(engine.ctx, loci, prior_stride=64)
    for k, L in enumerate(loci):
        w = po.ref_em_train(L["reads_per_sample"], L["bp_diff"], L["log_p1"], L["log_p2"], L["motif_len"], L["haploid"])
        _compare(got, k, w["trained"], w["n_iter"], w["params"], w["lls"][-1], w["log_gt_priors"], 9000 + k)


def test_malformed_batches_are_refused(engine):
    L = em_cases.em_locus(1)
    bad = dict(L, log_p1=[0.5] + list(L["log_p1"][1:]))          # a positive phasing term (the reference asserts)
    with pytest.raises(RuntimeError):
        abi.em_stutter_train(engine.ctx, [bad])
    with pytest.raises(RuntimeError):
        abi.em_stutter_train(engine.ctx, [dict(L, motif_len=0)])
    with pytest.raises(RuntimeError):
        abi.em_stutter_train(engine.ctx, [L], max_iter=0)
    assert abi.em_stutter_train(engine.ctx, [])["params"].shape == (0, 6)
