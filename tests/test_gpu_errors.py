"""GPU: error behaviour of the C ABI -- bad input is answered with an error code (the reference exits or asserts),
never with a crash, and results of valid calls are unaffected afterwards."""
import ctypes as C

import numpy as np
import pytest

import synth
from longtr_b200 import abi
from longtr_b200.engine import LongTRError

pytestmark = pytest.mark.gpu


def test_invalid_batches_are_rejected(engine):
    good = synth.make_pair_batch(3, n_loci=4)
    bad = dict(good)
    bad["read_off"] = good["read_off"].copy()
    bad["read_off"][2] = bad["read_off"][1]          # an empty read
    with pytest.raises(LongTRError, match="invalid"):
        engine.viterbi_ll(bad)
    bad = dict(good)
    bad["locus_hap_begin"] = good["locus_hap_begin"].copy()
    bad["locus_hap_begin"][1], bad["locus_hap_begin"][2] = good["locus_hap_begin"][2], good["locus_hap_begin"][1]
    vb, keep = abi.make_viterbi_batch(bad)                       # decreasing locus offsets
    p = abi.make_params()
    out = np.zeros(4 * abi.ll_size(good) + 64)
    assert engine.lib.ltr_viterbi_ll(engine.ctx, C.byref(p), C.byref(vb), abi.ptr(out, abi._dp), None) == -3
    with pytest.raises(LongTRError, match="invalid"):
        engine.viterbi_ll(good, indel_flank_len=99)
    lib = engine.lib
    assert lib.ltr_viterbi_ll(engine.ctx, None, None, None, None) == -3
    assert lib.ltr_job_run(engine.ctx, None) == -3
    # the context is still usable
    from oracle import pyoracle as po
    got, _ = engine.viterbi_ll(good)
    want, _ = po.viterbi_batch(good)
    assert np.array_equal(got, want)


def test_posterior_argument_checks(engine):
    ll = np.zeros((3, 2))
    p = np.full(3, -0.7)
    with pytest.raises(LongTRError, match="invalid"):
        engine.posteriors(ll, p, p, np.array([0, 1, 5], np.int32), 2)   # label out of range
    with pytest.raises(LongTRError, match="invalid"):
        engine.posteriors(ll, p, p, np.array([0, 0, 0], np.int32), 0)   # no samples


def test_flat_locus_errors(engine):
    loc = synth.make_locus(5, n_reads=2)
    loc["reads"][0]["cigar"] = "12Q"                                    # unknown CIGAR operation
    L, keep = synth.to_flat(loc)
    with pytest.raises(LongTRError, match="invalid"):
        engine.process_reads_flat(L, 2, len(loc["alleles"]))
    loc = synth.make_locus(6, n_reads=2, homopolymer=True)
    loc["alleles"] = loc["alleles"] + [""]                              # <DEL> allele on the homopolymer path
    L, keep = synth.to_flat(loc, switch_old_align_len=20)
    with pytest.raises(LongTRError, match="unsupported"):
        engine.process_reads_flat(L, 2, len(loc["alleles"]))


def test_context_creation_on_a_missing_device():
    lib = abi.load()
    ctx = C.c_void_p()
    assert lib.ltr_ctx_create(99, C.byref(ctx)) == -1 and not ctx.value
