"""CPU check of the warp-level Viterbi logic: the kernel's lane functions (viterbi_core.cuh),
driven by a host-side lane emulator, against the oracle -- bit for bit, no GPU needed."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import synth
from longtr_b200 import abi
from oracle import pyoracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
ONT = (-1.0, -0.458675, -1.0, -0.458675, -0.0202027, -4.60517, -4.60517)
ODD = (-0.7, -0.61, -0.35, -1.3, -0.013, -3.9, -4.4)


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call([os.path.join(HERE, "emu", "build_emu.sh")])
    lib = C.CDLL(os.path.join(HERE, "emu", "libltr_emu.so"))
    lib.ltr_emu_viterbi_batch.argtypes = [C.POINTER(abi.ViterbiBatch), C.POINTER(abi.Params), C.c_int, C.c_int,
                                          abi._dp, C.POINTER(C.c_uint64)]
    lib.ltr_emu_viterbi_batch.restype = C.c_int
    return lib


CASES = [
    (1, dict(n_loci=30), 16, None),
    (2, dict(n_loci=20, n_lo=20, n_hi=400), 16, None),
    (3, dict(n_loci=20, n_lo=20, n_hi=300), 2, None),            # forces multi-strip hand-off
    (4, dict(n_loci=20, n_lo=10, n_hi=150, weird=0.3), 3, ODD),  # non-integer parameters
    (5, dict(n_loci=20, n_lo=1, n_hi=40, weird=0.3), 1, None),   # tiny haplotypes, K=1 lanes
    (6, dict(n_loci=6, n_lo=100, n_hi=300, sub=0.02, indel=0.04), 4, ONT),
]


@pytest.mark.parametrize("seed,kw,kmax,params", CASES)
@pytest.mark.parametrize("fast", [0, 1])
def test_emulator_matches_oracle(emu, seed, kw, kmax, params, fast):
    b = synth.make_pair_batch(seed, **kw)
    want, _ = po.viterbi_batch(b, aln_params=params)
    vb, keep = abi.make_viterbi_batch(b)
    p = abi.make_params(params)
    out = np.full(len(want), 123.0)
    nf = C.c_uint64(0)
    rc = emu.ltr_emu_viterbi_batch(C.byref(vb), C.byref(p), kmax, fast, abi.ptr(out, abi._dp), C.byref(nf))
    assert rc == 0
    bad = np.nonzero(out != want)[0]
    assert len(bad) == 0, (bad[:10], out[bad[:10]], want[bad[:10]])
    if not fast:
        assert nf.value == 0


@pytest.mark.parametrize("seed", range(9))
def test_final_score_certificate_near_the_bailout_threshold(emu, seed):
    """MODE_FAST certifies 'no row bails out' from the final score alone.  Unrelated reads of
    a length chosen so that scores land around -600 exercise exactly the pairs where that matters: the reference's
    per-row bail-out (-700) and near-threshold scores must come out identical to the oracle."""
    rng = np.random.default_rng(9100 + seed)
    params, per_base = [(None, 2.056), (ONT, 1.732), (ODD, 1.060)][seed % 3]  # -score per base of unrelated pairs
    n0 = int(585.0 / per_base)
    lhb, lrb, hoff, roff, hb, rb = [0], [0], [0], [0], [], []
    for _l in range(6):
        n = n0 + int(rng.integers(-25, 26))
        hap = synth.rand_seq(rng, n + 60)
        hb.append(hap)
        hoff.append(hoff[-1] + len(hap))
        for _r in range(4):
            m = max(2, n + int(rng.integers(-30, 31)))
            s = (hap[30:30 + int(rng.integers(0, 40))] + synth.rand_seq(rng, m))[:m]
            rb.append(s)
            roff.append(roff[-1] + len(s))
        lhb.append(len(hb))
        lrb.append(len(rb))
    b = dict(locus_hap_begin=np.array(lhb, np.uint32), locus_read_begin=np.array(lrb, np.uint32),
             hap_off=np.array(hoff, np.uint32), read_off=np.array(roff, np.uint32),
             hap_bytes=np.frombuffer("".join(hb).encode(), np.uint8).copy(),
             read_bytes=np.frombuffer("".join(rb).encode(), np.uint8).copy())
    want, _ = po.viterbi_batch(b, aln_params=params, n_threads=4)
    assert np.sum((want > -640) & (want < -540)) + np.sum(want == -700) >= 6  # the case is on target
    vb, keep = abi.make_viterbi_batch(b)
    p = abi.make_params(params)
    out = np.full(len(want), 123.0)
    nf = C.c_uint64(0)
    assert emu.ltr_emu_viterbi_batch(C.byref(vb), C.byref(p), 16, 1, abi.ptr(out, abi._dp), C.byref(nf)) == 0
    assert np.array_equal(out, want)
