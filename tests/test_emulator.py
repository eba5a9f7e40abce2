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
    lib.ltr_emu_viterbi_batch_band.argtypes = [C.POINTER(abi.ViterbiBatch), C.POINTER(abi.Params), C.c_int, C.c_int,
                                               C.c_int, abi._dp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.ltr_emu_viterbi_batch_band.restype = C.c_int
    return lib


def run_band(emu, b, params, kmax, band_w):
    """Emulated band kernel + collect + stream kernels; returns (LL, [band pairs, uncertified, appended tasks])."""
    vb, keep = abi.make_viterbi_batch(b)
    p = abi.make_params(params)
    out = np.full(abi.ll_size(b), 123.0)
    nf = C.c_uint64(0)
    bs = (C.c_uint64 * 3)()
    rc = emu.ltr_emu_viterbi_batch_band(C.byref(vb), C.byref(p), kmax, 1, band_w, abi.ptr(out, abi._dp), C.byref(nf), bs)
    assert rc == 0
    return out, list(bs)


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


@pytest.mark.parametrize("seed,kw,kmax,params", [c for c in CASES if c[0] in (1, 2, 3, 6)])
@pytest.mark.parametrize("band_w", [0, 2, 8, 40])
def test_band_emulator_matches_oracle(emu, seed, kw, kmax, params, band_w):
    """Banded anti-diagonal kernel (band_core.cuh) + re-run of the uncertified pairs == oracle, bit for bit, whatever
    the requested margin (0 = automatic)."""
    b = synth.make_pair_batch(seed, **kw)
    want, _ = po.viterbi_batch(b, aln_params=params)
    out, stats = run_band(emu, b, params, kmax, band_w)
    assert np.array_equal(out, want)
    if band_w in (0, 2, 8):
        assert stats[0] > 0  # the case does exercise the band kernel


@pytest.mark.parametrize("band_w", [70, 110, 150, 0])
def test_band_wide_classes(emu, band_w):
    """Whole-warp band classes (W = 256, 384, 512) on long noisy pairs with ONT-like parameters (0 = the automatic
    margin, which grows with the haplotype length for such parameters)."""
    b = synth.make_pair_batch(61, n_loci=3, n_lo=500, n_hi=900, reads_hi=3, haps_hi=3, sub=0.02, indel=0.03)
    want, _ = po.viterbi_batch(b, aln_params=ONT, n_threads=4)
    out, stats = run_band(emu, b, ONT, 16, band_w)
    assert np.array_equal(out, want)
    assert stats[0] > 0


@pytest.mark.parametrize("seed", range(12))
def test_band_certificate_adversarial(emu, seed):
    """Reads whose best alignment leaves a narrow band (block insertions / deletions / duplications of up to 30 bases
    in low-complexity haplotypes) under three parameter sets: the certificate F_band > -g(|de| + 2w) must reject
    every pair whose banded score is not the reference's score."""
    rng = np.random.default_rng(777 + seed)
    params = [None, ONT, (-0.5, -0.4, -0.25, -0.3, -0.01, -2.0, -1.5)][seed % 3]
    lhb, lrb, hoff, roff, hb, rb = [0], [0], [0], [0], [], []
    for _l in range(8):
        n = int(rng.integers(60, 200))
        hap = synth.rand_seq(rng, n + 60)
        if rng.random() < 0.5:  # low complexity: off-diagonal paths are competitive
            motif = synth.rand_seq(rng, int(rng.integers(1, 5)))
            hap = hap[:30] + (motif * 400)[:n] + hap[30 + n:]
        hb.append(hap)
        hoff.append(hoff[-1] + len(hap))
        for _r in range(6):
            core = list(hap[30:30 + n])
            for _e in range(int(rng.integers(0, 4))):
                k, pos = int(rng.integers(1, 30)), int(rng.integers(0, max(1, len(core))))
                if rng.random() < 0.5:
                    core[pos:pos] = list(synth.rand_seq(rng, k)) if rng.random() < 0.5 else core[max(0, pos - k):pos]
                else:
                    del core[pos:pos + k]
            for _e in range(int(rng.integers(0, 4))):
                if core:
                    core[int(rng.integers(0, len(core)))] = "ACGT"[int(rng.integers(0, 4))]
            s = "".join(core)
            if len(s) < 2:
                s = "AC"
            rb.append(s)
            roff.append(roff[-1] + len(s))
        lhb.append(len(hb))
        lrb.append(len(rb))
    b = dict(locus_hap_begin=np.array(lhb, np.uint32), locus_read_begin=np.array(lrb, np.uint32),
             hap_off=np.array(hoff, np.uint32), read_off=np.array(roff, np.uint32),
             hap_bytes=np.frombuffer("".join(hb).encode(), np.uint8).copy(),
             read_bytes=np.frombuffer("".join(rb).encode(), np.uint8).copy())
    want, _ = po.viterbi_batch(b, aln_params=params, n_threads=4)
    n_band = n_unc = 0
    for band_w in (1, 3, 6):
        out, stats = run_band(emu, b, params, 16, band_w)
        assert np.array_equal(out, want)
        n_band += stats[0]
        n_unc += stats[1]
    assert n_band > 0 and 0 < n_unc < n_band  # both outcomes of the certificate occur


CHEAP = (-0.5, -0.4, -0.25, -0.3, -0.01, -0.6, -0.55)  # opening a gap barely dearer than extending it
P4 = (-0.3, -0.05, -0.3, -0.05, -0.001, -1.2, -1.2)


@pytest.mark.parametrize("seed", range(8))
def test_band_certificate_out_and_back(emu, seed):
    """The certificate charges a chain that leaves the band two gap openings (out and back).  Reads made of noisy tandem
    repeats with a block deleted at one place and (almost) the same block re-inserted elsewhere are exactly the inputs
    whose best path makes two long gap runs far from the diagonal; margins 1-7 and the automatic one, four parameter
    sets including one whose gap opening is barely dearer than its extension."""
    rng = np.random.default_rng(31337 + seed)
    params = [None, ONT, CHEAP, P4][seed % 4]
    lhb, lrb, hoff, roff, hb, rb = [0], [0], [0], [0], [], []
    for _l in range(6):
        n = int(rng.integers(70, 260))
        motif = synth.rand_seq(rng, int(rng.integers(1, 7)))
        core = list((motif * 400)[:n])
        for i in range(n):
            if rng.random() < 0.06:
                core[i] = "ACGT"[int(rng.integers(0, 4))]
        hap = synth.rand_seq(rng, 30) + "".join(core) + synth.rand_seq(rng, 30)
        hb.append(hap)
        hoff.append(hoff[-1] + len(hap))
        for _r in range(8):
            c = list(core)
            for _e in range(int(rng.integers(1, 3))):
                k = int(rng.integers(1, 40))
                a = int(rng.integers(0, max(1, len(c) - k)))
                seg = c[a:a + k]
                del c[a:a + k]
                k2 = max(0, k + int(rng.integers(-3, 4)))
                bpos = int(rng.integers(0, len(c) + 1))
                c[bpos:bpos] = seg[:k2] if rng.random() < 0.5 else list((motif * 50)[:k2])
            for _e in range(int(rng.integers(0, 3))):
                if c:
                    c[int(rng.integers(0, len(c)))] = "ACGT"[int(rng.integers(0, 4))]
            s = "".join(c)
            if len(s) < 2:
                s = "AC"
            rb.append(s)
            roff.append(roff[-1] + len(s))
        lhb.append(len(hb))
        lrb.append(len(rb))
    b = dict(locus_hap_begin=np.array(lhb, np.uint32), locus_read_begin=np.array(lrb, np.uint32),
             hap_off=np.array(hoff, np.uint32), read_off=np.array(roff, np.uint32),
             hap_bytes=np.frombuffer("".join(hb).encode(), np.uint8).copy(),
             read_bytes=np.frombuffer("".join(rb).encode(), np.uint8).copy())
    want, _ = po.viterbi_batch(b, aln_params=params, n_threads=4)
    n_band = n_unc = 0
    for band_w in (1, 2, 4, 7, 0):
        out, stats = run_band(emu, b, params, 16, band_w)
        assert np.array_equal(out, want)
        n_band += stats[0]
        n_unc += stats[1]
    assert n_band > 0 and 0 < n_unc < n_band


def test_band_geometry_properties():
    """band_geometry / band_cells (band_core.cuh) against brute force, through the emulator library's helpers."""
    lib = C.CDLL(os.path.join(HERE, "emu", "libltr_emu.so"))
    lib.ltr_emu_band_geometry.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                          C.POINTER(C.c_uint64)]
    rng = np.random.default_rng(5)
    for _ in range(400):
        n, m = int(rng.integers(2, 400)), int(rng.integers(2, 400))
        W = 16 * int(rng.choice([2, 3, 4, 6, 8]))
        dlo, w, cells = C.c_int(), C.c_int(), C.c_uint64()
        lib.ltr_emu_band_geometry(n, m, W, C.byref(dlo), C.byref(w), C.byref(cells))
        de = m - n
        if W - 1 - abs(de) < 0 or w.value < 0:
            assert w.value < 0 and W - 1 - abs(de) <= 1  # no margin left (the even alignment may cost one diagonal)
            continue
        dhi = dlo.value + W - 1
        assert dlo.value % 2 == 0
        assert dlo.value <= min(0, de) - w.value and dhi >= max(0, de) + w.value and w.value >= 0
        assert w.value >= (W - 1 - abs(de)) // 2 - 1  # at most one diagonal lost to the even alignment
        i = np.arange(1, n)[:, None]
        j = np.arange(1, m)[None, :]
        assert cells.value == int(np.sum((j - i >= dlo.value) & (j - i <= dhi)))


def test_band_policy_margins(emu):
    """band_policy / band_margin_needed (viterbi_host.h): off when the certificate's parameter condition fails, explicit
    margins honoured, automatic margin 7 for the Dindel defaults and growing with the haplotype for ONT-like parameters."""
    emu.ltr_emu_band_margin.argtypes = [C.POINTER(abi.Params), C.c_int, C.c_int]
    emu.ltr_emu_band_margin.restype = C.c_int
    dindel, ont = abi.make_params(None), abi.make_params(ONT)
    assert emu.ltr_emu_band_margin(C.byref(dindel), -1, 200) == -1            # switched off
    assert emu.ltr_emu_band_margin(C.byref(dindel), 33, 200) == 33            # explicit
    assert emu.ltr_emu_band_margin(C.byref(dindel), 0, 200) == 7              # automatic, HiFi-like parameters
    assert emu.ltr_emu_band_margin(C.byref(dindel), 0, 600) in (7, 8)
    w300, w760 = emu.ltr_emu_band_margin(C.byref(ont), 0, 300), emu.ltr_emu_band_margin(C.byref(ont), 0, 760)
    assert 25 <= w300 < w760 and 55 <= w760 <= 70                            # ONT-like: margin follows the expected errors
    # (budget factor 0.45 since the second band round exists: a narrow first attempt costs little when it fails)
    bad = abi.make_params((-1.0, -0.4, -1.0, -0.4, 0.01, -10.0, -10.0))       # a positive parameter: no certificate
    assert emu.ltr_emu_band_margin(C.byref(bad), 0, 200) == -1
    odd = abi.make_params(ODD)                                                # |I2I| < |D2D| ... condition of section 4 fails?
    assert emu.ltr_emu_band_margin(C.byref(odd), 0, 200) in (-1,) + tuple(range(2, 256))


@pytest.mark.parametrize("seed,kw,kmax,params", CASES)
@pytest.mark.parametrize("band_w", [-1, 0, 3, 90])
def test_plan_invariants(emu, seed, kw, kmax, params, band_w):
    """make_plan (viterbi_host.h): read de-duplication, length order of the distinct reads, and the partition of all
    (haplotype, distinct read) pairs into band tasks and stream tasks -- each pair exactly once, classes consistent."""
    emu.ltr_emu_plan_check.argtypes = [C.POINTER(abi.ViterbiBatch), C.POINTER(abi.Params), C.c_int, C.c_int,
                                       C.POINTER(C.c_uint64)]
    emu.ltr_emu_plan_check.restype = C.c_int
    b = synth.make_pair_batch(seed, **kw)
    # duplicate some reads inside their locus so that the de-duplication has something to do
    vb, keep = abi.make_viterbi_batch(b)
    p = abi.make_params(params)
    counts = (C.c_uint64 * 3)()
    assert emu.ltr_emu_plan_check(C.byref(vb), C.byref(p), kmax, band_w, counts) == 0
    if band_w < 0:
        assert counts[1] == 0


def test_plan_invariants_with_duplicate_reads(emu):
    emu.ltr_emu_plan_check.argtypes = [C.POINTER(abi.ViterbiBatch), C.POINTER(abi.Params), C.c_int, C.c_int,
                                       C.POINTER(C.c_uint64)]
    emu.ltr_emu_plan_check.restype = C.c_int
    from longtr_b200 import workloads
    w = workloads.generate(3, 150)   # config-3 loci: about half of the pooled reads are duplicates after trimming
    b, _ = w.subset(150)
    vb, keep = abi.make_viterbi_batch(b)
    p = abi.make_params(w.aln_params)
    counts = (C.c_uint64 * 3)()
    assert emu.ltr_emu_plan_check(C.byref(vb), C.byref(p), 16, 0, counts) == 0
    n_reads = int(b["locus_read_begin"][-1])
    assert counts[0] < 0.8 * n_reads and counts[1] > 10 * counts[2]
    w.close()


@pytest.mark.parametrize("seed", range(15))
def test_pathological_batches(emu, seed):
    """Loci without reads or without haplotypes, haplotypes of 0..62 bases next to 700-base ones, duplicated / one-base /
    1 500-base reads: plan invariants hold and the emulated kernels (band + stream) equal the oracle."""
    emu.ltr_emu_plan_check.argtypes = [C.POINTER(abi.ViterbiBatch), C.POINTER(abi.Params), C.c_int, C.c_int,
                                       C.POINTER(C.c_uint64)]
    emu.ltr_emu_plan_check.restype = C.c_int
    b = synth.make_pathological_batch(seed)
    vb, keep = abi.make_viterbi_batch(b)
    p = abi.make_params(None)
    for band_w in (-1, 0, 3):
        counts = (C.c_uint64 * 3)()
        assert emu.ltr_emu_plan_check(C.byref(vb), C.byref(p), 16, band_w, counts) == 0
    want, _ = po.viterbi_batch(b)
    out, _stats = run_band(emu, b, None, 16, 0)
    assert np.array_equal(out, want)


def _device_plan_check(emu, b, params, kmax, band_w):
    emu.ltr_emu_device_plan_check.argtypes = [C.POINTER(abi.ViterbiBatch), C.POINTER(abi.Params), C.c_int, C.c_int,
                                              C.POINTER(C.c_uint64)]
    emu.ltr_emu_device_plan_check.restype = C.c_int
    vb, keep = abi.make_viterbi_batch(b)
    p = abi.make_params(params)
    counts = (C.c_uint64 * 3)()
    return emu.ltr_emu_device_plan_check(C.byref(vb), C.byref(p), kmax, band_w, counts), list(counts)


@pytest.mark.parametrize("seed,kw,kmax,params", CASES)
@pytest.mark.parametrize("band_w", [-1, 0, 3, 90])
def test_device_plan_matches_host_plan(emu, seed, kw, kmax, params, band_w):
    """plan_device.cuh (the plan kernels' per-item functions, run serially on the host) gives make_plan's numbering of the
    distinct reads, offsets, bytes, read map, statistics and task sets."""
    b = synth.make_pair_batch(seed, **kw)
    rc, _ = _device_plan_check(emu, b, params, kmax, band_w)
    assert rc == 0


def test_device_plan_with_duplicate_reads(emu):
    from longtr_b200 import workloads
    w = workloads.generate(3, 300)
    b, _ = w.subset(300)
    rc, counts = _device_plan_check(emu, b, w.aln_params, 16, 0)
    assert rc == 0
    assert counts[0] < 0.8 * int(b["locus_read_begin"][-1])
    w4 = workloads.generate(4, 12)
    b4, _ = w4.subset(12)
    rc, counts = _device_plan_check(emu, b4, w4.aln_params, 16, 0)
    assert rc == 0
    w.close()
    w4.close()


@pytest.mark.parametrize("seed", range(15))
def test_device_plan_pathological_batches(emu, seed):
    b = synth.make_pathological_batch(seed)
    for band_w in (-1, 0, 3):
        rc, _ = _device_plan_check(emu, b, None, 16, band_w)
        assert rc == 0


def test_device_plan_flags_malformed_read_offsets(emu):
    b = synth.make_pair_batch(11, n_loci=6)
    off = b["read_off"].copy()
    off[3] = off[2]  # an empty read
    b2 = dict(b, read_off=off)
    rc, _ = _device_plan_check(emu, b2, None, 16, 0)
    assert rc == -1  # both plans reject it


def test_band_second_round_certifies(emu):
    """Second band round (band_retry_class): an uncertified pair whose banded score F_band lies above the threshold of a
    wider class is re-run there and MUST come back certified (the emulator aborts otherwise); results stay the oracle's.
    Narrow margins on long noisy pairs make sure the round is exercised."""
    emu.ltr_emu_band_retried.restype = C.c_uint64
    emu.ltr_emu_band_retried.argtypes = [C.c_int]
    emu.ltr_emu_band_retried(1)
    cases = [(ONT, synth.make_pair_batch(61, n_loci=3, n_lo=500, n_hi=900, reads_hi=3, haps_hi=3, sub=0.02, indel=0.03)),
             (ONT, synth.make_pair_batch(62, n_loci=4, n_lo=200, n_hi=500, reads_hi=4, haps_hi=3, sub=0.01, indel=0.015)),
             (None, synth.make_pair_batch(63, n_loci=6, n_lo=100, n_hi=300, reads_hi=5, haps_hi=4, sub=0.01, indel=0.02))]
    for params, b in cases:
        want, _ = po.viterbi_batch(b, aln_params=params, n_threads=4)
        for band_w in (1, 3, 12, 30):
            out, stats = run_band(emu, b, params, 16, band_w)
            assert np.array_equal(out, want)
    assert emu.ltr_emu_band_retried(0) > 20
