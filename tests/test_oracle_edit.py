"""N2 (candidate-haplotype clustering) on the CPU: the restatement (oracle/longtr_oracle_edit.c) against the reference's
own HaplotypeGenerator::needleman_wunsch / greedy_clustering compiled in place (oracle/_ref) and against the golden file
recorded from it; the kernels' per-lane functions (longtr_b200/csrc/edit_core.cuh) through the lane emulator against
both.  Integer arithmetic: every comparison is exact."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

import edit_cases as ec
from longtr_b200 import abi
from oracle import pyoracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "edit.json")


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call([os.path.join(HERE, "emu", "build_emu.sh")])
    lib = C.CDLL(os.path.join(HERE, "emu", "libltr_emu_edit.so"))
    lib.ltr_emu_edit_score.restype = C.c_int32
    lib.ltr_emu_edit_score.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
    lib.ltr_emu_greedy_cluster.restype = C.c_int32
    lib.ltr_emu_greedy_cluster.argtypes = [abi._u8p, abi._u32p, abi._u32p, C.c_int32, C.c_int32, abi._i32p, abi._i32p]
    return lib


def emu_score(emu, a, b, T):
    flag = C.c_int32(0)
    s = emu.ltr_emu_edit_score(a.encode(), len(a), b.encode(), len(b), T, C.byref(flag))
    return int(s), int(flag.value)


def emu_cluster(emu, seqs, T):
    data, off = abi.pack_seqs(seqs)
    items = np.arange(len(seqs), dtype=np.uint32)
    cent = np.full(max(1, len(seqs)), -1, dtype=np.int32)
    n = C.c_int32(0)
    ok = emu.ltr_emu_greedy_cluster(abi.ptr(data, abi._u8p), abi.ptr(off, abi._u32p), abi.ptr(items, abi._u32p), len(seqs),
                                    T, abi.ptr(cent, abi._i32p), C.byref(n))
    return int(ok), cent[:len(seqs)], n.value


def oracle_cluster(seqs, T, which="oracle"):
    data, off = abi.pack_seqs(seqs)
    return po.greedy_cluster(data, off, np.arange(len(seqs), dtype=np.uint32), T, which)


def plain_distance(a, b):
    return po.edit_score(a, b, 999)  # 999 >= any distance of strings up to 999 bases: no early exit can fire below it


def test_golden_pairs_oracle():
    g = json.load(open(GOLDEN))
    assert len(g["pairs"]) > 300
    for a, b, T, want in g["pairs"]:
        assert po.edit_score(a, b, T) == want, (len(a), len(b), T)


def test_golden_pairs_cover_the_threshold_corner():
    """The recorded scores contain both answers the reference gives when the distance equals T."""
    g = json.load(open(GOLDEN))
    at_T = [w for a, b, T, w in g["pairs_at_threshold"] if w in (T, T + 1) and plain_distance(a, b) == T]
    assert any(w == T for (a, b, T, w) in g["pairs_at_threshold"] if plain_distance(a, b) == T)
    assert any(w == T + 1 for (a, b, T, w) in g["pairs_at_threshold"] if plain_distance(a, b) == T and len(a) and len(b))
    assert len(at_T) > 50
    for a, b, T, want in g["pairs_at_threshold"]:
        assert po.edit_score(a, b, T) == want


def test_golden_clusters_oracle():
    g = json.load(open(GOLDEN))
    n_fail = 0
    for c in g["clusters"]:
        ok, cent, n = oracle_cluster(c["seqs"], c["T"])
        assert ok == c["ok"]
        n_fail += (ok == 0)
        if ok:
            assert cent.tolist() == c["centroid_of"] and n == c["n_centroids"]
    assert n_fail >= 3


@pytest.mark.skipif(not po.ref_available(), reason="oracle/_ref not built")
def test_oracle_vs_reference_pairs():
    cases = ec.pair_cases(seed=23, n_random=150)
    cases += ec.at_threshold_cases(cases, plain_distance)
    for a, b, T in cases:
        assert po.edit_score(a, b, T) == po.edit_score(a, b, T, "ref"), (len(a), len(b), T)


@pytest.mark.skipif(not po.ref_available(), reason="oracle/_ref not built")
def test_oracle_vs_reference_clusters():
    for seqs, T in ec.cluster_cases(seed=9, n_sets=25):
        ok, cent, n = oracle_cluster(seqs, T)
        rok, rcent, rn = oracle_cluster(seqs, T, "ref")
        assert ok == rok
        if ok:
            assert cent.tolist() == rcent.tolist() and n == rn


def test_emulator_pairs_vs_oracle(emu):
    cases = ec.pair_cases(seed=11)
    cases += ec.at_threshold_cases(cases, plain_distance)
    n_flag = 0
    for a, b, T in cases:
        got, flag = emu_score(emu, a, b, T)
        n_flag += flag
        assert got == po.edit_score(a, b, T), (len(a), len(b), T)
    assert n_flag > 50  # the exact pass was exercised


def test_emulator_golden(emu):
    g = json.load(open(GOLDEN))
    for a, b, T, want in g["pairs"] + g["pairs_at_threshold"]:
        assert emu_score(emu, a, b, T)[0] == want
    for c in g["clusters"]:
        ok, cent, n = emu_cluster(emu, c["seqs"], c["T"])
        assert ok == c["ok"]
        if ok:
            assert cent.tolist() == c["centroid_of"] and n == c["n_centroids"]


def test_emulator_clusters_vs_oracle(emu):
    for seqs, T in ec.cluster_cases(seed=5):
        ok, cent, n = emu_cluster(emu, seqs, T)
        ook, ocent, on = oracle_cluster(seqs, T)
        assert ok == ook
        if ok:
            assert cent.tolist() == ocent.tolist() and n == on
