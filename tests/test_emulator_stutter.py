"""CPU check of the warp-level logic of kernel 2 (homopolymer / stutter path): the kernel's per-lane functions
(longtr_b200/csrc/stutter_core.cuh) driven by a host-side lane emulator, against the golden vectors recorded from
the reference and against the oracle -- bit for bit, no GPU needed."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import golden_util as gu
import synth
from longtr_b200 import abi
from longtr_b200.flat import DEFAULT_STUTTER
from oracle import pyoracle as po

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call([os.path.join(HERE, "emu", "build_emu.sh")])
    lib = C.CDLL(os.path.join(HERE, "emu", "libltr_emu_stutter.so"))
    lib.ltr_emu_stutter_pair.argtypes = [C.c_char_p] * 5 + [C.c_int32, C.POINTER(C.c_double), C.c_int32,
                                                           C.POINTER(C.c_float), C.POINTER(C.c_double)]
    lib.ltr_emu_stutter_pair.restype = C.c_int
    return lib


def emu_locus(emu, case):
    P, H = len(case["reads"]), len(case["alleles"])
    out = np.full((P, H), case.get("fill", 0.0))
    params = case["aln_params"] or abi.DEFAULT_ALN_PARAMS
    par = (C.c_float * 7)(*params)
    st = (C.c_double * 6)(*DEFAULT_STUTTER)
    for r, read in enumerate(case["reads"]):
        if case.get("realign_read") is not None and not case["realign_read"][r]:
            continue
        seed = case["seeds"][r]
        if seed < 0:
            out[r, :] = 0.0
            continue
        for a, allele in enumerate(case["alleles"]):
            if case.get("realign_to_hap") is not None and not case["realign_to_hap"][a]:
                continue
            v = C.c_double(0)
            rc = emu.ltr_emu_stutter_pair(case["lflank"].encode(), case["rflank"].encode(), allele.encode(),
                                          read["seq"].encode(), read["qual"].encode(), seed, st, len(case["motif"]),
                                          par, C.byref(v))
            assert rc == 0
            out[r, a] = v.value
    return out


SHORT_CASES = [c for c in gu.load("appendix_a") if c["switch"] != 0] + gu.load("process_reads_short")


@pytest.mark.parametrize("case", SHORT_CASES, ids=lambda c: c["name"])
def test_emulator_matches_reference_golden(emu, case):
    P, H = len(case["reads"]), len(case["alleles"])
    got = emu_locus(emu, case)
    want = gu.unhex(case["ll"], (P, H))
    assert np.array_equal(got, want), (got, want)


@pytest.mark.parametrize("seed", range(10))
def test_emulator_matches_oracle_long_flanks_and_blocks(emu, seed):
    """Flanks longer than one 64-row strip and longer homopolymers (multi-strip hand-off)."""
    loc = synth.make_locus(8000 + seed, n_reads=3, homopolymer=True, ref_len=20 + 9 * seed, flank=35 + 15 * seed,
                           ctx=30, sub=0.01, indel=0.02)
    L, keep = synth.to_flat(loc, switch_old_align_len=20)
    P, H = len(loc["reads"]), len(loc["alleles"])
    want, seeds, _ = po.process_reads(L, P, H)
    case = dict(lflank=loc["lflank"], rflank=loc["rflank"], alleles=loc["alleles"], reads=loc["reads"],
                motif=loc["motif"], aln_params=None, seeds=[int(s) for s in seeds])
    got = emu_locus(emu, case)
    assert np.array_equal(got, want)
