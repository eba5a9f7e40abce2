"""GPU parity of kernel 2 (homopolymer / --stutter-align-len path) through ltr_process_reads_flat against the values
recorded from the unmodified reference (tests/golden: Appendix A3 + 24 seeded loci) and the oracle on fresh loci.
Same doubles, same operation order, bit-faithful fasterexp/fasterlog -> bit equality expected and required."""
import numpy as np
import pytest

import golden_util as gu
import synth
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

SHORT_CASES = [c for c in gu.load("appendix_a") if c["switch"] != 0] + gu.load("process_reads_short")


@pytest.mark.parametrize("case", SHORT_CASES, ids=lambda c: c["name"])
def test_short_path_matches_reference(engine, case):
    L, keep = gu.flat_locus(case)
    P, H = len(case["reads"]), len(case["alleles"])
    ll, seeds = engine.process_reads_flat(L, P, H, fill=case.get("fill", 0.0))
    want = gu.unhex(case["ll"], (P, H))
    assert np.array_equal(ll, want), (ll - want)
    for r in range(P):
        if case.get("realign_read") is None or case["realign_read"][r]:
            assert seeds[r] == case["seeds"][r]


@pytest.mark.parametrize("seed", range(12))
def test_short_path_matches_oracle_on_fresh_loci(engine, seed):
    loc = synth.make_locus(8100 + seed, n_reads=8, homopolymer=True, ref_len=12 + 7 * seed, flank=35 + 12 * (seed % 4),
                           ctx=30 + 20 * (seed % 3), sub=0.004 * (seed % 3), indel=0.01 * (seed % 4))
    L, keep = synth.to_flat(loc, switch_old_align_len=20)
    P, H = len(loc["reads"]), len(loc["alleles"])
    want, wseeds, _ = po.process_reads(L, P, H)
    got, gseeds = engine.process_reads_flat(L, P, H)
    assert np.array_equal(got, want)
    assert np.array_equal(gseeds, wseeds)


def test_stutter_batch_config5_matches_oracle(engine):
    """A batch of BASELINE config-5 loci through ltr_stutter_ll (one launch) against the oracle, locus by locus."""
    from longtr_b200 import workloads
    work = workloads.generate_stutter(48)
    got, st = engine.stutter_ll(work.batch)
    assert st.n_pairs > 0 and st.n_launches == 1
    off = 0
    for l in range(work.n_loci):
        (L, keep), (P, H) = work.flat_locus(l)
        want, _seeds, _ = po.process_reads(L, P, H)
        assert np.array_equal(got[off:off + P * H].reshape(P, H), want), l
        off += P * H
    assert off == len(got)
