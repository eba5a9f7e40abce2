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


def test_bad_locus_fails_alone_on_the_short_path(engine):
    """ltr_stutter_ll_status: an empty (<DEL>) allele, a seed outside the read and a broken stutter model each fail their own
    locus only; the rows of the other loci equal an undisturbed run bit for bit, the rows of the failed loci stay untouched.
    ltr_stutter_ll (all or nothing) fails as a whole on the same batch and leaves the output array as it was."""
    from longtr_b200 import workloads
    from longtr_b200.engine import LongTRError
    work = workloads.generate_stutter(24)
    good, _st = engine.stutter_ll(work.batch)
    b = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in work.batch.items()}
    lab, lrb = b["locus_allele_begin"], b["locus_read_begin"]
    # locus 3: its second allele becomes empty (shift nothing: just make two offsets equal by emptying that allele)
    a = int(lab[3]) + 1
    cut = int(b["allele_off"][a + 1] - b["allele_off"][a])
    b["allele_bytes"] = np.concatenate([b["allele_bytes"][:b["allele_off"][a]], b["allele_bytes"][b["allele_off"][a + 1]:]])
    b["allele_off"] = b["allele_off"].copy()
    b["allele_off"][a + 1:] -= cut
    # locus 7: a seed beyond the read; locus 11: a stutter model that is not a probability distribution
    b["read_seed"] = b["read_seed"].copy()
    b["read_seed"][lrb[7]] = 10 ** 6
    b["stutter"] = b["stutter"].copy()
    b["stutter"][6 * 11 + 1] = 0.99
    sentinel = 123.25
    out = np.full(len(good), sentinel)
    got, st, status = engine.stutter_ll(b, out=out, per_locus_status=True)
    assert status[3] == -5 and status[7] == -3 and status[11] == -3
    assert (np.delete(status, [3, 7, 11]) == 0).all()
    H = np.diff(lab).astype(np.int64)
    P = np.diff(lrb).astype(np.int64)
    off = np.concatenate([[0], np.cumsum(H * P)])
    for l in range(work.n_loci):
        rows = got[off[l]:off[l + 1]]
        if l in (3, 7, 11):
            assert (rows == sentinel).all()
        else:
            assert np.array_equal(rows, good[off[l]:off[l + 1]]), l
    out2 = np.full(len(good), sentinel)
    with pytest.raises(LongTRError):
        engine.stutter_ll(b, out=out2)
    assert (out2 == sentinel).all()
