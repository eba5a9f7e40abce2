"""GPU: ltr_pipeline_* -- loci submitted one at a time come back, by tag, with the values of the per-locus entry point."""
import numpy as np
import pytest

import synth
from longtr_b200 import Pipeline

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(120)
def test_pipeline_matches_per_locus_calls(engine):
    loci = []
    for seed in range(150):
        loc = synth.make_locus(12000 + seed, n_reads=8 + seed % 9, sub=0.01, indel=0.02, ref_len=40 + 7 * (seed % 40))
        L, keep = synth.to_flat(loc)
        loci.append((L, keep, len(loc["reads"]), len(loc["alleles"])))
    want = {}
    for tag, (L, keep, P, H) in enumerate(loci):
        want[tag] = engine.process_reads_flat(L, P, H, fill=3.5)
    pipe = Pipeline(0, batch_loci=32, slots=2)
    got = {}
    for tag, (L, keep, P, H) in enumerate(loci):
        pipe.submit(L, tag, fill=3.5)
        r = pipe.next(wait=False)          # the producer picks up whatever is finished as it goes
        if r is not None:
            got[r[0]] = r
    while True:                            # waiting also sends the last, partially filled batch on its way
        r = pipe.next(wait=True)
        if r is None:
            break
        got[r[0]] = r
    assert pipe.next(wait=False) is None
    pipe.close()
    assert sorted(got) == list(range(len(loci)))
    for tag in got:
        _, ll, seeds, status = got[tag]
        assert status == 0
        assert np.array_equal(ll, want[tag][0]) and np.array_equal(seeds, want[tag][1])


@pytest.mark.timeout(60)
def test_pipeline_reports_a_bad_locus_without_failing_its_batch(engine):
    good = []
    for seed in range(6):
        loc = synth.make_locus(13000 + seed, n_reads=6, ref_len=60)
        L, keep = synth.to_flat(loc)
        good.append((L, keep, len(loc["reads"]), len(loc["alleles"])))
    bad_loc = synth.make_locus(13100, n_reads=4, ref_len=60)
    Lb, keepb = synth.to_flat(bad_loc)
    Lb.period = 0                          # the flat API rejects the locus (LTR_ERR_INVALID)
    pipe = Pipeline(0, batch_loci=16, slots=1)
    for tag, (L, keep, P, H) in enumerate(good):
        pipe.submit(L, tag)
    pipe.submit(Lb, 99)
    res = {}
    while True:
        r = pipe.next(wait=True)
        if r is None:
            break
        res[r[0]] = r
    pipe.close()
    assert res[99][3] != 0
    for tag, (L, keep, P, H) in enumerate(good):
        ll, seeds = engine.process_reads_flat(L, P, H)
        assert res[tag][3] == 0 and np.array_equal(res[tag][1], ll)
