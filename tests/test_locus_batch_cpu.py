"""CPU: host-only pieces of the batch genotyper (ltr_genotyper_*, csrc/host/locus_batcher.cpp): the run-wise
HapAligner::trim_alignment on BAM-encoded CIGARs against the host mirror (base-by-base, like the reference) and the
oracle's restatement, on seeded loci with indel-rich reads and on the reference-recorded fixtures."""
import ctypes as C

import numpy as np
import pytest

import golden_util as gu
import synth
from longtr_b200 import abi, locus_batch as lb
from oracle import pyoracle as po


def _as_batch_locus(loc, rng=None):
    reads = [dict(start=r["start"], stop=r["stop"], seq=r["seq"], cigar=r["cigar"], sample=0, log_p1=-0.7, log_p2=-0.7)
             for r in loc["reads"]]
    return dict(lflank=loc["lflank"], rflank=loc["rflank"], alleles=loc["alleles"], repeat_start=loc["repeat_start"],
                repeat_end=loc["repeat_end"], n_samples=1, reads=reads)


@pytest.mark.parametrize("seed", range(40))
def test_trim_matches_mirror_and_oracle(seed):
    rng = np.random.default_rng(seed)
    loc = synth.make_locus(3000 + seed, n_reads=10, sub=float(rng.choice([1e-3, 0.02])), indel=float(rng.choice([1e-3, 0.05, 0.15])),
                           ctx=int(rng.integers(0, 80)))
    flat, keep = synth.to_flat(loc)
    s, keep2 = lb.make_locus_batch(lb.build_locus_batch([_as_batch_locus(loc)]))
    for r in range(len(loc["reads"])):
        want = abi.trim_read(flat, r)
        assert want == po.trim_read(flat, r)
        if len(want) == 0:  # the caller of trim_alignment substitutes the 10 bp pseudo read (HapAligner.cpp:820-823)
            want = (loc["lflank"][-5:] + loc["rflank"][:5]).encode()
        assert lb.trim_read(s, 0, r) == want


def test_trim_of_reference_recorded_loci():
    """The loci of tests/golden/pruning.json (incl. reads whose CIGAR is shorter than the read, see dropin_cases.make_case)."""
    from longtr_b200.flat import make_flat_locus
    n = 0
    for c in gu.load("pruning"):
        if c["reads"] is None:
            continue
        s, keep = lb.make_locus_batch(lb.build_locus_batch([dict(lflank=c["lflank"], rflank=c["rflank"], alleles=c["alleles"],
                                                                 repeat_start=c["repeat_start"], repeat_end=c["repeat_end"],
                                                                 n_samples=c["S"], reads=c["reads"])]))
        reads = [(r["start"], r["stop"], r["seq"], "I" * len(r["seq"]), r["cigar"]) for r in c["reads"]]
        flat, keep_f = make_flat_locus(c["lflank"], c["alleles"], c["rflank"], c["repeat_start"], c["repeat_end"], 2, reads, motif="AC")
        for r in range(len(reads)):
            want = abi.trim_read(flat, r)
            if len(want) == 0:
                want = (c["lflank"][-5:] + c["rflank"][:5]).encode()
            assert lb.trim_read(s, 0, r) == want
            n += 1
    assert n > 300


def test_trim_rejects_what_the_reference_dies_on():
    loc = synth.make_locus(7, n_reads=2)
    b = _as_batch_locus(loc)
    b["reads"][0]["cigar"] = "10N" + b["reads"][0]["cigar"]
    s, keep = lb.make_locus_batch(lb.build_locus_batch([b]))
    with pytest.raises(RuntimeError):
        lb.trim_read(s, 0, 0)
    assert len(lb.trim_read(s, 0, 1)) > 0
