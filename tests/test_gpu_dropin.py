"""GPU drop-in check: LongTR's own SeqStutterGenotyper / HaplotypeGenerator / VCF writer (reference objects,
compiled in place) linked with integration/reference_binding.cpp, which replaces HapAligner::process_reads and
Genotyper::calc_log_sample_posteriors by calls into liblongtr_b200.so.  The VCF records written that way must be
IDENTICAL, character for character, to the all-CPU reference's (tests/golden/vcf_records.json incl. SURVEY A4):
same GT, allele sequences, GB, Q, PQ, DP, GLDIFF, ALLREADS."""
import pytest

import dropin_cases as dc
import golden_util as gu
from oracle import pyoracle as po

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not po.full_available("gpu"), reason="oracle/_ref/ltr_ref_gpu not built")]


def test_vcf_records_identical_to_reference():
    cases = [dc.case_a4()] + dc.seeded_cases()
    recs = po.full_locus_records(cases, "gpu")
    gold = {g["name"]: g["record"] for g in gu.load("vcf_records")}
    assert recs[0] == dc.A4_RECORD
    for c, r in zip(cases, recs):
        assert r == gold[c["name"]], (c["name"], r, gold[c["name"]])


def test_real_data_vcf_records_identical_to_reference():
    """The shipped trio reads (BASELINE.json configs[0]: HG002; configs[1]: HG002+HG003+HG004 joint genotyping, the
    multi-sample posterior path) on the shipped BED regions: LongTR's genotyper on top of the GPU library writes
    the same VCF records, character for character, as the all-CPU reference (fixtures: tools/real_cases.py)."""
    cases = gu.load_real_cases()
    recs = po.full_locus_records(cases, "gpu")
    bad = [c["name"] for c, r in zip(cases, recs) if r != c["record"]]
    assert not bad, bad
