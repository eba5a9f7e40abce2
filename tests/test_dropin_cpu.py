"""CPU: the all-reference full-locus driver (oracle/_ref/ltr_ref_full: SeqStutterGenotyper ctor -> genotype ->
write_vcf_record, IO-less) reproduces SURVEY Appendix A4 and the committed golden VCF records.  This pins the
fixtures that tests/test_gpu_dropin.py holds the GPU drop-in build to."""
import pytest

import dropin_cases as dc
import golden_util as gu
from oracle import pyoracle as po

pytestmark = pytest.mark.skipif(not po.full_available("full"), reason="oracle/_ref/ltr_ref_full not built")


def test_reference_reproduces_appendix_a4_and_golden_records():
    cases = [dc.case_a4()] + dc.seeded_cases()
    recs = po.full_locus_records(cases, "full")
    assert recs[0] == dc.A4_RECORD
    gold = {g["name"]: g["record"] for g in gu.load("vcf_records")}
    for c, r in zip(cases, recs):
        assert r == gold[c["name"]], c["name"]
        assert r, "genotype() failed for " + c["name"]
