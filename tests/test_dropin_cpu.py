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


@pytest.mark.skipif(not po.full_available("gpu"), reason="oracle/_ref/ltr_ref_gpu not built")
def test_gpu_binding_is_linked_and_refuses_to_run_without_a_gpu():
    """The drop-in build really routes HapAligner::process_reads through the C ABI: without a CUDA device it dies
    through LongTR's own printErrorAndDie instead of silently using the reference's CPU code."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        po.full_locus_records([dc.case_a4()], "gpu")


def test_reference_reproduces_real_data_records():
    """BASELINE.json configs[0] / [1] as parity cases: HG002 alone and the HG002+HG003+HG004 trio on the shipped BED
    regions (reads decoded by tools/real_cases.py), through the all-CPU reference genotyper."""
    cases = gu.load_real_cases()
    assert len(cases) >= 40
    recs = po.full_locus_records(cases, "full")
    for c, r in zip(cases, recs):
        assert r == c["record"], c["name"]


@pytest.mark.skipif(not po.full_available("lazy"), reason="oracle/_ref/ltr_ref_lazy not built")
def test_eliding_the_haplotype_alignment_changes_no_record():
    """N1: the reference's per-locus genotyper with Haplotype::aln_haps_to_ref replaced by
    integration/lazy_haplotype_alignment.cpp (no Needleman-Wunsch of the haplotypes against the reference allele,
    src/SeqAlignment/Haplotype.cpp:58-86) writes the very same VCF records -- seeded loci, the allele-pruning loci and the
    shipped HG002 / trio loci -- and so does the same binary with the original re-enabled through its env switch."""
    import os
    cases = [dc.case_a4()] + dc.seeded_cases() + dc.pruning_cases() + gu.load_real_cases()
    want = po.full_locus_records(cases, "full")
    assert po.full_locus_records(cases, "lazy") == want
    os.environ["LONGTR_B200_EAGER_HAP_ALIGNMENT"] = "1"
    try:
        assert po.full_locus_records(cases, "lazy") == want
    finally:
        del os.environ["LONGTR_B200_EAGER_HAP_ALIGNMENT"]
