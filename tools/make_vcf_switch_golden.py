"""Golden VCF records under the reference's output switches (--hide-allreads, --hide-mallreads, --output-gls, --output-pls,
--output-phased-gls, --output-filters: src/hipstr_main.cpp:178-183 -> Genotyper::OUTPUT_*): the 52 shipped HG002 / trio loci of
tests/golden/real_cases.json.gz and the seeded drop-in loci (tests/dropin_cases.py, haploid ones among them) through the
UNMODIFIED reference (oracle/_ref/ltr_ref_full, LTR_REF_OUTPUT_SWITCHES) once per switch combination ->
tests/golden/vcf_switches.json.gz = {mask: {case name: record}} with the masks in the library's LTR_VCF_* bits; "haploid": the
first five seeded loci as a haploid chromosome, the default switches (3) included.
Run where /root/reference is mounted:  python tools/make_vcf_switch_golden.py"""
import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

MASKS = (0, 7, 60, 63)   # nothing optional / ALLREADS+MALLREADS+GL / GL+PL+PHASEDGL+FILTER / everything


def main():
    import dropin_cases as dc
    import golden_util as gu
    from oracle import pyoracle as po
    cases = gu.load_real_cases() + [dc.case_a4()] + dc.seeded_cases()
    out = {}
    for m in MASKS:
        recs = po.full_locus_records(cases, "full", switches=m)
        out[str(m)] = {c["name"]: r for c, r in zip(cases, recs) if r}
        print("mask", m, len(out[str(m)]), "records")
    hap = {}
    for m in (3,) + MASKS:
        recs = po.full_locus_records(dc.haploid_cases(), "full", switches=m)
        hap[str(m)] = {c["name"]: r for c, r in zip(dc.haploid_cases(), recs) if r}
        print("haploid, mask", m, len(hap[str(m)]), "records")
    # the default switches must give the records already pinned
    recs = po.full_locus_records(cases[:52], "full", switches=3)
    assert all(r == c["record"] for c, r in zip(cases[:52], recs) if "record" in c)
    path = os.path.join(ROOT, "tests", "golden", "vcf_switches.json.gz")
    with gzip.open(path, "wt") as f:
        json.dump(dict(generator="tools/make_vcf_switch_golden.py", source="oracle/_ref/ltr_ref_full with LTR_REF_OUTPUT_SWITCHES",
                       masks=out, haploid=hap), f, separators=(",", ":"))
    print(path, os.path.getsize(path), "bytes")
    k = next(iter(out["63"]))
    print(out["63"][k].split("\t")[8:])
    k = next(iter(hap["63"]))
    print(hap["63"][k].split("\t")[8:])


if __name__ == "__main__":
    main()
