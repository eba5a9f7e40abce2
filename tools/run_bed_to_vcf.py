#!/usr/bin/env python
"""The whole flow through the C ABI, as a few calls (an example, not a command-line tool: LongTR's CLI is out of scope):

    python tools/run_bed_to_vcf.py --bams a.bam,b.bam --fasta ref.fa --regions regions.bed --out calls.vcf [--samples A,B]

BAM files (one per sample) + indexed FASTA + region file (CHROM START STOP MOTIF [NAME]) -> VCF text: ltr_bam_open,
ltr_fasta_open, ltr_run_bed with vcf_records (host threads: read filters, trimming, candidate alleles; GPU: alignment,
posteriors, removal of uncalled alleles), ltr_vcf_header + the records in region order."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--bams", required=True)
    ap.add_argument("--fasta", required=True)
    ap.add_argument("--regions", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--samples", default="")
    ap.add_argument("--devices", default="0")
    ap.add_argument("--no-phased-bam", action="store_true", help="ignore HP tags (the library's default is --phased-bam)")
    # the reference's output switches (src/hipstr_main.cpp:178-183) -> LTR_VCF_* (ltr_regions_opts.vcf_switches, ltr_vcf_header_ex)
    for flag in ("hide-allreads", "hide-mallreads", "output-gls", "output-pls", "output-phased-gls", "output-filters"):
        ap.add_argument("--" + flag, action="store_true")
    a = ap.parse_args(argv)
    from longtr_b200 import Genotyper, abi
    paths = a.bams.split(",")
    samples = a.samples.split(",") if a.samples else [os.path.basename(p).split(".")[0] for p in paths]
    bams = [abi.BamFile(p) for p in paths]
    for b in bams:
        if not b.has_index:
            b.build_index()
    fasta = abi.FastaFile(a.fasta)
    switches = ((0 if a.hide_allreads else abi.VCF_ALLREADS) | (0 if a.hide_mallreads else abi.VCF_MALLREADS) |
                (abi.VCF_GLS if a.output_gls else 0) | (abi.VCF_PLS if a.output_pls else 0) |
                (abi.VCF_PHASED_GLS if a.output_phased_gls else 0) | (abi.VCF_FILTERS if a.output_filters else 0))
    g = Genotyper(devices=tuple(int(d) for d in a.devices.split(",")))
    try:
        run = g.run_bed(bams, fasta, a.regions, vcf_records=True, vcf_switches=switches,
                        **(dict(phased_bam=0) if a.no_phased_bam else {}))
    finally:
        g.close()
    n = 0
    with open(a.out, "w") as f:
        f.write(abi.vcf_header(fasta, a.fasta, " ".join(["run_bed_to_vcf.py"] + (argv if argv is not None else sys.argv[1:])),
                               samples, switches=switches))
        for res in run["per_chrom"]:
            for rec in res["records"] or []:
                if rec:
                    f.write(rec + "\n")
                    n += 1
    print("%d records -> %s" % (n, a.out))
    return n


if __name__ == "__main__":
    main()
