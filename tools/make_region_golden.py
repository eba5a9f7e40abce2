"""Records tests/golden/regions.json: what the REFERENCE's genotyper (SeqStutterGenotyper ctor -> genotype(), oracle/_ref/
ltr_ref_trace: its own HaplotypeGenerator, HapAligner and posteriors) makes of the reads that the library's region loop
(ltr_bam_* -> ltr_region_collect) prepares from synthetic BAM files (tests/bam_writer.py).  The reference's generator runs its
assembly branch on the restated spoa (oracle/poa_restatement.hpp: spoa is un-vendored, the consensus itself is unpinned); the
candidate alleles it arrives at must equal ltr_candidate_alleles' (asserted below) before anything is recorded.  The GPU test
tests/test_gpu_regions.py rebuilds the same BAM files from the same seeds and holds ltr_regions_run to these answers.

    python tools/make_region_golden.py          (needs /root/reference for oracle/_ref; no GPU)"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import bam_writer as bw  # noqa: E402
from longtr_b200 import abi  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

WORLDS = [dict(name="one_sample", n_loci=120, n_samples=1, first_locus=0),
          dict(name="two_samples", n_loci=80, n_samples=2, first_locus=500)]


def gapped(seq, cigar):
    import re
    out, si = [], 0
    for n, op in re.findall(r"(\d+)([=XID])", cigar):
        n = int(n)
        if op == "D":
            out.append("-" * n)
        else:
            out.append(seq[si:si + n])
            si += n
    return "".join(out)


def main():
    assert po.full_available("trace")
    gold = dict(generator="tools/make_region_golden.py", source="oracle/_ref/ltr_ref_trace on ltr_region_collect's reads", worlds=[])
    for W in WORLDS:
        world = bw.synthetic_world(W["n_loci"], config=3, first_locus=W["first_locus"], n_samples=W["n_samples"])
        d = tempfile.mkdtemp()
        bams = [abi.BamFile(p) for p in bw.write_world(world, d)]
        for b in bams:
            b.build_index()
        cases, meta = [], []
        for ri, (s, e, per) in enumerate(world["regions"]):
            got = abi.region_collect(bams, "chrS", s, e, world["chrom_seq"], 0, candidates=dict(period=per))
            c = got["candidates"]
            meta.append(dict(region=ri, status=c["status"]))
            if c["status"] != 0:
                continue
            S = len(got["samples"])
            motif = world["chrom_seq"][s:s + per]
            reads = [dict(start=r["start"], stop=r["stop"], rev=0, sample=r["sample"], name=r["name"], seq=r["seq"],
                          qual=r["qual"], aln=gapped(r["seq"], r["cigar"]), cigar=r["cigar"], log_p1=r["log_p1"],
                          log_p2=r["log_p2"]) for r in got["reads"]]
            cases.append(dict(name="%s_%d" % (W["name"], ri), chrom_name="chrS", chrom_seq=world["chrom_seq"], region_start=s,
                              region_stop=e, motif=motif, region_name="R%d" % ri, samples=["S%d" % f for f in got["samples"]],
                              n_p1s=[sum(1 for r in got["reads"] if r["sample"] == k and r["hp"] == 1) for k in range(S)],
                              n_p2s=[sum(1 for r in got["reads"] if r["sample"] == k and r["hp"] == 2) for k in range(S)], reads=reads, stutter_motif="A", stutter_period=per))
            meta[-1]["case"] = len(cases) - 1
            meta[-1]["alleles"] = c["alleles"]
            meta[-1]["inexact"] = c["inexact"]
            meta[-1]["samples"] = got["samples"]
        traces = po.full_locus_traces(cases)
        regions = []
        for m in meta:
            if "case" not in m:
                regions.append(dict(region=m["region"], status=m["status"]))
                continue
            rec, calls = traces[m["case"]]
            if not calls:
                regions.append(dict(region=m["region"], status=m["status"], reference_failed=True))
                continue
            first, last = calls[0], calls[-1]
            assert first["alleles"] == m["alleles"], (m["region"], first["alleles"], m["alleles"])
            regions.append(dict(region=m["region"], status=0, alleles=m["alleles"], inexact=m["inexact"], samples=m["samples"],
                                S=first["S"],
                                block=[first["repeat_start"], first["repeat_end"]], lflank=first["lflank"], rflank=first["rflank"],
                                kept=[first["alleles"].index(a) for a in last["alleles"]], out_gts=last["gts"],
                                record=rec, motif=world["chrom_seq"][world["regions"][m["region"]][0]:world["regions"][m["region"]][0] + world["regions"][m["region"]][2]],
                                out_post=last["post"], out_totals=last["totals"]))
        gold["worlds"].append(dict(W, regions=regions))
        print(W["name"], "regions", len(regions), "genotyped by the reference", sum(1 for r in regions if "kept" in r))
    path = os.path.join(ROOT, "tests", "golden", "regions.json")
    json.dump(gold, open(path, "w"), separators=(",", ":"))
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
