"""Records tests/golden/assembly.json: (1) seeded read clusters with the consensus HaplotypeGenerator::poa (reference, compiled
in place: oracle/_ref/libltr_ref_hapgen_poa.so) returns on top of the restated spoa (oracle/poa_restatement.hpp -- spoa itself
is un-vendored: parity of the consensus is UNPINNED, the fixture pins the product to the restatement); (2) candidate alleles
of synthetic regions that need the assembly, from the reference's own add_haplotype_block on the reads ltr_region_collect
prepared.  tests/test_assembly.py replays both without the reference.

    python tools/make_assembly_golden.py          (needs oracle/_ref; no GPU)"""
import json
import os
import random
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import bam_writer as bw  # noqa: E402
from longtr_b200 import abi  # noqa: E402
from oracle import pyregion as pr  # noqa: E402

WORLDS = [dict(config=3, n_loci=60, n_samples=1, first_locus=2000), dict(config=3, n_loci=40, n_samples=3, first_locus=2100),
          dict(config=4, n_loci=8, n_samples=2, first_locus=2200)]


def mutate(rng, s, sub, indel):
    out = []
    for ch in s:
        r = rng.random()
        if r < sub:
            out.append(rng.choice("ACGT"))
        elif r < sub + indel / 2:
            continue
        elif r < sub + indel:
            out.append(ch)
            out.append(rng.choice("ACGT"))
        else:
            out.append(ch)
    return "".join(out)


def clusters(seed=20261017, n=120):
    rng = random.Random(seed)
    for it in range(n):
        motif = "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 12)))
        L = rng.randint(3, 260)
        truth = (motif * (L // len(motif) + 1))[:L]
        k = rng.randint(1, 28)
        sub, indel = rng.choice([(0.001, 0.002), (0.01, 0.02), (0.05, 0.08), (0.15, 0.15)])
        seqs = [mutate(rng, truth, sub, indel) for _ in range(k)]
        if it % 7 == 3:
            seqs.append("")  # an empty sequence is skipped by the graph
        if it % 11 == 5:
            seqs = [s.replace("A", "N", 1) for s in seqs]
        yield seqs


def main():
    assert pr.ref_hapgen_poa_available()
    gold = dict(generator="tools/make_assembly_golden.py",
                source="reference HaplotypeGenerator compiled in place on oracle/poa_restatement.hpp (spoa restated, unpinned)",
                poa=[], worlds=[])
    for seqs in clusters():
        gold["poa"].append(dict(seqs=seqs, consensus=pr.ref_poa(seqs) if any(seqs) else ""))
    for W in WORLDS:
        world = bw.synthetic_world(W["n_loci"], config=W["config"], first_locus=W["first_locus"], n_samples=W["n_samples"])
        d = tempfile.mkdtemp()
        bams = [abi.BamFile(p) for p in bw.write_world(world, d)]
        for b in bams:
            b.build_index()
        regs = []
        for ri, (s, e, per) in enumerate(world["regions"]):
            got = abi.region_collect(bams, "chrS", s, e, world["chrom_seq"], 0, candidates=dict(period=per, flags=1))
            if not got["reads"] or got["candidates"]["status"] != 3:
                continue
            want = pr.ref_candidate_alleles(got["reads"], len(got["samples"]), s, e, world["chrom_seq"][s:s + per],
                                            world["chrom_seq"], 5, assemble=True)
            assert want["status"] == "ok"
            regs.append(dict(region=ri, alleles=want["alleles"], inexact=want["inexact"],
                             block=[want["block_start"], want["block_end"]]))
        gold["worlds"].append(dict(W, regions=regs))
        print("world", W, "regions with assembly:", len(regs), "inexact alleles:", sum(sum(r["inexact"]) for r in regs))
    path = os.path.join(ROOT, "tests", "golden", "assembly.json")
    json.dump(gold, open(path, "w"), separators=(",", ":"))
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
