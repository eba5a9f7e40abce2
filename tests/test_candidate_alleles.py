"""N2: ltr_candidate_alleles (csrc/host/candidate_alleles.cpp) against the reference's own HaplotypeGenerator
(add_haplotype_block + fuse_haplotype_blocks compiled in place, oracle/hapgen_driver.cpp) on the reads ltr_region_collect
prepares from the shipped trio BAMs: same alleles in the same order, same block, same flanks, same verdicts ("no spanning
alignments", "needs the assembly").  Host only; runs where /root/reference is mounted."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from longtr_b200 import abi  # noqa: E402
from oracle import pyregion as pr  # noqa: E402

DATA = os.path.join(os.environ.get("LONGTR_REFERENCE", "/root/reference"), "test_data")
SAMPLES = ["HG002", "HG003", "HG004"]
pytestmark = [pytest.mark.skipif(not os.path.exists(os.path.join(DATA, "HG002_sample_reads.bam")),
                                 reason="reference test data not mounted"),
              pytest.mark.skipif(not pr.ref_hapgen_available(), reason="oracle/_ref/libltr_ref_hapgen.so not built")]
STATUS = {1: "Haplotype blocks are too near to the chromosome ends", 2: "No spanning alignments", 3: "needs assembly"}


@pytest.fixture(scope="module")
def world():
    import real_cases
    bams = [abi.BamFile(os.path.join(DATA, s + "_sample_reads.bam")) for s in SAMPLES]
    tid = [i for i, (n, _) in enumerate(bams[0].refs) if n == "chr1"][0]
    allr = [r for b in bams for r in b.fetch(tid, 0, 1 << 29)]
    lo, hi = min(r["pos"] for r in allr), max(r["end"] for r in allr)
    ref = real_cases.build_pseudo_reference([dict(r, seq=r["seq"].upper()) for r in allr], lo, hi)
    return dict(bams=bams, ref=ref, ref_start=lo, regions=[r for r in real_cases.regions() if r["chrom"] == "chr1"])


@pytest.mark.parametrize("which,flank,assemble", [("trio", 5, False), ("single", 5, False), ("trio", 12, False),
                                                  ("trio", 5, True), ("single", 5, True)])
def test_candidates_match_the_reference(world, which, flank, assemble):
    """assemble=False: the reference build whose spoa stand-in throws, against LTR_CAND_FLAG_NO_ASSEMBLY (same verdict "needs
    assembly"); assemble=True: the build on the restated spoa against the default call (same alleles, inexact flags)."""
    if assemble and not pr.ref_hapgen_poa_available():
        pytest.skip("oracle/_ref/libltr_ref_hapgen_poa.so not built")
    bams = world["bams"] if which == "trio" else world["bams"][:1]
    # the driver sees a chromosome that starts at position 0: pad the slice in front
    chrom = "N" * world["ref_start"] + world["ref"]
    seen = {"ok": 0, "needs assembly": 0, "other": 0}
    n_multi = 0
    for reg in world["regions"]:
        motif = reg["motif"].split(",")[0]
        got = abi.region_collect(bams, "chr1", reg["start"], reg["stop"], world["ref"], world["ref_start"],
                                 candidates=dict(period=len(motif), indel_flank_len=flank, flags=0 if assemble else 1))
        if not got["reads"]:
            continue
        c = got["candidates"]
        want = pr.ref_candidate_alleles(got["reads"], len(got["samples"]), reg["start"], reg["stop"], motif, chrom, flank,
                                        assemble=assemble)
        if want["status"] != "ok":
            assert STATUS.get(c["status"]) == want["status"], (reg["name"], c["status"], want["status"])
            seen["needs assembly" if want["status"] == "needs assembly" else "other"] += 1
            continue
        assert c["status"] == 0, (reg["name"], c["status"])
        assert c["alleles"] == want["alleles"] and c["inexact"] == want["inexact"], reg["name"]
        seen["assembled"] = seen.get("assembled", 0) + (c["assembly_threshold"] > 0)
        assert (c["block_start"], c["block_end"], c["lflank_start"]) == \
               (want["block_start"], want["block_end"], want["lflank_start"])
        assert c["lflank"] == want["lflank"] and c["rflank"] == want["rflank"]
        seen["ok"] += 1
        n_multi += len(c["alleles"]) > 1
    assert seen["ok"] >= 15 and n_multi >= 8, seen
    if assemble:
        assert seen["needs assembly"] == 0 and seen["assembled"] >= (3 if which == "trio" else 0), seen
