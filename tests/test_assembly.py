"""N2, assembly branch of the candidate alleles (csrc/host/poa.cpp, csrc/host/candidate_alleles.cpp; reference
src/SeqAlignment/HaplotypeGenerator.cpp:167-292, 397-471).  Host only.

The consensus comes from spoa in the reference; spoa is un-vendored and unpinned, so its parity is UNPINNED: the product's
implementation is held to an independent restatement of the published algorithm (oracle/poa_restatement.hpp) and to what
that algorithm must deliver (a cluster of identical reads, majority reads).  Everything around the consensus -- clustering
ladder, consensus / merge rounds, support tests, ordering, trimming -- is held to the reference's own HaplotypeGenerator,
compiled in place on top of that restatement: recorded in tests/golden/assembly.json (tools/make_assembly_golden.py) and, where
oracle/_ref is present, live on fresh worlds."""
import json
import os
import random

import pytest

import bam_writer as bw
from longtr_b200 import abi
from oracle import pyregion as pr

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "assembly.json")


@pytest.fixture(scope="module")
def gold():
    return json.load(open(GOLD))


def test_consensus_matches_the_recorded_restatement(gold):
    assert len(gold["poa"]) >= 100
    for k, case in enumerate(gold["poa"]):
        assert abi.poa_consensus(case["seqs"]) == case["consensus"], k


def test_consensus_properties():
    rng = random.Random(5)
    assert abi.poa_consensus([]) == "" and abi.poa_consensus(["", ""]) == ""
    assert abi.poa_consensus(["ACGTTGCA"]) == "ACGTTGCA"
    for _ in range(40):
        truth = "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 400)))
        assert abi.poa_consensus([truth] * rng.randint(1, 12)) == truth
        # a minority of reads with a private substitution / insertion / deletion does not change the consensus
        reads = [truth] * 7
        for _ in range(3):
            p = rng.randrange(len(truth))
            reads.append(truth[:p] + rng.choice(["", "A", "CG"]) + truth[p + rng.randint(0, 1):])
        rng.shuffle(reads)
        assert abi.poa_consensus(reads) == truth


def test_consensus_buffer_too_small():
    import ctypes as C

    import numpy as np
    lib = abi.load()
    data, off = abi.pack_seqs(["ACGTACGT", "ACGTACGT"])
    out = np.zeros(4, dtype=np.uint8)
    n = C.c_uint32(0)
    u8p, u32p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32)
    rc = lib.ltr_poa_consensus(data.ctypes.data_as(u8p), off.ctypes.data_as(u32p), 2, out.ctypes.data_as(u8p), 4, C.byref(n))
    assert rc == -3 and n.value == 8


def _world(W, tmp):
    world = bw.synthetic_world(W["n_loci"], config=W["config"], first_locus=W["first_locus"], n_samples=W["n_samples"])
    bams = [abi.BamFile(p) for p in bw.write_world(world, str(tmp))]
    for b in bams:
        b.build_index()
    return world, bams


@pytest.mark.parametrize("wi", [0, 1, 2])
def test_assembled_alleles_match_the_recorded_reference(gold, tmp_path, wi):
    W = gold["worlds"][wi]
    world, bams = _world(W, tmp_path)
    n_inexact = 0
    for g in W["regions"]:
        s, e, per = world["regions"][g["region"]]
        stop = abi.region_collect(bams, "chrS", s, e, world["chrom_seq"], 0, candidates=dict(period=per, flags=1))["candidates"]
        assert stop["status"] == 3 and stop["cluster_sets"] and stop["assembly_threshold"] == 0
        c = abi.region_collect(bams, "chrS", s, e, world["chrom_seq"], 0, candidates=dict(period=per))["candidates"]
        assert c["status"] == 0 and c["assembly_threshold"] >= 20 and c["n_consensus"] >= 1
        assert c["alleles"] == g["alleles"] and c["inexact"] == g["inexact"], g["region"]
        assert [c["block_start"], c["block_end"]] == g["block"]
        assert c["cluster_sets"] == stop["cluster_sets"]
        n_inexact += sum(c["inexact"])
    assert len(W["regions"]) >= 8 and n_inexact >= 10


@pytest.mark.skipif(not pr.ref_hapgen_poa_available(), reason="oracle/_ref/libltr_ref_hapgen_poa.so not built")
@pytest.mark.parametrize("config,n_loci,n_samples,first", [(3, 60, 1, 7000), (3, 40, 2, 7100), (4, 6, 3, 7200)])
def test_assembled_alleles_match_the_reference_on_fresh_worlds(tmp_path, config, n_loci, n_samples, first):
    world, bams = _world(dict(config=config, n_loci=n_loci, n_samples=n_samples, first_locus=first), tmp_path)
    n_asm = 0
    for s, e, per in world["regions"]:
        got = abi.region_collect(bams, "chrS", s, e, world["chrom_seq"], 0, candidates=dict(period=per))
        if not got["reads"]:
            continue
        c = got["candidates"]
        want = pr.ref_candidate_alleles(got["reads"], len(got["samples"]), s, e, world["chrom_seq"][s:s + per],
                                        world["chrom_seq"], 5, assemble=True)
        if want["status"] != "ok":
            assert c["status"] in (1, 2)
            continue
        assert c["status"] == 0
        assert c["alleles"] == want["alleles"] and c["inexact"] == want["inexact"]
        assert (c["block_start"], c["block_end"], c["lflank_start"]) == (want["block_start"], want["block_end"], want["lflank_start"])
        n_asm += c["assembly_threshold"] > 0
    assert n_asm >= 5


@pytest.mark.skipif(not pr.ref_hapgen_poa_available(), reason="oracle/_ref/libltr_ref_hapgen_poa.so not built")
def test_consensus_matches_the_restatement_on_fresh_clusters():
    rng = random.Random(99)
    for _ in range(150):
        motif = "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 30)))
        L = rng.randint(1, 500)
        truth = (motif * (L // len(motif) + 1))[:L]
        sub, indel = rng.choice([(0.002, 0.002), (0.01, 0.03), (0.1, 0.1), (0.3, 0.3)])
        seqs = []
        for _ in range(rng.randint(1, 29)):
            out = []
            for ch in truth:
                r = rng.random()
                if r < sub:
                    out.append(rng.choice("ACGT"))
                elif r < sub + indel / 2:
                    continue
                elif r < sub + indel:
                    out.append(ch + rng.choice("ACGT"))
                else:
                    out.append(ch)
            seqs.append("".join(out))
        if not any(seqs):
            continue
        assert abi.poa_consensus(seqs) == pr.ref_poa(seqs)


def test_host_edit_distances_fuzz(tmp_path):
    """ltr::edit_distance / ltr::bounded_edit_distance (the clustering's distances, host) against the plain recurrence."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "host_edit_fuzz")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + os.path.join(root, "longtr_b200", "csrc", "host"),
                           os.path.join(root, "tests", "emu", "host_edit_fuzz.cpp"),
                           os.path.join(root, "longtr_b200", "csrc", "host", "poa.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and " bad 0" in out.stdout, out.stdout[-400:]


def test_consensus_with_wide_cells():
    """Sequences long enough that the alignment matrix needs 32-bit cells (16-bit cells serve N + 3 len < 32 000)."""
    if not pr.ref_hapgen_poa_available():
        pytest.skip("oracle/_ref/libltr_ref_hapgen_poa.so not built")
    rng = random.Random(3)
    truth = "".join(rng.choice("ACGT") for _ in range(9000))
    seqs = []
    for _ in range(3):
        out = []
        for ch in truth:
            r = rng.random()
            if r < 0.002:
                out.append(rng.choice("ACGT"))
            elif r < 0.003:
                continue
            elif r < 0.004:
                out.append(ch + rng.choice("ACGT"))
            else:
                out.append(ch)
        seqs.append("".join(out))
    assert abi.poa_consensus(seqs) == pr.ref_poa(seqs)
