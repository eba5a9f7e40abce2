"""Loaders for tests/golden/*.json (written by tools/make_golden.py from oracle/_ref)."""
import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(GOLD, name + ".json")) as f:
        return json.load(f)["cases"]


def unhex(xs, shape=None):
    a = np.array([float.fromhex(x) for x in xs], dtype=np.float64)
    return a if shape is None else a.reshape(shape)


def flat_locus(case):
    from longtr_b200.flat import make_flat_locus
    reads = [(r["start"], r["stop"], r["seq"], r["qual"], r["cigar"]) for r in case["reads"]]
    return make_flat_locus(case["lflank"], case["alleles"], case["rflank"], case["repeat_start"],
                           case["repeat_end"], case["period"], reads, motif=case["motif"],
                           switch_old_align_len=case["switch"], aln_params=case["aln_params"],
                           realign_to_hap=case.get("realign_to_hap"), realign_read=case.get("realign_read"))


def pair_batch(case):
    haps, reads = case["haps"], case["reads"]
    hoff = np.concatenate([[0], np.cumsum([len(s) for s in haps])]).astype(np.uint32)
    roff = np.concatenate([[0], np.cumsum([len(s) for s in reads])]).astype(np.uint32)
    return dict(locus_hap_begin=np.array(case["locus_hap_begin"], dtype=np.uint32),
                locus_read_begin=np.array(case["locus_read_begin"], dtype=np.uint32),
                hap_off=hoff, read_off=roff,
                hap_bytes=np.frombuffer("".join(haps).encode(), dtype=np.uint8).copy(),
                read_bytes=np.frombuffer("".join(reads).encode(), dtype=np.uint8).copy())


def load_real_cases():
    """Per-locus inputs cut from the reference's shipped HG002/HG003/HG004 reads + the VCF records the reference's own
    per-locus genotyper writes for them (tools/real_cases.py)."""
    import gzip
    with gzip.open(os.path.join(GOLD, "real_cases.json.gz"), "rt") as f:
        return json.load(f)["cases"]
