"""Seeded loci for the stutter-model EM tests: 1-3 samples, 3-40 reads each, motif length 1-6, diploid or haploid genotypes
within +-3 repeat units, in-frame and out-of-frame stutter, phased (HP-like) and unphased reads."""
import numpy as np


def em_locus(seed, haploid=None, big=False):
    rng = np.random.default_rng(20261017 + seed)
    if haploid is None:
        haploid = seed % 7 == 3
    S = int(rng.integers(1, 4))
    motif = int(rng.integers(1, 7))
    rps = [int(rng.integers(3, 41 if not big else 400)) for _ in range(S)]
    up, down, off = rng.uniform(0.01, 0.12), rng.uniform(0.01, 0.15), rng.uniform(0.0, 0.06)
    bd, p1, p2 = [], [], []
    for s in range(S):
        g = rng.integers(-3, 4, size=2) * motif
        if haploid:
            g[1] = g[0]
        for _ in range(rps[s]):
            h = int(rng.integers(0, 2))
            b = int(g[h])
            u = rng.random()
            if u < up:
                b += motif * int(rng.integers(1, 3))
            elif u < up + down:
                b -= motif * int(rng.integers(1, 3))
            elif u < up + down + off:
                b += int(rng.integers(-2, 3))
            bd.append(b)
            if rng.random() < 0.7:
                p1.append(-1e-6 if h == 0 else -1000.0)
                p2.append(-1000.0 if h == 0 else -1e-6)
            else:
                p1.append(0.0)
                p2.append(0.0)
    return dict(reads_per_sample=rps, bp_diff=bd, log_p1=p1, log_p2=p2, motif_len=motif, haploid=bool(haploid))
