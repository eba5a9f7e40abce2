"""Seeded inputs for the edit-distance / clustering parity tests (N2): pairs of TR-like strings and sets of distinct
strings the way HaplotypeGenerator::gen_candidate_seqs builds them (the skipped sequences of one sample, first element
kept, the rest ordered by length and sequence).  numpy only, no reference access."""
import numpy as np

BASES = "ACGT"
THRESHOLDS = [20, 50, 80, 100, 150, 200, 300, 400, 500, 600, 700]  # HaplotypeGenerator.cpp:403


def rand_seq(rng, n, alphabet=BASES):
    return "".join(alphabet[i] for i in rng.integers(0, len(alphabet), size=n))


def mutate(rng, s, sub, indel, alphabet=BASES):
    out = []
    for c in s:
        r = rng.random()
        if r < indel / 2:
            continue
        if r < indel:
            out.append(alphabet[rng.integers(0, len(alphabet))])
            out.append(c)
        elif r < indel + sub:
            out.append(alphabet[rng.integers(0, len(alphabet))])
        else:
            out.append(c)
    return "".join(out)


def tr_allele(rng, motif, copies, flank=0):
    return rand_seq(rng, flank) + motif * copies + rand_seq(rng, flank)


def pair_cases(seed=11, n_random=260):
    """[(cent, read, T)] -- includes empty strings, non-ACGT bytes, strings past one strip of either kernel (1024 / 256
    rows) and thresholds around the true distance."""
    rng = np.random.default_rng(seed)
    cases = [("", "", 20), ("", "ACGT", 20), ("ACGT", "", 20), ("A", "A", 0), ("A", "C", 0), ("A", "C", 1),
             ("ACGT" * 5, "ACGT" * 5, 20), ("ACGT" * 30, "ACGT" * 5, 20), ("ACGT" * 30, "ACGT" * 5, 100),
             ("ACGTN" * 8, "ACGTN" * 8, 20), ("ACGTN" * 8, "ACGTA" * 8, 20), ("acgt" * 8, "ACGT" * 8, 50),
             ("NNNNNNNN", "NNNNNNN", 20), ("RYKM" * 10, "RYKM" * 9 + "ACGT", 20)]
    # a row of the reference's test fires although the distance equals T: i leading bases, then a copy of the other string
    for i, T in ((3, 3), (7, 7), (20, 20), (50, 50)):
        body = rand_seq(rng, 40)
        cases.append(("G" * i + body.replace("G", "A"), body.replace("G", "A"), T))
        cases.append((body.replace("G", "A"), "G" * i + body.replace("G", "A"), T))
    for k in range(n_random):
        kind = k % 6
        if kind == 0:    # short STR-like
            motif = rand_seq(rng, int(rng.integers(1, 7)))
            a = tr_allele(rng, motif, int(rng.integers(5, 40)), int(rng.integers(0, 10)))
        elif kind == 1:  # VNTR-like, one strip of the Myers kernel
            motif = rand_seq(rng, int(rng.integers(10, 60)))
            a = tr_allele(rng, motif, int(rng.integers(5, 20)), 5)
        elif kind == 2:  # past 1024 rows
            motif = rand_seq(rng, int(rng.integers(20, 60)))
            a = tr_allele(rng, motif, int(rng.integers(25, 50)), 5)
        elif kind == 3:  # random strings
            a = rand_seq(rng, int(rng.integers(1, 400)))
        elif kind == 4:  # around the strip boundaries
            a = rand_seq(rng, int(rng.choice([31, 32, 33, 255, 256, 257, 1023, 1024, 1025, 2047, 2048, 2049])))
        else:
            a = rand_seq(rng, int(rng.integers(200, 700)), "ACGTN")
        rate = float(rng.choice([0.0, 0.005, 0.02, 0.05, 0.15]))
        b = mutate(rng, a, rate, rate)
        if rng.random() < 0.3:  # whole-motif length change
            cut = int(rng.integers(0, max(1, len(b) // 4)))
            b = b[cut:] if rng.random() < 0.5 else b + a[:cut]
        if rng.random() < 0.1:
            b = rand_seq(rng, len(a) + int(rng.integers(-5, 6)) if len(a) > 5 else 3)
        if rng.random() < 0.5:
            a, b = b, a
        cases.append((a, b, int(rng.choice(THRESHOLDS))))
    return cases


def at_threshold_cases(cases, distance):
    """The same pairs with T set to the true distance and its neighbours (distance: callable (a, b) -> int)."""
    out = []
    for a, b, _ in cases:
        if len(a) * len(b) == 0 or len(a) * len(b) > 600 * 600:
            continue
        d = distance(a, b)
        for T in (d - 1, d, d + 1):
            if 0 <= T <= 999:
                out.append((a, b, T))
    return out


def cluster_set(rng, n_alleles, per_allele, length, rate, motif_len=None):
    """Distinct noisy copies of n_alleles repeat alleles, in the reference's order: first string kept in front, the rest
    by (length, sequence) -- stringops.cpp orderByLengthAndSequence."""
    motif = rand_seq(rng, motif_len or int(rng.integers(2, 40)))
    copies = max(2, length // len(motif))
    seqs = set()
    for a in range(n_alleles):
        allele = tr_allele(rng, motif, copies + int(rng.integers(-4, 5)), 3)
        if rng.random() < 0.5:
            allele = mutate(rng, allele, 0.02, 0.0)
        for _ in range(per_allele):
            seqs.add(mutate(rng, allele, rate, rate))
    seqs = sorted(seqs)  # std::map key order
    rest = sorted(seqs[1:], key=lambda s: (len(s), s))
    return [seqs[0]] + rest


def cluster_cases(seed=5, n_sets=40):
    """[(seqs, T)]: sets that finish, sets that need more than 15 centroids at their threshold, tiny sets."""
    rng = np.random.default_rng(seed)
    out = [(["ACGTACGT"], 20), (["ACGTACGT", "ACGTACGA"], 20), (["A" * 10, "C" * 40, "G" * 80], 20)]
    for k in range(n_sets):
        length = int(rng.choice([60, 150, 400, 900, 1300] if k % 5 else [60, 150, 300]))
        n_alleles = int(rng.integers(1, 5)) if k % 5 else int(rng.integers(16, 22))
        per = int(rng.integers(2, 12)) if n_alleles < 10 else 2
        seqs = cluster_set(rng, n_alleles, per, length, float(rng.choice([0.002, 0.01, 0.03])))
        if k % 5 == 0:  # far-apart alleles: random strings cannot share a centroid at a small threshold
            seqs = seqs[:1] + sorted({rand_seq(rng, length + int(rng.integers(0, 15))) for _ in range(n_alleles)},
                                     key=lambda s: (len(s), s)) + seqs[1:]
            seqs = list(dict.fromkeys(seqs))
        out.append((seqs, int(rng.choice(THRESHOLDS[:6] if k % 5 else THRESHOLDS[:2]))))
    return out


def pack_sets(cases):
    """-> (all sequences, set_begin, set_items, set_T): every set gets its own copies of the strings."""
    seqs, begin, items, Ts = [], [0], [], []
    for s, T in cases:
        items += list(range(len(seqs), len(seqs) + len(s)))
        seqs += s
        begin.append(len(items))
        Ts.append(T)
    return seqs, np.array(begin, dtype=np.uint32), np.array(items, dtype=np.uint32), np.array(Ts, dtype=np.int32)
