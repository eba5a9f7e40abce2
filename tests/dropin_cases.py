"""Seeded IO-less loci for the drop-in check (reference SeqStutterGenotyper -> VCF record), SURVEY.md Appendix A4."""
import numpy as np


class MT19937:
    """std::mt19937 (32-bit), so that case A4 reproduces SURVEY Appendix A4 character by character."""

    def __init__(self, seed):
        self.mt = [0] * 624
        self.idx = 624
        self.mt[0] = seed & 0xFFFFFFFF
        for i in range(1, 624):
            self.mt[i] = (1812433253 * (self.mt[i - 1] ^ (self.mt[i - 1] >> 30)) + i) & 0xFFFFFFFF

    def __call__(self):
        if self.idx >= 624:
            for i in range(624):
                y = (self.mt[i] & 0x80000000) | (self.mt[(i + 1) % 624] & 0x7FFFFFFF)
                self.mt[i] = self.mt[(i + 397) % 624] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
            self.idx = 0
        y = self.mt[self.idx]
        self.idx += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF


def make_case(name, chrom, motif, units, sample_offsets, reads_per_sample=14, lo=800, span=200, qual="I",
              haploid=False, params=None, extra=None, exact_cigar=False):
    """chrom = left(1000) + motif*units + right(1000); sample_offsets[s] = (k_hp0, k_hp1) repeat-unit offsets.
    extra[s] = (k, count): the last `count` reads of sample s carry a third allele (offset k) and no phase information --
    enough support to become a candidate haplotype, not enough to be called (the allele-pruning cases)."""
    p = len(motif)
    rs, re = 1000, 1000 + p * units
    reads, n1, n2 = [], [], []
    for s, (k0, k1) in enumerate(sample_offsets):
        c1 = c2 = 0
        for r in range(reads_per_sample):
            hp = r % 2
            k = k0 if hp == 0 else k1
            unphased = False
            if extra is not None and extra[s] is not None and r >= reads_per_sample - extra[s][1]:
                k, unphased = extra[s][0], True
            seq = chrom[lo:rs] + motif * (units + k) + chrom[re:re + span]
            left, right = rs - lo, span
            if k > 0:
                # (the historical form stops the last '=' after `span` bases -- the recorded fixtures, SURVEY A4 included,
                # were made with it and short repeats never notice; exact_cigar covers the whole read)
                cigar, aln = "%d=%dI%d=" % (left, p * k, right + (p * units if exact_cigar else 0)), seq
            elif k < 0:
                cigar = "%d=%dD%d=" % (left, -p * k, right + p * k + (p * units - (-p * k)) - (p * units + p * k) + 0)
                cigar = "%d=%dD%d=" % (left, -p * k, len(seq) - left)
                aln = seq[:left] + "-" * (-p * k) + seq[left:]
            else:
                cigar, aln = "%d=" % len(seq), seq
            # left-aligned indel placement: the indel sits at the first repeat unit
            reads.append(dict(start=lo, stop=re + span - 1, rev=(r % 3 == 0), sample=s, name="read%d_%d" % (s, r),
                              seq=seq, qual=qual * len(seq), aln=aln, cigar=cigar,
                              log_p1=(-0.6931471805599453 if unphased else (-1e-6 if hp == 0 else -1000.0)),
                              log_p2=(-0.6931471805599453 if unphased else (-1000.0 if hp == 0 else -1e-6))))
            c1 += hp == 0
            c2 += hp == 1
        n1.append(c1)
        n2.append(c2)
    return dict(name=name, chrom_name="chrT", chrom_seq=chrom, region_start=rs, region_stop=re, motif=motif,
                region_name="locus1", samples=["S%d" % (s + 1) for s in range(len(sample_offsets))], n_p1s=n1, n_p2s=n2,
                reads=reads, stutter_motif="A", stutter_period=p, haploid=haploid, aln_params=params)


def case_a4():
    g = MT19937(7)
    rnd = lambda n: "".join("ACGT"[g() & 3] for _ in range(n))
    left = rnd(1000)
    chrom = left + "CAG" * 15 + rnd(1000)
    return make_case("A4", chrom, "CAG", 15, [(0, 3), (-2, -2)])


A4_RECORD = ("chrT\t1001\tlocus1\tCAGCAGCAGCAGCAGCAGCAGCAGCAGCAGCAGCAGCAGCAGCAG\t"
             "CAGCAGCAGCAGCAGCAGCAGCAGCAGCAGCAGCAGCAG,CAGCAGCAGCAGCAGCAGCAGCAGCAGCAGCAGCAGCAGCAGCAGCAGCAGCAG\t.\t.\t"
             "START=1001;END=1045;MOTIF=CAG;PERIOD=3;NSKIP=0;NFILT=0;INEXACT_ALLELE=0,0;BPDIFFS=-6,9;DP=28;DSNP=28;"
             "DFLANKINDEL=0;AN=4;REFAC=1;AC=2,1\tGT:GB:Q:PQ:DP:DSNP:DFLANKINDEL:PDP:PSNP:GLDIFF:ALLREADS:MALLREADS\t"
             "0|2:0|9:1.00:1.00:14:14:0:7|7:7|7:48.35:0|7;9|7:0|7;9|7\t1|1:-6|-6:1.00:1.00:14:14:0:7|7:7|7:48.36:-6|14:-6|14")


def seeded_cases():
    out = []
    ONT = (-1.0, -0.458675, -1.0, -0.458675, -0.0202027, -4.60517, -4.60517)
    for seed in range(10):
        rng = np.random.default_rng(4400 + seed)
        rnd = lambda n: "".join("ACGT"[int(x)] for x in rng.integers(0, 4, size=n))
        p = int(rng.integers(2, 7))
        while True:
            motif = rnd(p)
            if all(motif != motif[:q] * (p // q) for q in range(1, p) if p % q == 0):
                break
        units = int(rng.integers(6, 30))
        chrom = rnd(1000) + motif * units + rnd(1000)
        S = int(rng.integers(1, 4))
        offs = [(int(rng.integers(-3, 4)), int(rng.integers(-3, 4))) for _ in range(S)]
        offs = [(max(a, 2 - units), max(b, 2 - units)) for a, b in offs]
        out.append(make_case("dropin%02d" % seed, chrom, motif, units, offs, reads_per_sample=int(rng.integers(10, 17)),
                             params=ONT if seed % 4 == 3 else None))
    return out


def haploid_cases():
    """The first five seeded loci genotyped as a haploid chromosome (--haploid-chrs): haploid priors, GT / GL / PL over single
    alleles, no PHASEDGL."""
    return [dict(c, name=c["name"] + "_haploid", haploid=True) for c in seeded_cases()[:5]]


def pruning_cases():
    """Loci in which some candidate allele ends up in no sample's optimal pair: SeqStutterGenotyper::genotype drops it
    and recomputes the posteriors (src/seq_stutter_genotyper.cpp:636-645)."""
    out = []
    for seed in range(8):
        rng = np.random.default_rng(5500 + seed)
        rnd = lambda n: "".join("ACGT"[int(x)] for x in rng.integers(0, 4, size=n))
        p = int(rng.integers(2, 6))
        while True:
            motif = rnd(p)
            if all(motif != motif[:q] * (p // q) for q in range(1, p) if p % q == 0):
                break
        units = int(rng.integers(8, 24))
        chrom = rnd(1000) + motif * units + rnd(1000)
        S = int(rng.integers(1, 4))
        offs = [(int(rng.integers(-2, 3)), int(rng.integers(-2, 3))) for _ in range(S)]
        used = {k for o in offs for k in o}
        extra = []
        for s in range(S):
            cand = [k for k in range(-4, 5) if k not in used and k != 0 and units + k >= 2]
            extra.append((int(rng.choice(cand)), int(rng.integers(3, 5))) if (s == 0 or rng.random() < 0.5) else None)
        out.append(make_case("prune%02d" % seed, chrom, motif, units, offs, reads_per_sample=int(rng.integers(14, 20)),
                             extra=extra))
    return out
