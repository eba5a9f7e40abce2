"""Seeded synthetic TR loci for the parity tests (numpy only, no reference access).

A locus is generated the way SURVEY.md section 8(d) describes config C3/C4: a random
non-periodic motif, a reference repeat, random flanks, two true alleles that differ by
whole repeat units, and reads drawn from the alleles with substitution / indel errors and
a '=XID' CIGAR against the reference window.
"""
import numpy as np

BASES = "ACGT"


def rand_seq(rng, n):
    return "".join(BASES[i] for i in rng.integers(0, 4, size=n))


def rand_motif(rng, period):
    while True:
        m = rand_seq(rng, period)
        if period == 1 or all(m != m[:q] * (period // q) for q in range(1, period) if period % q == 0):
            return m


def rle(ops):
    out, prev, cnt = [], None, 0
    for o in ops:
        if o == prev:
            cnt += 1
        else:
            if prev is not None:
                out.append("%d%s" % (cnt, prev))
            prev, cnt = o, 1
    if prev is not None:
        out.append("%d%s" % (cnt, prev))
    return "".join(out)


def simulate_read(rng, window, win_start, rep_lo, rep_hi, allele, sub=1e-3, indel=1e-3, qual_lo=20, qual_hi=40,
                  var_at_end=True):
    """Read from haplotype window[:rep_lo]+allele+window[rep_hi:] (indices into window)."""
    ref_rep = window[rep_lo:rep_hi]
    d = len(allele) - len(ref_rep)
    ev = [("=", c) for c in window[:rep_lo]]
    # allele vs reference repeat: shared prefix/suffix, length change placed at one end
    if d >= 0:
        if var_at_end:
            ev += [("=" if a == b else "X", a) for a, b in zip(allele[:len(ref_rep)], ref_rep)]
            ev += [("I", c) for c in allele[len(ref_rep):]]
        else:
            ev += [("I", c) for c in allele[:d]]
            ev += [("=" if a == b else "X", a) for a, b in zip(allele[d:], ref_rep)]
    else:
        if var_at_end:
            ev += [("=" if a == b else "X", a) for a, b in zip(allele, ref_rep[:len(allele)])]
            ev += [("D", None)] * (-d)
        else:
            ev += [("D", None)] * (-d)
            ev += [("=" if a == b else "X", a) for a, b in zip(allele, ref_rep[-d:])]
    ev += [("=", c) for c in window[rep_hi:]]
    # sequencing errors (never on the first / last 3 events so start/stop stay put)
    out = []
    for k, (op, c) in enumerate(ev):
        edge = k < 3 or k >= len(ev) - 3
        if op in "=X" and not edge:
            u = rng.random()
            if u < sub:
                alt = BASES[(BASES.index(c) + int(rng.integers(1, 4))) % 4]
                ref_base = None
                out.append(("X", alt))
                continue
            if u < sub + indel / 2:
                out.append(("D", None))
                continue
            if u < sub + indel:
                out.append((op, c))
                out.append(("I", BASES[int(rng.integers(0, 4))]))
                continue
        out.append((op, c))
    seq = "".join(c for op, c in out if op != "D")
    ops = [op for op, _ in out]
    n_ref = sum(1 for op in ops if op in "=XD")
    qual = "".join(chr(33 + int(q)) for q in rng.integers(qual_lo, qual_hi + 1, size=len(seq)))
    return dict(start=win_start, stop=win_start + n_ref - 1, seq=seq, qual=qual, cigar=rle(ops))


def make_locus(seed, period=None, ref_len=None, n_reads=12, ctx=60, pad=5, flank=35, sub=1e-3, indel=1e-3,
               n_decoys=1, max_units=3, pos0=1000, homopolymer=False):
    """Returns a dict describing one locus (strings + reads) -- feed to flat.make_flat_locus."""
    rng = np.random.default_rng(seed)
    if period is None:
        period = 1 if homopolymer else int(rng.integers(1, 7))
    motif = rand_motif(rng, period)
    if ref_len is None:
        ref_len = int(rng.integers(12, 120))
    units = max(2, ref_len // period)
    # repeat block = pad + repeat + pad (HaplotypeGenerator keeps the padding inside the block, Appendix A)
    lctx, rctx = rand_seq(rng, ctx), rand_seq(rng, ctx)
    lflank, rflank = rand_seq(rng, flank), rand_seq(rng, flank)
    lpad, rpad = rand_seq(rng, pad), rand_seq(rng, pad)
    ref_allele = lpad + motif * units + rpad
    ks = sorted(set(int(k) for k in rng.integers(-max_units, max_units + 1, size=2)))
    true_alleles = [lpad + motif * max(1, units + k) + rpad for k in ks]
    decoys = []
    for _ in range(n_decoys):
        k = int(rng.integers(-max_units - 2, max_units + 3))
        decoys.append(lpad + motif * max(1, units + k) + rpad)
    alts = sorted(set(true_alleles + decoys) - {ref_allele}, key=lambda s: (len(s), s))
    alleles = [ref_allele] + alts
    window = lctx + lflank + ref_allele + rflank + rctx
    win_start = pos0 - ctx - flank
    rep_lo = ctx + flank
    rep_hi = rep_lo + len(ref_allele)
    reads = []
    for r in range(n_reads):
        a = true_alleles[int(rng.integers(0, len(true_alleles)))]
        reads.append(simulate_read(rng, window, win_start, rep_lo, rep_hi, a, sub=sub, indel=indel,
                                   var_at_end=bool(rng.integers(0, 2))))
    return dict(lflank=lflank, rflank=rflank, alleles=alleles, repeat_start=pos0,
                repeat_end=pos0 + len(ref_allele), period=period, motif=motif, reads=reads,
                true_alleles=true_alleles)


def to_flat(loc, **kw):
    from longtr_b200.flat import make_flat_locus
    reads = [(r["start"], r["stop"], r["seq"], r["qual"], r["cigar"]) for r in loc["reads"]]
    return make_flat_locus(loc["lflank"], loc["alleles"], loc["rflank"], loc["repeat_start"],
                           loc["repeat_end"], loc["period"], reads, motif=loc["motif"], **kw)


def make_pair_batch(seed, n_loci, n_lo=20, n_hi=200, reads_lo=1, reads_hi=6, haps_lo=1, haps_hi=4,
                    sub=0.01, indel=0.01, flank=30, weird=0.1):
    """Flattened kernel-level batch: full haplotypes (flank+allele+flank) and already trimmed reads.

    ``weird`` is the fraction of reads replaced by unrelated / very short / very long sequences
    so that sentinel paths (-700 bail-out, |n-m|>600, m==1, m>n) are exercised.
    """
    rng = np.random.default_rng(seed)
    lhb, lrb, hoff, roff = [0], [0], [0], [0]
    hbytes, rbytes = [], []
    for _ in range(n_loci):
        period = int(rng.integers(1, 7))
        motif = rand_motif(rng, period)
        core_len = int(rng.integers(n_lo, n_hi + 1))
        units = max(1, core_len // period)
        lf, rf = rand_seq(rng, flank + 5), rand_seq(rng, flank + 5)
        H = int(rng.integers(haps_lo, haps_hi + 1))
        alleles = []
        for h in range(H):
            k = 0 if h == 0 else int(rng.integers(-4, 5))
            alleles.append(motif * max(1, units + k))
        for a in alleles:
            s = lf + a + rf
            if rng.random() < weird * 0.3:
                s = s[:int(rng.integers(40, 64))]  # exercises the <=60 sentinel
            hbytes.append(s)
            hoff.append(hoff[-1] + len(s))
        P = int(rng.integers(reads_lo, reads_hi + 1))
        for _r in range(P):
            a = alleles[int(rng.integers(0, H))]
            s = list(lf[flank:] + a + rf[:5])
            out = []
            for c in s:
                u = rng.random()
                if u < sub:
                    out.append(BASES[int(rng.integers(0, 4))])
                elif u < sub + indel / 2:
                    continue
                elif u < sub + indel:
                    out.append(c)
                    out.append(BASES[int(rng.integers(0, 4))])
                else:
                    out.append(c)
            s = "".join(out) or "A"
            u = rng.random()
            if u < weird * 0.4:
                s = rand_seq(rng, int(rng.integers(1, 4)))
            elif u < weird * 0.7:
                s = rand_seq(rng, int(rng.integers(5, 2 * n_hi)))
            elif u < weird:
                s = rand_seq(rng, len(s) + int(rng.integers(590, 640)))
            rbytes.append(s)
            roff.append(roff[-1] + len(s))
        lhb.append(len(hoff) - 1)
        lrb.append(len(roff) - 1)
    return dict(locus_hap_begin=np.array(lhb, dtype=np.uint32), locus_read_begin=np.array(lrb, dtype=np.uint32),
                hap_off=np.array(hoff, dtype=np.uint32), read_off=np.array(roff, dtype=np.uint32),
                hap_bytes=np.frombuffer("".join(hbytes).encode(), dtype=np.uint8).copy(),
                read_bytes=np.frombuffer("".join(rbytes).encode(), dtype=np.uint8).copy(),
                haps=hbytes, reads=rbytes)


def make_pathological_batch(seed):
    """Loci without reads or without haplotypes, haplotypes of 0..62 bases next to 700-base ones, duplicated / one-base /
    1 500-base reads (kernel-level batch like make_pair_batch)."""
    rng = np.random.default_rng(99 + seed)
    lhb, lrb, hoff, roff, hb, rb = [0], [0], [0], [0], [], []
    for _l in range(int(rng.integers(1, 8))):
        H, R = int(rng.integers(0, 4)), int(rng.integers(0, 7))
        n = int(rng.choice([0, 1, 2, 5, 40, 61, 62, 90, 200, 700]))
        base = rand_seq(rng, n + 60) if n > 0 else rand_seq(rng, int(rng.integers(0, 61)))
        for _h in range(H):
            s = base if rng.random() < 0.6 else rand_seq(rng, int(rng.integers(0, 130)))
            hb.append(s)
            hoff.append(hoff[-1] + len(s))
        pool = []
        for _r in range(R):
            u = rng.random()
            if pool and u < 0.4:
                s = pool[int(rng.integers(0, len(pool)))]
            elif u < 0.5:
                s = "A"
            elif u < 0.6:
                s = rand_seq(rng, int(rng.integers(700, 1500)))
            else:
                core = base[30:30 + max(1, n)] if len(base) > 60 else "ACGT"
                s = core[:max(1, len(core) - int(rng.integers(0, 5)))] + rand_seq(rng, int(rng.integers(0, 4)))
            pool.append(s)
            rb.append(s)
            roff.append(roff[-1] + len(s))
        lhb.append(len(hb))
        lrb.append(len(rb))
    return dict(locus_hap_begin=np.array(lhb, np.uint32), locus_read_begin=np.array(lrb, np.uint32),
                hap_off=np.array(hoff, np.uint32), read_off=np.array(roff, np.uint32),
                hap_bytes=np.frombuffer("".join(hb).encode(), np.uint8).copy() if hb else np.zeros(0, np.uint8),
                read_bytes=np.frombuffer("".join(rb).encode(), np.uint8).copy() if rb else np.zeros(0, np.uint8))
