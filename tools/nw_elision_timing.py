"""N1: what Haplotype::aln_haps_to_ref costs per locus and what eliding it changes.
All-CPU: the reference's per-locus genotyper as is (`ltr_ref_full`) against the same objects with the member replaced by
integration/lazy_haplotype_alignment.cpp (`ltr_ref_lazy`), and -- same binary, env switch -- with the original re-enabled.
VCF records must be identical.  Usage: python tools/nw_elision_timing.py [repeats]   (no GPU needed)"""
import os
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", "tests"))
import dropin_cases as dc  # noqa: E402
import golden_util  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def run(which, text, env=None):
    exe = os.path.join(HERE, "..", "oracle", "_ref", "ltr_ref_%s" % which)
    e = dict(os.environ)
    e.update(env or {})
    t0 = time.perf_counter()
    p = subprocess.run([exe], input=text, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=6000, env=e)
    dt = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError(p.stderr.decode()[-300:])
    return dt, p.stdout


def vntr_cases(n, units=33, period=30):
    """VNTR-sized loci: ~1 kb repeats of a 30-base motif, several alleles per sample set, ONT-like parameters."""
    ONT = (-1.0, -0.458675, -1.0, -0.458675, -0.0202027, -4.60517, -4.60517)
    out = []
    for seed in range(n):
        rng = np.random.default_rng(8800 + seed)
        rnd = lambda k: "".join("ACGT"[int(x)] for x in rng.integers(0, 4, size=k))
        motif = rnd(period)
        chrom = rnd(1000) + motif * units + rnd(1000)
        offs = [(int(rng.integers(-3, 4)), int(rng.integers(-3, 4))) for _ in range(3)]
        out.append(dc.make_case("vntr%02d" % seed, chrom, motif, units, offs, reads_per_sample=10, params=ONT, lo=400, span=600, exact_cigar=True))
    return out


def measure(name, cases):
    text = "".join(po._case_text(c) for c in cases).encode()
    rows = []
    for label, which, env in (("reference as is (ltr_ref_full)", "full", None),
                              ("NW elided (ltr_ref_lazy)", "lazy", None),
                              ("ltr_ref_lazy, original re-enabled", "lazy", {"LONGTR_B200_EAGER_HAP_ALIGNMENT": "1"})):
        dt, out = min((run(which, text, env) for _ in range(3)), key=lambda x: x[0])  # best of three
        rows.append((label, dt, out))
        print("%-16s %-36s %4d loci  %8.3f s  %9.3f ms / locus" % (name, label, len(cases), dt, 1e3 * dt / len(cases)))
    same = rows[0][2] == rows[1][2] == rows[2][2]
    print("%-16s identical VCF records: %s;  per-locus time saved by the elision: %.3f ms (%.1f %% of the all-CPU locus)" %
          (name, same, 1e3 * (rows[0][1] - rows[1][1]) / len(cases), 100.0 * (rows[0][1] - rows[1][1]) / rows[0][1]))
    return same


def main():
    rep = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    print("python tools/nw_elision_timing.py %d   (all-CPU, best of three runs per line, %d host cores visible)" %
          (rep, os.cpu_count() or 1))
    std = [dc.case_a4()] + dc.seeded_cases() + dc.pruning_cases() + golden_util.load_real_cases()
    ok = measure("STR loci", std * rep)
    ok = measure("VNTR ~1 kb", vntr_cases(3)) and ok
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
