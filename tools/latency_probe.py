"""Per-locus latency of the reference-facing one-locus calls (what a per-locus drop-in pays per call)."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import synth
from longtr_b200 import Engine

eng = Engine(0)
loci = []
for s in range(60):
    loc = synth.make_locus(100 + s, n_reads=24, ref_len=60 + 4 * s, sub=0.002, indel=0.002)
    loci.append((synth.to_flat(loc), len(loc["reads"]), len(loc["alleles"])))
for rep in range(3):
    t0 = time.perf_counter()
    for (L, keep), P, H in loci:
        eng.process_reads_flat(L, P, H)
    dt = (time.perf_counter() - t0) / len(loci)
    print("process_reads_flat: %.3f ms per locus (%.0f loci/s)" % (dt * 1e3, 1 / dt), flush=True)
rng = np.random.default_rng(0)
ll = -rng.exponential(20, size=(24, 4)); p1 = np.full(24, -0.69); p2 = np.full(24, -0.69)
for rep in range(2):
    t0 = time.perf_counter()
    for _ in range(100):
        eng.genotype_locus(ll, p1, p2, [24])
    dt = (time.perf_counter() - t0) / 100
    print("genotype_locus: %.3f ms per locus" % (dt * 1e3), flush=True)
# the same loci handed over in ONE call (ltr_process_reads_flat_batch): one flattened GPU job for all of them
many = (loci * 34)[:2000]
flat = [x[0][0] for x in many]
shapes = [(x[1], x[2]) for x in many]
for rep in range(3):
    t0 = time.perf_counter()
    eng.process_reads_flat_batch(flat, shapes)
    dt = (time.perf_counter() - t0) / len(many)
    print("process_reads_flat_batch (2000 loci per call): %.4f ms per locus (%.0f loci/s)" % (dt * 1e3, 1 / dt), flush=True)
