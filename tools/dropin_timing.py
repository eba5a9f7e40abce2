"""Whole-locus timing of the drop-in: the reference's own per-locus genotyper (SeqStutterGenotyper ctor -> genotype ->
write_vcf_record, IO-less driver oracle/full_driver.cpp) all-CPU (`ltr_ref_full`) against the same code with
HapAligner::process_reads / Genotyper::calc_log_sample_posteriors bound to the GPU library through
integration/reference_binding.cpp (`ltr_ref_gpu`), one locus at a time like LongTR's region loop.
Run on the GPU box: python tools/dropin_timing.py [n_loci]."""
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", "tests"))
import dropin_cases  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def run(which, text):
    exe = os.path.join(HERE, "..", "oracle", "_ref", "ltr_ref_%s" % which)
    t0 = time.perf_counter()
    p = subprocess.run([exe], input=text, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=3000)
    dt = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError(p.stderr.decode()[-300:])
    return dt, p.stdout


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    import golden_util
    cases = [dropin_cases.case_a4()] + dropin_cases.seeded_cases() + golden_util.load_real_cases()
    cases = (cases * (n // len(cases) + 1))[:n]  # 11 seeded + 47 real (HG002 / trio) loci, repeated
    text = "".join(po._case_text(c) for c in cases).encode()
    one = po._case_text(cases[0]).encode()
    res = {}
    for which in ("full", "gpu"):
        t_one, _ = run(which, one)              # process start-up (+ CUDA context for the GPU build)
        t_all, out = run(which, text)
        res[which] = (t_one, t_all, out)
        print("%-4s: %d loci in %.3f s (start-up run with 1 locus: %.3f s) -> %.2f ms per locus after start-up" %
              (which, n, t_all, t_one, 1e3 * (t_all - t_one) / max(1, n - 1)))
    print("identical VCF records: %s" % (res["full"][2] == res["gpu"][2]))
    print("per-locus speed-up of the unbatched drop-in: %.1fx" %
          ((res["full"][1] - res["full"][0]) / max(1e-9, res["gpu"][1] - res["gpu"][0])))


if __name__ == "__main__":
    main()
