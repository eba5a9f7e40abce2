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
sys.path.insert(0, HERE)
import dropin_cases  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def run(which, text, env=None):
    exe = os.path.join(HERE, "..", "oracle", "_ref", "ltr_ref_%s" % which)
    e = dict(os.environ)
    e.update(env or {})
    t0 = time.perf_counter()
    p = subprocess.run([exe], input=text, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=3000, env=e)
    dt = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError(p.stderr.decode()[-300:])
    return dt, p.stdout


def per_locus(which, env, cases, reps=(1, 5)):
    """Seconds per locus as a difference quotient: (T(5N) - T(N)) / 4N, best of two runs each -- the CUDA start-up of
    the GPU build (1-3 s, varies from box to box and run to run) cancels instead of being estimated separately."""
    t = {}
    out = None
    for r in reps:
        text = "".join(po._case_text(c) for c in cases * r).encode()
        t[r], out = min((run(which, text, env) for _ in range(2)), key=lambda x: x[0])
    n = len(cases)
    return (t[reps[1]] - t[reps[0]]) / (n * (reps[1] - reps[0])), t, out


def main():
    import golden_util
    cases = [dropin_cases.case_a4()] + dropin_cases.seeded_cases() + golden_util.load_real_cases()
    eager = {"LONGTR_B200_EAGER_HAP_ALIGNMENT": "1"}
    timing = {"LONGTR_B200_TIMING": "1"}
    rows = (("all-CPU reference (ltr_ref_full)", "full", None), ("GPU drop-in, NW as in the reference", "gpu", eager),
            ("GPU drop-in, NW elided", "gpu", None))
    print("11 seeded + 47 real HG002 / trio loci, N = %d and 5N loci per run, per-locus time = (T(5N) - T(N)) / 4N" % len(cases))
    res = {}
    for label, which, env in rows:
        dt, t, out = per_locus(which, env, cases)
        res[label] = (dt, out)
        print("%-38s %7.3f ms per locus   (runs: %.3f s / %.3f s)" % (label, 1e3 * dt, t[1], t[5]))
    outs = [v[1] for v in res.values()]
    print("identical VCF records: %s" % (outs[0] == outs[1] == outs[2]))
    base = res[rows[0][0]][0]
    for label, _w, _e in rows[1:]:
        print("per-locus speed-up of the unbatched drop-in (%s): %.2fx" % (label, base / res[label][0]))
    text = "".join(po._case_text(c) for c in cases * 5).encode()
    e = dict(os.environ)
    e.update(timing)
    p = subprocess.run([os.path.join(HERE, "..", "oracle", "_ref", "ltr_ref_gpu")], input=text, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, env=e)
    print(p.stderr.decode().strip().splitlines()[-1])
    # VNTR-sized loci (~1 kb repeats, ONT-like parameters): where Haplotype::aln_haps_to_ref is what is left on the host
    import nw_elision_timing as nw
    v = nw.vntr_cases(4)
    print("VNTR ~1 kb loci, N = 4 and 3N loci per run")
    vres = {}
    for label, which, env in rows:
        dt, t, out = per_locus(which, env, v, reps=(1, 3))
        vres[label] = out
        print("%-38s %7.1f ms per locus   (runs: %.3f s / %.3f s)" % (label, 1e3 * dt, t[1], t[3]))
    outs = list(vres.values())
    print("VNTR identical VCF records: %s" % (outs[0] == outs[1] == outs[2]))


if __name__ == "__main__":
    main()
