"""Records tests/golden/em.json: EMStutterGenotyper(...).train(100, 0.01, 0.001) of the REFERENCE, compiled in place
(oracle/em_driver.cpp -> oracle/_ref/libltr_ref_em.so), on the seeded loci of tests/em_cases.py: trained flag, the six model
parameters, allele log-frequencies, iteration count and final log-likelihood (doubles as hex floats).

    python tools/make_golden_em.py          (needs oracle/_ref; no GPU)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import em_cases  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

N = 120


def main():
    assert po.ref_em_available()
    cases = []
    for seed in range(N):
        L = em_cases.em_locus(seed)
        r = po.ref_em_train(L["reads_per_sample"], L["bp_diff"], L["log_p1"], L["log_p2"], L["motif_len"], L["haploid"])
        cases.append(dict(seed=seed, trained=r["trained"], n_iter=r["n_iter"], params=[float(x).hex() for x in r["params"]],
                          ll=float(r["lls"][-1]).hex(), log_gt_priors=[float(x).hex() for x in r["log_gt_priors"]]))
    short = [dict(seed=1000 + k, max_iter=2) for k in range(10)]   # not converged within two iterations: trained = false
    for c in short:
        L = em_cases.em_locus(c["seed"])
        r = po.ref_em_train(L["reads_per_sample"], L["bp_diff"], L["log_p1"], L["log_p2"], L["motif_len"], L["haploid"],
                            max_iter=c["max_iter"])
        c.update(trained=r["trained"], n_iter=r["n_iter"], params=[float(x).hex() for x in r["params"]],
                 ll=float(r["lls"][-1]).hex(), log_gt_priors=[float(x).hex() for x in r["log_gt_priors"]])
    path = os.path.join(ROOT, "tests", "golden", "em.json")
    json.dump(dict(generator="tools/make_golden_em.py", source="oracle/_ref/libltr_ref_em.so (reference compiled in place)",
                   cases=cases, short=short), open(path, "w"), separators=(",", ":"))
    print(path, os.path.getsize(path), "bytes;", sum(c["trained"] for c in cases), "of", len(cases), "trained;",
          "iterations", min(c["n_iter"] for c in cases), "-", max(c["n_iter"] for c in cases),
          "; short runs trained:", sum(c["trained"] for c in short))


if __name__ == "__main__":
    main()
