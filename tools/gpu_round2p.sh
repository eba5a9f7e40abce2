#!/usr/bin/env bash
out=gpurun_out
python - <<'PY'
import sys, os, subprocess, time
sys.path.insert(0,'.'); sys.path.insert(0,'tests'); sys.path.insert(0,'tools')
import dropin_cases, golden_util
from oracle import pyoracle as po
cases = [dropin_cases.case_a4()] + dropin_cases.seeded_cases() + golden_util.load_real_cases()
cases = cases * int(os.environ.get("REP", "5"))
text = "".join(po._case_text(c) for c in cases).encode()
env = dict(os.environ); env["LONGTR_B200_TIMING"] = "1"
for which in ("gpu", "full"):
    t0 = time.perf_counter()
    p = subprocess.run(["oracle/_ref/ltr_ref_%s" % which], input=text, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    print(which, len(cases), "loci", time.perf_counter() - t0, "s", p.stderr.decode()[-400:])
PY
