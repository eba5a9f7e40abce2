#!/usr/bin/env bash
python - <<'PY'
import sys, os, subprocess, time, resource
sys.path.insert(0,'.'); sys.path.insert(0,'tests'); sys.path.insert(0,'tools')
import dropin_cases, golden_util
from oracle import pyoracle as po
cases = [dropin_cases.case_a4()] + dropin_cases.seeded_cases() + golden_util.load_real_cases()
for rep in (1, 5):
    text = "".join(po._case_text(c) for c in cases*rep).encode()
    open('/tmp/in.txt','wb').write(text)
    for which in ("full", "gpu"):
        r0 = resource.getrusage(resource.RUSAGE_CHILDREN)
        t0 = time.perf_counter()
        with open('/tmp/in.txt','rb') as f:
            p = subprocess.run(["oracle/_ref/ltr_ref_%s" % which], stdin=f, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        dt = time.perf_counter() - t0
        r1 = resource.getrusage(resource.RUSAGE_CHILDREN)
        print(which, len(cases)*rep, "loci wall %.3f user %.3f sys %.3f minflt %d vcsw %d ivcsw %d maxrss %d" % (dt, r1.ru_utime-r0.ru_utime, r1.ru_stime-r0.ru_stime, r1.ru_minflt-r0.ru_minflt, r1.ru_nvcsw-r0.ru_nvcsw, r1.ru_nivcsw-r0.ru_nivcsw, r1.ru_maxrss))
PY
for lib in "" longtr_b200/csrc/variants/liblongtr_b200_u2.so longtr_b200/csrc/variants/liblongtr_b200_u4.so; do
  LONGTR_B200_LIB=$lib python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline --no-raw 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); print('lib=$lib', d['value'], d['config']['viterbi_ms_per_step'], d['roofline']['frac'])"
done
