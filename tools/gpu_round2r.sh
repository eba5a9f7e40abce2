#!/usr/bin/env bash
python - <<'PY'
import sys, os
sys.path.insert(0,'.'); sys.path.insert(0,'tests'); sys.path.insert(0,'tools')
import dropin_cases, golden_util
from oracle import pyoracle as po
cases = [dropin_cases.case_a4()] + dropin_cases.seeded_cases() + golden_util.load_real_cases()
open('/tmp/in5.txt','w').write("".join(po._case_text(c) for c in cases*5))
PY
which strace perf gdb ltrace 2>&1 | head
for b in full gpu; do
  echo "== $b"; ( /usr/bin/time -v oracle/_ref/ltr_ref_$b < /tmp/in5.txt > /dev/null ) 2>&1 | egrep "Elapsed|User time|System time|Voluntary|Involuntary|Maximum resident|Minor"
done
echo "== gpu with LONGTR_B200_TIMING"; LONGTR_B200_TIMING=1 oracle/_ref/ltr_ref_gpu < /tmp/in5.txt 2>&1 >/dev/null | tail -2
if which strace >/dev/null 2>&1; then strace -c -f -o /tmp/st.txt oracle/_ref/ltr_ref_gpu < /tmp/in5.txt > /dev/null 2>&1; head -15 /tmp/st.txt; fi
