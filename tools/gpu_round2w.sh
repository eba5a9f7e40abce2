#!/usr/bin/env bash
# GPU visit r2w: 11 band classes + second round (full GPU suite, config-4 margin sweep, config-3 check, config-4 launch list).
out=gpurun_out; tag=r2w
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
summ='import json,sys
d=json.load(sys.stdin); c=d["config"]
print(sys.argv[1], "value %.0f vit_ms %.1f banded %d uncert %d second %d frac %.3f cells %.1fG e2e %.0f" % (d["value"], c["viterbi_ms_per_step"], c["pairs_banded_per_gpu"], c["pairs_band_uncertified_per_gpu"], c.get("pairs_band_second_round_per_gpu",-1), d["roofline"]["frac"], c["cells_evaluated_per_gpu"]/1e9, d["e2e"]["value"]))'
for combo in "0.45 100" "0.5 100" "0.55 100" "0.45 80" "0.6 100"; do
  set -- $combo
  LTR_BENCH_DEPTH=1 LTR_BAND_BUDGET=$1 LTR_BAND_RETRY_RHO=$2 timeout 300 python bench.py --config 4 --steps 2 --warmup 2 --no-cpu-baseline --no-raw --no-extra 2>$out/${tag}_c4.err | python -c "$summ" "c4 budget=$1 rho=$2" | tee -a $out/${tag}_sweep.txt
done
for b in 0.45 0.6; do
  LTR_BENCH_DEPTH=2 LTR_BAND_BUDGET=$b timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-raw --no-extra 2>$out/${tag}_c3.err | python -c "$summ" "c3 budget=$b" | tee -a $out/${tag}_sweep.txt
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_c4.csv python bench.py --config 4 --loci 2000 --steps 1 --warmup 1 --no-cpu-baseline --no-raw --no-extra > $out/${tag}_launches_c4.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/r2w_launches_c4.csv")) if len(r)>10]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
t=collections.Counter(); n=collections.Counter()
for r in rows[1:]:
    t[r[ki]]+=float(r[vi].replace(",","")); n[r[ki]]+=1
tot=sum(t.values())
for k,v in t.most_common(16): print("%-70s n=%3d %.3f ms share %.3f"%(k[:70],n[k],v/1e6,v/tot))
PY
ls -la $out | tail -6
