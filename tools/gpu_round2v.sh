#!/usr/bin/env bash
# GPU visit r2v: N2 after the begin-kernel fix, config-4 margin sweep with the second band round, ncu capture of kernel 2.
out=gpurun_out; tag=r2v
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_edit.py -x -q > $out/${tag}_pytest_edit.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_edit.log
tail -3 $out/${tag}_pytest_edit.log
timeout 300 python bench.py --n2 --steps 3 --warmup 3 > $out/${tag}_bench_n2.json 2> $out/${tag}_bench_n2.err; tail -c 600 $out/${tag}_bench_n2.json
summ='import json,sys
d=json.load(sys.stdin); c=d["config"]
print(sys.argv[1], "value %.0f vit_ms %.1f banded %d uncert %d second %d frac %.3f cells %.1fG e2e %.0f" % (d["value"], c["viterbi_ms_per_step"], c["pairs_banded_per_gpu"], c["pairs_band_uncertified_per_gpu"], c.get("pairs_band_second_round_per_gpu",-1), d["roofline"]["frac"], c["cells_evaluated_per_gpu"]/1e9, d["e2e"]["value"]))'
for combo in "0.5 100" "0.4 100" "0.35 100" "0.45 90" "0.5 90" "0.55 100"; do
  set -- $combo
  LTR_BENCH_DEPTH=1 LTR_BAND_BUDGET=$1 LTR_BAND_RETRY_RHO=$2 timeout 300 python bench.py --config 4 --steps 2 --warmup 2 --no-cpu-baseline --no-raw --no-extra 2>$out/${tag}_c4.err | python -c "$summ" "c4 budget=$1 rho=$2" | tee -a $out/${tag}_c4_sweep.txt
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stutter_pair -c 1 -o $out/${tag}_stutter_full -f python bench.py --config 5 --loci 6000 --steps 1 --warmup 0 --no-cpu-baseline > $out/${tag}_stutter_ncu.log 2>&1
ncu -i $out/${tag}_stutter_full.ncu-rep --page raw --csv > $out/${tag}_stutter_raw.csv 2>/dev/null
ncu -i $out/${tag}_stutter_full.ncu-rep --page source --csv > $out/${tag}_stutter_source.csv 2>/dev/null
ls -la $out | tail -8
