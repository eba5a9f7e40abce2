#!/usr/bin/env bash
# GPU visit r2u: N2 kernels (tests, bench line, ncu capture), second band round on config 4 (margin / rho sweep), config 3 check.
out=gpurun_out; tag=r2u
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > $out/${tag}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
timeout 300 python bench.py --n2 --steps 3 --warmup 3 > $out/${tag}_bench_n2.json 2> $out/${tag}_bench_n2.err; tail -c 1500 $out/${tag}_bench_n2.json
summ='import json,sys
d=json.load(sys.stdin); c=d["config"]
print(sys.argv[1], "value %.0f vit_ms %.1f banded %d uncert %d frac %.3f cells %.1fG e2e %.0f" % (d["value"], c["viterbi_ms_per_step"], c["pairs_banded_per_gpu"], c["pairs_band_uncertified_per_gpu"], d["roofline"]["frac"], c["cells_evaluated_per_gpu"]/1e9, d["e2e"]["value"]))'
for combo in "0.6 0" "0.6 80" "0.6 100" "0.45 80" "0.45 100" "0.3 100" "0.75 80"; do
  set -- $combo
  LTR_BENCH_DEPTH=1 LTR_BAND_BUDGET=$1 LTR_BAND_RETRY_RHO=$2 timeout 300 python bench.py --config 4 --steps 2 --warmup 2 --no-cpu-baseline --no-raw --no-extra 2>$out/${tag}_c4.err | python -c "$summ" "c4 budget=$1 rho=$2" | tee -a $out/${tag}_c4_sweep.txt
done
for rho in 0 80; do
  LTR_BENCH_DEPTH=2 LTR_BAND_RETRY_RHO=$rho timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-raw --no-extra 2>$out/${tag}_c3.err | python -c "$summ" "c3 rho=$rho" | tee -a $out/${tag}_c4_sweep.txt
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:edit_myers -c 3 -o $out/${tag}_n2_full -f python bench.py --n2 --loci 256 --steps 1 --warmup 1 --no-cpu-baseline > $out/${tag}_n2_ncu.log 2>&1
ncu -i $out/${tag}_n2_full.ncu-rep --page raw --csv > $out/${tag}_n2_full_raw.csv 2>/dev/null
ls -la $out | tail -12
