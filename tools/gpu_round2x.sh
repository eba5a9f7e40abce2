#!/usr/bin/env bash
# GPU visit r2x: config-4 band share threshold sweep, ncu capture of the K=12 two-strip stream kernel and of a whole-warp band class.
out=gpurun_out; tag=r2x
mkdir -p $out
summ='import json,sys
d=json.load(sys.stdin); c=d["config"]
print(sys.argv[1], "value %.0f vit_ms %.1f banded %d uncert %d second %d frac %.3f cells %.1fG e2e %.0f" % (d["value"], c["viterbi_ms_per_step"], c["pairs_banded_per_gpu"], c["pairs_band_uncertified_per_gpu"], c.get("pairs_band_second_round_per_gpu",-1), d["roofline"]["frac"], c["cells_evaluated_per_gpu"]/1e9, d["e2e"]["value"]))'
for combo in "55 100" "65 100" "75 100" "85 100" "75 90"; do
  set -- $combo
  LTR_BENCH_DEPTH=1 LTR_BAND_MAX_SHARE=$1 LTR_BAND_RETRY_RHO=$2 timeout 300 python bench.py --config 4 --steps 2 --warmup 2 --no-cpu-baseline --no-raw --no-extra 2>$out/${tag}_c4.err | python -c "$summ" "c4 share=$1 rho=$2" | tee -a $out/${tag}_sweep.txt
done
for s in 55 75; do
  LTR_BENCH_DEPTH=2 LTR_BAND_MAX_SHARE=$s timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-raw --no-extra 2>$out/${tag}_c3.err | python -c "$summ" "c3 share=$s" | tee -a $out/${tag}_sweep.txt
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"viterbi_stream_kernel<12" -c 1 -o $out/${tag}_stream12_full -f python bench.py --config 4 --loci 2000 --steps 1 --warmup 0 --no-cpu-baseline --no-raw --no-extra > $out/${tag}_stream12_ncu.log 2>&1
ncu -i $out/${tag}_stream12_full.ncu-rep --page raw --csv > $out/${tag}_stream12_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"viterbi_band_kernel<4, 32" -c 1 -o $out/${tag}_band432_full -f python bench.py --config 4 --loci 2000 --steps 1 --warmup 0 --no-cpu-baseline --no-raw --no-extra > $out/${tag}_band432_ncu.log 2>&1
ncu -i $out/${tag}_band432_full.ncu-rep --page raw --csv > $out/${tag}_band432_raw.csv 2>/dev/null
ls -la $out | tail -8
