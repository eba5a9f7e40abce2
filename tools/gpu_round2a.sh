#!/usr/bin/env bash
# Round-2 first visit: ncu captures of kernel 2 (stutter_pair_kernel, config 5) and of the full-matrix stream kernel on
# config 4 (the K = 12 two-strip instance), and a timing of compute-sanitizer over part of the GPU suite.
tag="${1:-r2a}"
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $out/${tag}_smi.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stutter_pair_kernel -c 2 -f -o /tmp/${tag}_c5 \
    python bench.py --config 5 --steps 1 --warmup 1 --loci 4000 --no-cpu-baseline > $out/${tag}_c5_bench.log 2>&1
ncu -i /tmp/${tag}_c5.ncu-rep --page raw --csv > $out/${tag}_c5_raw.csv 2>/dev/null
ncu -i /tmp/${tag}_c5.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $out/${tag}_c5_sass.csv.gz
ncu -i /tmp/${tag}_c5.ncu-rep --page source --csv --print-source cuda 2>/dev/null | gzip > $out/${tag}_c5_cuda.csv.gz
timeout 900 ncu --set full --clock-control none --import-source on -k regex:viterbi_stream_kernel -c 12 -f -o /tmp/${tag}_c4 \
    python bench.py --config 4 --steps 1 --warmup 0 --loci 1000 --no-cpu-baseline > $out/${tag}_c4_bench.log 2>&1
ncu -i /tmp/${tag}_c4.ncu-rep --page raw --csv > $out/${tag}_c4_raw.csv 2>/dev/null
ncu -i /tmp/${tag}_c4.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $out/${tag}_c4_sass.csv.gz
LTR_BAND=-1 timeout 600 python bench.py --config 4 --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_c4_noband.json 2> $out/${tag}_c4_noband.err
( time timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_viterbi.py tests/test_gpu_posteriors.py -x -q -m gpu ) > $out/${tag}_memcheck_probe.log 2>&1
echo "memcheck rc=$?" >> $out/${tag}_memcheck_probe.log
ls -la $out | tail -20
