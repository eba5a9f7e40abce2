#!/usr/bin/env bash
tag="${1:-r2c}"
out=gpurun_out
mkdir -p $out
python -m pytest tests/test_gpu_plan_async.py tests/test_gpu_viterbi.py tests/test_gpu_band.py -x -q -m gpu > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
LTR_TIMING=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-extra > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.err
cat $out/${tag}_bench_c3.json | cut -c1-2500
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > $out/${tag}_launches_bench.log 2>&1
