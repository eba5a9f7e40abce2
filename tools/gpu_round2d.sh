#!/usr/bin/env bash
tag="${1:-r2d}"
out=gpurun_out
mkdir -p $out
python -m pytest tests -x -q -m gpu > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -6 $out/${tag}_pytest.log
LTR_TIMING=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-extra > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.err
cat $out/${tag}_bench_c3.json | cut -c1-2200
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > $out/${tag}_launches_bench.log 2>&1
