#!/usr/bin/env bash
# Round-2 visit after the device plan + asynchronous jobs: parity suite, default bench line, sanitizer over the new paths.
tag="${1:-r2b}"
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -5 $out/${tag}_pytest.log
LTR_TIMING=1 timeout 900 python bench.py --steps 5 --warmup 3 > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.err
cat $out/${tag}_bench_c3.json | cut -c1-3000
( time timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_plan_async.py tests/test_gpu_viterbi.py tests/test_gpu_band.py -x -q -m gpu ) > $out/${tag}_memcheck.log 2>&1
echo "memcheck rc=$?" >> $out/${tag}_memcheck.log
tail -8 $out/${tag}_memcheck.log
