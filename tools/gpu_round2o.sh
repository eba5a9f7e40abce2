#!/usr/bin/env bash
tag="${1:-r2o}"
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_fullsize.py > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -4 $out/${tag}_pytest.log
python tools/latency_probe.py > $out/${tag}_latency.txt 2>&1; cat $out/${tag}_latency.txt
python tools/dropin_timing.py 1160 > $out/${tag}_dropin_timing.txt 2>&1; cat $out/${tag}_dropin_timing.txt
python bench.py --config 5 --steps 2 --warmup 2 --no-cpu-baseline > $out/${tag}_bench_c5.json 2> $out/${tag}_bench_c5.err
python -c "
import json
d=json.load(open('$out/${tag}_bench_c5.json'))
print('c5', d['value'], d['ms_per_step'], d['roofline']['frac'])
"
( time timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_stutter.py -x -q -m gpu ) > $out/${tag}_racecheck_stutter.log 2>&1
echo "racecheck rc=$?" >> $out/${tag}_racecheck_stutter.log; tail -5 $out/${tag}_racecheck_stutter.log
