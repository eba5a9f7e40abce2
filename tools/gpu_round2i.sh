#!/usr/bin/env bash
tag="${1:-r2i}"
out=gpurun_out
mkdir -p $out
python -m pytest tests/test_gpu_stutter.py -x -q -m gpu > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -4 $out/${tag}_pytest.log
python bench.py --config 5 --steps 2 --warmup 2 --no-cpu-baseline > $out/${tag}_bench_c5.json 2> $out/${tag}_bench_c5.err
python -c "
import json
d=json.load(open('$out/${tag}_bench_c5.json'))
print('c5', d['value'], d['ms_per_step'], d['roofline']['frac'])
"
