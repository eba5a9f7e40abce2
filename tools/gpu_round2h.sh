#!/usr/bin/env bash
tag="${1:-r2h}"
out=gpurun_out
mkdir -p $out
python -m pytest tests/test_gpu_genotyper.py tests/test_gpu_posteriors.py -x -q -m gpu > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -25 $out/${tag}_pytest.log
LTR_TIMING=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.err
python -c "
import json
d=json.load(open('$out/${tag}_bench_c3.json'))
print(d['value'], d['e2e']['value'], json.dumps(d['e2e_from_flat_loci'], indent=1))
"
tail -5 $out/${tag}_bench_c3.err
