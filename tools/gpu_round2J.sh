#!/usr/bin/env bash
# GPU visit r2J: reads as a 4-bit stream (ltr_ctx_set_read_encoding) -- parity tests and the end-to-end arm with both encodings
out=gpurun_out; tag=r2J
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_plan_async.py -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log; tail -12 $out/${tag}_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-extra --no-raw --no-cpu-baseline > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2J_bench_c3.json') if l.startswith('{')][-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"],2)); print(json.dumps(d["e2e"], indent=0)[:1500])
PY
tail -3 $out/${tag}_bench_c3.err
