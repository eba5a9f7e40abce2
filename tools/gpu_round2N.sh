#!/usr/bin/env bash
# GPU visit r2N (N GPUs): the driver's SCALE launch with 20 steps (end-to-end arm with both read encodings)
n="${1:-2}"; out=gpurun_out; tag=r2N
mkdir -p $out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $n --steps 20 --warmup 3 --no-raw > $out/${tag}_scale${n}.json 2> $out/${tag}_scale${n}.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2N_scale$n.json') if l.startswith('{')][-1])
e=d["e2e"]
print("N", d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(e["value"]), "4-bit", e.get("value_4bit_reads"))
PY
