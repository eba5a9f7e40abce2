#!/usr/bin/env bash
# GPU visit r2K (8 GPUs): the driver's SCALE launch at N = 8 with 20 steps, end-to-end arm with both read encodings
out=gpurun_out; tag=r2K; n=8
mkdir -p $out
start=$(date +%s)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $n --steps 20 --warmup 3 --no-raw > $out/${tag}_scale${n}.json 2> $out/${tag}_scale${n}.err
echo "wall $(( $(date +%s) - start )) s" >> $out/${tag}_scale${n}.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2K_scale8.json') if l.startswith('{')][-1])
print("N", d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"],2)); print(json.dumps(d["e2e"])[:1600])
PY
tail -3 $out/${tag}_scale${n}.err
