#!/usr/bin/env bash
# bench.py under torchrun on N GPUs of one box (what the driver's SCALE step launches), e2e + raw arms included.
n="${1:-2}"; tag="${2:-r2}"
out=gpurun_out
mkdir -p $out
start=$(date +%s)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $n --steps 5 --warmup 3 > $out/${tag}_scale${n}.json 2> $out/${tag}_scale${n}.err
echo "wall $(( $(date +%s) - start )) s" >> $out/${tag}_scale${n}.err
grep -E '^\{' $out/${tag}_scale${n}.json | python -c "
import json,sys
for line in sys.stdin:
    d=json.loads(line)
    print('N=%d value %.0f e2e %.0f raw %s ms %.2f' % (d['n_gpus'], d['value'], d['e2e']['value'], (d.get('e2e_from_flat_loci') or {}).get('value'), d['ms_per_step']))
"
tail -3 $out/${tag}_scale${n}.err
