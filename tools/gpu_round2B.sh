#!/usr/bin/env bash
# GPU visit r2B: kernel-2 diagnostics (resident CTAs, alleles per task).
out=gpurun_out; tag=r2B
mkdir -p $out
run() { python bench.py --config 5 --loci 20000 --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'value %.0f frac %.3f ms %.1f' % (d['value'], d['roofline']['frac'], d['ms_per_step']))" | tee -a $out/${tag}_k2.txt; }
run default
LTR_STUT_BLOCKS_PER_SM=4 run bps4
LTR_STUT_BLOCKS_PER_SM=5 run bps5
LTR_STUT_BLOCKS_PER_SM=8 run bps8
LTR_STUT_MAX_ALLELES=1 run alleles1
LTR_STUT_MAX_ALLELES=1 LTR_STUT_BLOCKS_PER_SM=8 run alleles1_bps8
LTR_STUT_MAX_ALLELES=2 run alleles2
