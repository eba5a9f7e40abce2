#!/usr/bin/env bash
# GPU visit r2C (2 GPUs): posterior kernel far beyond the synthetic sizes, N=2 bench with ranks bound to their GPU's CPUs.
out=gpurun_out; tag=r2C
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_posteriors.py tests/test_gpu_stutter.py -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log; tail -3 $out/${tag}_pytest.log
bash tools/gpu_scale.sh 2 r2C
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2C_scale2.json') if l.startswith('{')][-1]); print('bound cpus', d['config'].get('rank_bound_to_cpus_of_its_gpu'))"
nproc; lscpu | grep -i "numa\|socket\|model name" | head -8
