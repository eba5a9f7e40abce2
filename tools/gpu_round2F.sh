#!/usr/bin/env bash
# GPU visit r2F (2 GPUs): one process driving two devices -- same calls, and what it buys on config 4 (GPU-bound) / 3 (host-bound)
out=gpurun_out; tag=r2F
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_genotyper.py -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log; tail -3 $out/${tag}_pytest.log
timeout 600 python bench.py --one-process-devices 2 --config 4 --steps 2 > $out/${tag}_oneproc_c4.json 2> $out/${tag}_oneproc_c4.err; cat $out/${tag}_oneproc_c4.json | head -c 1800; echo
timeout 600 python bench.py --one-process-devices 2 --config 3 --steps 2 > $out/${tag}_oneproc_c3.json 2> $out/${tag}_oneproc_c3.err; cat $out/${tag}_oneproc_c3.json | head -c 1800; echo
