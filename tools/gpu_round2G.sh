#!/usr/bin/env bash
# GPU visit r2G (2 GPUs): where the host time of the one-process two-device run goes (submit timing)
out=gpurun_out; tag=r2G
mkdir -p $out
timeout 600 python bench.py --one-process-devices 2 --config 4 --steps 3 > $out/${tag}_oneproc_c4.json 2> $out/${tag}_oneproc_c4.err; cat $out/${tag}_oneproc_c4.json | head -c 2500; echo
LTR_TIMING=1 timeout 600 python bench.py --one-process-devices 2 --config 4 --steps 1 --loci 4000 > $out/${tag}_timing.json 2> $out/${tag}_timing.err; tail -40 $out/${tag}_timing.err
