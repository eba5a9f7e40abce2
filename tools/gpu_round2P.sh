#!/usr/bin/env bash
# GPU visit r2P: compute-sanitizer over the code paths added late in round 2 (4-bit read stream / unpack kernel, LL download and
# per-read allele assignment, VCF records through the region loop, EM kernel, multi-device chunking on one device)
out=gpurun_out; tag=r2P
mkdir -p $out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_plan_async.py -x -q -k "4bit or packed or submit_wait" > $out/${tag}_memcheck_4bit.log 2>&1; echo "memcheck rc=$?" >> $out/${tag}_memcheck_4bit.log; tail -4 $out/${tag}_memcheck_4bit.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_real_data.py tests/test_gpu_regions.py tests/test_gpu_genotyper.py -x -q > $out/${tag}_memcheck_regions.log 2>&1; echo "memcheck rc=$?" >> $out/${tag}_memcheck_regions.log; tail -4 $out/${tag}_memcheck_regions.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_plan_async.py -x -q -k "4bit" > $out/${tag}_racecheck_4bit.log 2>&1; echo "racecheck rc=$?" >> $out/${tag}_racecheck_4bit.log; tail -4 $out/${tag}_racecheck_4bit.log
