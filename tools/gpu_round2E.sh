#!/usr/bin/env bash
# GPU visit r2E: ltr_run_bed, drop-in on the 52 real loci (5 with the assembly branch)
out=gpurun_out; tag=r2E
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_regions.py tests/test_gpu_dropin.py -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -15 $out/${tag}_pytest.log
