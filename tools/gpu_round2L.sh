#!/usr/bin/env bash
# GPU visit r2L: the shipped HG002 / trio loci through the library's own pipeline against the reference's VCF records
out=gpurun_out; tag=r2L
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_real_data.py -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log; tail -30 $out/${tag}_pytest.log
