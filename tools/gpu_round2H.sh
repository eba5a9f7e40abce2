#!/usr/bin/env bash
# GPU visit r2H: stutter-model EM kernel against the reference
out=gpurun_out; tag=r2H
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_em.py -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log; tail -40 $out/${tag}_pytest.log
