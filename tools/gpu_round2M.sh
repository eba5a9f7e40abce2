#!/usr/bin/env bash
# GPU visit r2M: VCF records through ltr_regions_run (synthetic worlds against the reference's records) + real-data records
out=gpurun_out; tag=r2M
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_regions.py tests/test_gpu_real_data.py tests/test_gpu_genotyper.py -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log; tail -40 $out/${tag}_pytest.log
