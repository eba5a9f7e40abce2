#!/usr/bin/env bash
# GPU visit r2D: regions pipeline with the assembly branch (tests + extra.n3 line)
out=gpurun_out; tag=r2D
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_regions.py tests/test_gpu_genotyper.py -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -5 $out/${tag}_pytest.log
timeout 600 python bench.py --n3 > $out/${tag}_bench_n3.json 2> $out/${tag}_bench_n3.err; cat $out/${tag}_bench_n3.json | head -c 1500; echo
