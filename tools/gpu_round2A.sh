#!/usr/bin/env bash
# GPU visit r2A: kernel 2 with persistent warps and phase A shared by the alleles of a read (parity, config-5 bench, sanitizers),
# extra.n3 (BAM files -> calls).
out=gpurun_out; tag=r2A
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_stutter.py tests/test_gpu_fullsize.py tests/test_gpu_regions.py -x -q -k "stutter or homopolymer or regions" > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
timeout 600 python bench.py --config 5 --steps 2 --warmup 3 > $out/${tag}_bench_c5.json 2> $out/${tag}_bench_c5.err; python -c "
import json; d=json.loads(open('$out/${tag}_bench_c5.json').read().strip().splitlines()[-1]); print('c5 value %.0f e2e %.0f frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['frac']))"
timeout 600 python bench.py --n3 --steps 2 --warmup 1 > $out/${tag}_bench_n3.json 2> $out/${tag}_bench_n3.err; tail -c 1200 $out/${tag}_bench_n3.json; tail -3 $out/${tag}_bench_n3.err
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_stutter.py -x -q > $out/${tag}_racecheck_stutter.log 2>&1; tail -4 $out/${tag}_racecheck_stutter.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_stutter.py tests/test_gpu_edit.py tests/test_gpu_regions.py -x -q > $out/${tag}_memcheck.log 2>&1; tail -4 $out/${tag}_memcheck.log
