#!/usr/bin/env bash
# GPU visit r2O (final state of round 2): full GPU suite, the default bench line with every sub-line, reference arm, launch
# list, full ncu capture of the dominant kernel (W = 32 band class).
out=gpurun_out; tag=r2O
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
timeout 1200 python bench.py > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.err; tail -c 300 $out/${tag}_bench_c3.json; echo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref_c3.json 2> $out/${tag}_bench_ref_c3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --loci 20000 --no-cpu-baseline --no-raw --no-extra > $out/${tag}_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'band_kernel<.int.4, .int.4,' -c 1 -o $out/${tag}_full -f python bench.py --steps 1 --warmup 0 --loci 20000 --no-cpu-baseline --no-raw --no-extra > $out/${tag}_full_bench.log 2>&1
ncu -i $out/${tag}_full.ncu-rep --page raw --csv > $out/${tag}_full_raw.csv 2>/dev/null
ls -la $out | grep $tag
