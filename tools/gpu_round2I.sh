#!/usr/bin/env bash
# GPU visit r2I: full GPU suite, the default bench line with every sub-line (c4, c5, n2, n3, em), reference arm, launch list,
# one full ncu capture of the dominant kernel and of the EM kernel, sanitizers over the EM kernel tests.
out=gpurun_out; tag=r2I
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
timeout 1200 python bench.py > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.err; tail -c 600 $out/${tag}_bench_c3.json; echo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref_c3.json 2> $out/${tag}_bench_ref_c3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --loci 20000 --no-cpu-baseline --no-raw --no-extra > $out/${tag}_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"viterbi_band_kernel<4, 4" -c 1 -o $out/${tag}_full -f python bench.py --steps 1 --warmup 0 --loci 20000 --no-cpu-baseline --no-raw --no-extra > $out/${tag}_full_bench.log 2>&1
ncu -i $out/${tag}_full.ncu-rep --page raw --csv > $out/${tag}_full_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"em_stutter_kernel" -c 1 -o $out/${tag}_em_full -f python bench.py --em --steps 1 --warmup 0 --no-cpu-baseline > $out/${tag}_em_ncu.log 2>&1
ncu -i $out/${tag}_em_full.ncu-rep --page raw --csv > $out/${tag}_em_raw.csv 2>/dev/null
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_em.py -x -q > $out/${tag}_memcheck_em.log 2>&1; echo "memcheck rc=$?" >> $out/${tag}_memcheck_em.log; tail -4 $out/${tag}_memcheck_em.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_em.py -x -q -k "recorded" > $out/${tag}_racecheck_em.log 2>&1; echo "racecheck rc=$?" >> $out/${tag}_racecheck_em.log; tail -4 $out/${tag}_racecheck_em.log
ls -la $out | grep $tag
