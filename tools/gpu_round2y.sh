#!/usr/bin/env bash
# GPU visit r2y: full GPU suite (regions pipeline included), the default bench line with every sub-line, reference arm,
# launch list + full ncu captures (band W=32 class, K=12 stream kernel, whole-warp band class).
out=gpurun_out; tag=r2y
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
timeout 900 python bench.py > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.err; tail -c 400 $out/${tag}_bench_c3.json | head -c 400; echo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref_c3.json 2> $out/${tag}_bench_ref_c3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --loci 20000 --no-cpu-baseline --no-raw --no-extra > $out/${tag}_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"viterbi_band_kernel<4, 4" -c 1 -o $out/${tag}_full -f python bench.py --steps 1 --warmup 0 --loci 20000 --no-cpu-baseline --no-raw --no-extra > $out/${tag}_full_bench.log 2>&1
ncu -i $out/${tag}_full.ncu-rep --page raw --csv > $out/${tag}_full_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"viterbi_stream_kernel<12" -c 1 -o $out/${tag}_stream12_full -f python bench.py --config 4 --loci 2000 --steps 1 --warmup 0 --no-cpu-baseline --no-raw --no-extra > $out/${tag}_stream12_ncu.log 2>&1
ncu -i $out/${tag}_stream12_full.ncu-rep --page raw --csv > $out/${tag}_stream12_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"viterbi_band_kernel<4, 32" -c 1 -o $out/${tag}_band432_full -f python bench.py --config 4 --loci 2000 --steps 1 --warmup 0 --no-cpu-baseline --no-raw --no-extra > $out/${tag}_band432_ncu.log 2>&1
ncu -i $out/${tag}_band432_full.ncu-rep --page raw --csv > $out/${tag}_band432_raw.csv 2>/dev/null
ls -la $out | grep $tag
