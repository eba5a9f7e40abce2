#!/usr/bin/env bash
# Round-2 full visit to a B200 box: parity suite, sanitizer runs, default bench line (with extra.c4 / extra.c5 and the
# raw-loci arm), reference arm, drop-in timing, ncu launch list + one full capture of the dominant kernel.
# Usage (from the repo root, under gpurun): bash tools/gpu_round2.sh <tag>
tag="${1:-r2}"
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -4 $out/${tag}_pytest.log
python tools/first_gpu_probe.py > $out/${tag}_probe.log 2>&1
python bench.py --steps 5 --warmup 3 > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.err
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_ref_c3.json 2> $out/${tag}_bench_ref_c3.err
cut -c1-1200 $out/${tag}_bench_c3.json; cat $out/${tag}_bench_ref_c3.json | cut -c1-400
( time timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q \
    --deselect tests/test_gpu_fullsize.py --deselect tests/test_gpu_dropin.py ) > $out/${tag}_memcheck.log 2>&1
echo "memcheck rc=$?" >> $out/${tag}_memcheck.log; tail -6 $out/${tag}_memcheck.log
( time timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_viterbi.py \
    tests/test_gpu_band.py tests/test_gpu_plan_async.py tests/test_gpu_posteriors.py tests/test_gpu_stutter.py \
    tests/test_gpu_genotyper.py -x -q -m gpu ) > $out/${tag}_racecheck.log 2>&1
echo "racecheck rc=$?" >> $out/${tag}_racecheck.log; tail -6 $out/${tag}_racecheck.log
python tools/dropin_timing.py 290 > $out/${tag}_dropin_timing.txt 2>&1; cat $out/${tag}_dropin_timing.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --loci 20000 --no-cpu-baseline --no-extra --no-raw > $out/${tag}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:viterbi_band_kernel -c 3 -f -o $out/${tag}_full \
    python bench.py --steps 1 --warmup 1 --loci 20000 --no-cpu-baseline --no-extra --no-raw > $out/${tag}_full_bench.log 2>&1
ncu -i $out/${tag}_full.ncu-rep --page raw --csv > $out/${tag}_full_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:stutter_pair_kernel -c 1 -f -o /tmp/${tag}_stutter \
    python bench.py --config 5 --steps 1 --warmup 1 --loci 4000 --no-cpu-baseline > $out/${tag}_stutter_bench.log 2>&1
ncu -i /tmp/${tag}_stutter.ncu-rep --page raw --csv > $out/${tag}_stutter_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:viterbi_stream_kernel.12 -c 2 -f -o /tmp/${tag}_c4stream \
    python bench.py --config 4 --steps 1 --warmup 0 --no-cpu-baseline --no-raw > $out/${tag}_c4stream_bench.log 2>&1
ncu -i /tmp/${tag}_c4stream.ncu-rep --page raw --csv > $out/${tag}_c4stream_raw.csv 2>/dev/null
ls -la $out | tail -5
