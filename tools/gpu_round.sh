#!/usr/bin/env bash
# One GPU-box visit: parity tests, clean bench lines, ncu launch list and one full capture.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag>
tag="${1:-r1}"
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
python tools/first_gpu_probe.py > $out/${tag}_probe.log 2>&1
python bench.py --steps 5 --warmup 3 > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.err
python bench.py --config 4 --steps 3 --warmup 3 > $out/${tag}_bench_c4.json 2> $out/${tag}_bench_c4.err
python bench.py --config 5 --steps 3 --warmup 3 > $out/${tag}_bench_c5.json 2> $out/${tag}_bench_c5.err
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_ref_c3.json 2> $out/${tag}_bench_ref_c3.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --loci 20000 --no-cpu-baseline > $out/${tag}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:viterbi_band_kernel -c 3 -f -o $out/${tag}_full \
    python bench.py --steps 1 --warmup 1 --loci 20000 --no-cpu-baseline > $out/${tag}_full_bench.log 2>&1
tail -3 $out/${tag}_pytest.log; cat $out/${tag}_bench_c3.json $out/${tag}_bench_c4.json $out/${tag}_bench_c5.json $out/${tag}_bench_ref_c3.json
