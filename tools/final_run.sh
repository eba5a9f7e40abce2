# Round-end GPU evidence (one gpurun call): the GPU suite, the default bench line, extra.n3 alone, memcheck over the suite
# without the full-size cases (100 000 loci under memcheck take many minutes).
set -x
mkdir -p gpurun_out
T=${LTR_RUN_TAG:-r2X}
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -3 gpurun_out/${T}_pytest.log
python bench.py > gpurun_out/${T}_bench_c3.json 2> gpurun_out/${T}_bench_c3.err; echo "bench rc=$?"; tail -c 400 gpurun_out/${T}_bench_c3.json
python bench.py --n3 > gpurun_out/${T}_bench_n3.json 2> gpurun_out/${T}_bench_n3.err; echo "n3 rc=$?"; head -c 700 gpurun_out/${T}_bench_n3.json
(time timeout 130 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -x -q -k "not full_size") > gpurun_out/${T}_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/${T}_memcheck.log; tail -6 gpurun_out/${T}_memcheck.log
