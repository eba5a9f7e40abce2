set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2U_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2U_pytest.log; tail -3 gpurun_out/r2U_pytest.log
python bench.py > gpurun_out/r2U_bench_c3.json 2> gpurun_out/r2U_bench_c3.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2U_bench_c3.json
python bench.py --impl reference > gpurun_out/r2U_bench_ref_c3.json 2> gpurun_out/r2U_bench_ref.err; echo "ref rc=$?"
python bench.py --n3 > gpurun_out/r2U_bench_n3.json 2> gpurun_out/r2U_bench_n3.err; echo "n3 rc=$?"; tail -c 1500 gpurun_out/r2U_bench_n3.json
(time timeout 300 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -x -q) > gpurun_out/r2U_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2U_memcheck.log; tail -4 gpurun_out/r2U_memcheck.log
for f in viterbi band stutter posteriors plan_async edit em genotyper regions real_data; do
  (time timeout 240 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_$f.py -x -q) > gpurun_out/r2U_racecheck_$f.log 2>&1; echo "racecheck $f rc=$?" | tee -a gpurun_out/r2U_racecheck_$f.log; grep -c "Race reported" gpurun_out/r2U_racecheck_$f.log
done
