# Host-side probe of the region loop on the GPU box: threads sweep of the preparation alone + extra.n3 with its host timing.
set -x
nproc; lscpu | grep -E "Model name|Thread|Core|Socket|MHz" | head -8
g++ -O2 -std=c++17 -Iinclude tools/region_prepare_bench.cpp -Llongtr_b200/csrc -llongtr_b200 -Wl,-rpath,$PWD/longtr_b200/csrc -pthread -o /tmp/region_prepare_bench
python - <<'P'
import sys
sys.path.insert(0, "tests")
import bam_writer as bw
w = bw.synthetic_world(1500, config=3, first_locus=0, n_samples=1)
bw.write_world(w, "/tmp")
open("/tmp/regions.txt", "w").write("".join("%d %d %d\n" % r for r in w["regions"]))
open("/tmp/chrom.txt", "w").write(w["chrom_seq"])
P
for t in 1 4 8 16 32; do /tmp/region_prepare_bench /tmp/synth_0.bam /tmp/regions.txt /tmp/chrom.txt $t | tail -1; done
python bench.py --n3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({k:d[k] for k in ('value','ms_per_step','host_ms','host_ms_with_vcf_records','genotyper_ms','host_threads_per_gpu')}))"
