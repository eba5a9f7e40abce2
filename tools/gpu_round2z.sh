#!/usr/bin/env bash
# GPU visit r2z: full ncu captures (band W=32 class on config 3; K=12 stream kernel and a whole-warp band class on config 4;
# Myers kernel), N2 bench line with its roofline.
out=gpurun_out; tag=r2z
mkdir -p $out
cap() {  # name, regex, bench args...
  local name=$1 re=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$re" -c 1 -o $out/${tag}_${name} -f python bench.py "$@" > $out/${tag}_${name}_ncu.log 2>&1
  ncu -i $out/${tag}_${name}.ncu-rep --page raw --csv > $out/${tag}_${name}_raw.csv 2>/dev/null
  ls -la $out/${tag}_${name}.ncu-rep 2>/dev/null
}
cap band44 'band_kernel<.int.4, .int.4,' --steps 1 --warmup 0 --loci 20000 --no-cpu-baseline --no-raw --no-extra
cap stream12 'stream_kernel<.int.12, .int.2' --config 4 --loci 2000 --steps 1 --warmup 0 --no-cpu-baseline --no-raw --no-extra
cap band432 'band_kernel<.int.4, .int.32,' --config 4 --loci 2000 --steps 1 --warmup 0 --no-cpu-baseline --no-raw --no-extra
cap myers 'edit_myers_kernel' --n2 --loci 256 --steps 1 --warmup 0 --no-cpu-baseline
timeout 300 python bench.py --n2 --steps 3 --warmup 3 > $out/${tag}_bench_n2.json 2> $out/${tag}_bench_n2.err; tail -c 900 $out/${tag}_bench_n2.json
