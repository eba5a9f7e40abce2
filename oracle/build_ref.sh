#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.
# Builds oracle/_ref/libltr_ref.so: the reference's own hot-path translation units,
# compiled IN PLACE from /root/reference/src (no source is copied into this repo)
# with the reference Makefile's flags (Makefile:8: -O3 -g -std=c++0x -DMACOSX ...;
# never -O0), against the compile-only htslib shim in
# oracle/shim/, plus oracle/ref_driver.cpp (ours).  Outputs only into oracle/_ref/.
# If /root/reference is absent (GPU box) the prebuilt library is used as is.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ref="${LONGTR_REFERENCE:-/root/reference}"
out="$here/_ref"
mkdir -p "$out/obj"
if [ ! -d "$ref/src" ]; then
  echo "build_ref.sh: $ref/src not present; keeping prebuilt $out/libltr_ref.so" >&2
  exit 0
fi
CXX="${CXX:-g++}"
FLAGS="-O3 -g -std=c++0x -DMACOSX -D__STDC_LIMIT_MACROS -D_FILE_OFFSET_BITS=64 -fPIC -w"
TUS="SeqAlignment/HapAligner SeqAlignment/Haplotype SeqAlignment/HapBlock
     SeqAlignment/NeedlemanWunsch SeqAlignment/StutterAlignerClass
     SeqAlignment/AlignmentTraceback mathops stutter_model base_quality error
     stringops region read_pooler genotyper fasta_reader"
objs=""
for tu in $TUS; do
  o="$out/obj/$(basename "$tu").o"
  if [ ! -f "$o" ] || [ "$ref/src/$tu.cpp" -nt "$o" ]; then
    $CXX $FLAGS -I"$here/shim" -c "$ref/src/$tu.cpp" -o "$o"
  fi
  objs="$objs $o"
done
$CXX -O2 -g -std=c++11 -fPIC -w -I"$here/shim" -I"$ref/src" -I"$here/../include" \
     -c "$here/ref_driver.cpp" -o "$out/obj/ref_driver.o"
$CXX -O2 -std=c++11 -fPIC -w -I"$here/shim" -c "$here/shim/hts_stubs.cpp" -o "$out/obj/hts_stubs.o"
$CXX -shared -o "$out/libltr_ref.so" "$out/obj/ref_driver.o" "$out/obj/hts_stubs.o" $objs -Wl,--no-undefined -lm -lpthread
echo "built $out/libltr_ref.so"
