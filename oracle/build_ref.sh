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
# candidate-haplotype clustering (N2): the reference's HaplotypeGenerator behind oracle/edit_driver.cpp
o="$out/obj/HaplotypeGenerator.o"
if [ ! -f "$o" ] || [ "$ref/src/SeqAlignment/HaplotypeGenerator.cpp" -nt "$o" ]; then
  $CXX $FLAGS -I"$here/shim" -c "$ref/src/SeqAlignment/HaplotypeGenerator.cpp" -o "$o"
fi
$CXX -O2 -g -std=c++11 -fPIC -w -fno-access-control -I"$here/shim" -I"$ref/src" \
     -c "$here/edit_driver.cpp" -o "$out/obj/edit_driver.o"
$CXX -shared -o "$out/libltr_ref.so" "$out/obj/ref_driver.o" "$out/obj/edit_driver.o" "$out/obj/HaplotypeGenerator.o" \
     "$out/obj/hts_stubs.o" $objs -Wl,--no-undefined -lm -lpthread
echo "built $out/libltr_ref.so"

# ---- libltr_ref_io.so: the reference's alignment input (bam_io.cpp, compiled in place) on top of the library's BAM reader
#      through integration/hts_compat.cpp (the htslib names LongTR binds, implemented on ltr_bam_*) ----------------------
#      (bam_reader.cpp is plain C++: it is compiled into this library, which therefore does not pull liblongtr_b200.so and a
#      second C++ runtime into the test process)
o="$out/obj/bam_io.o"
if [ ! -f "$o" ] || [ "$ref/src/bam_io.cpp" -nt "$o" ]; then
  $CXX $FLAGS -I"$here/shim" -c "$ref/src/bam_io.cpp" -o "$o"
fi
$CXX -O2 -g -std=c++11 -fPIC -w -I"$here/shim" -I"$here/../include" -c "$here/../integration/hts_compat.cpp" -o "$out/obj/hts_compat.o"
$CXX -O2 -g -std=c++11 -fPIC -w -I"$here/../include" -c "$here/../longtr_b200/csrc/host/bam_reader.cpp" -o "$out/obj/bam_reader.o"
$CXX -O2 -g -std=c++11 -fPIC -w -I"$here/shim" -I"$ref/src" -c "$here/io_driver.cpp" -o "$out/obj/io_driver.o"
# linked against the shared C++ runtime (a private static copy next to the process's own breaks iostream's locale facets)
LINKXX="$CXX"; [ -x /usr/bin/g++ ] && LINKXX=/usr/bin/g++
$LINKXX -shared -o "$out/libltr_ref_io.so" "$out/obj/io_driver.o" "$out/obj/bam_io.o" "$out/obj/hts_compat.o" "$out/obj/bam_reader.o" \
     "$out/obj/error.o" "$out/obj/stringops.o" -Wl,--no-undefined -lz -lm -lpthread
echo "built $out/libltr_ref_io.so"

# ---- libltr_ref_hapgen.so: the reference's HaplotypeGenerator (compiled with the spoa stand-in throwing) behind
#      oracle/hapgen_driver.cpp: candidate alleles of a region from flat reads ---------------------------------------------
$CXX $FLAGS -DLTR_SPOA_THROW -I"$here/shim" -c "$ref/src/SeqAlignment/HaplotypeGenerator.cpp" -o "$out/obj/HaplotypeGenerator_throw.o"
$CXX -O2 -g -std=c++11 -fPIC -w -fno-access-control -I"$here/shim" -I"$ref/src" -c "$here/hapgen_driver.cpp" -o "$out/obj/hapgen_driver.o"
$LINKXX -shared -o "$out/libltr_ref_hapgen.so" "$out/obj/hapgen_driver.o" "$out/obj/HaplotypeGenerator_throw.o" \
     "$out/obj/HapBlock.o" "$out/obj/error.o" "$out/obj/stringops.o" "$out/obj/stutter_model.o" "$out/obj/mathops.o" \
     "$out/obj/region.o" -Wl,--no-undefined -lm -lpthread
echo "built $out/libltr_ref_hapgen.so"
# the same with the spoa names served by oracle/poa_restatement.hpp: the reference's assembly branch runs to completion
$CXX $FLAGS -DLTR_SPOA_RESTATEMENT -I"$here/shim" -c "$ref/src/SeqAlignment/HaplotypeGenerator.cpp" -o "$out/obj/HaplotypeGenerator_poa.o"
$CXX -O2 -g -std=c++11 -fPIC -w -fno-access-control -DLTR_SPOA_RESTATEMENT -I"$here/shim" -I"$ref/src" -c "$here/hapgen_driver.cpp" -o "$out/obj/hapgen_driver_poa.o"
$LINKXX -shared -o "$out/libltr_ref_hapgen_poa.so" "$out/obj/hapgen_driver_poa.o" "$out/obj/HaplotypeGenerator_poa.o" \
     "$out/obj/HapBlock.o" "$out/obj/error.o" "$out/obj/stringops.o" "$out/obj/stutter_model.o" "$out/obj/mathops.o" \
     "$out/obj/region.o" -Wl,--no-undefined -lm -lpthread
echo "built $out/libltr_ref_hapgen_poa.so"

# ---- libltr_ref_fasta.so: the reference's FastaReader + Genotyper::get_vcf_header on integration/faidx_compat.cpp -> ltr_fasta_* ----
$CXX -O2 -g -std=c++11 -fPIC -w -I"$here/../include" -c "$here/../integration/faidx_compat.cpp" -o "$out/obj/faidx_compat.o"
$CXX -O2 -g -std=c++17 -fPIC -w -I"$here/../include" -c "$here/../longtr_b200/csrc/host/fasta_reader.cpp" -o "$out/obj/ltr_fasta_reader.o"
$CXX -O2 -g -std=c++11 -fPIC -w -I"$here/shim" -I"$ref/src" -c "$here/fasta_driver.cpp" -o "$out/obj/fasta_driver.o"
$LINKXX -shared -o "$out/libltr_ref_fasta.so" "$out/obj/fasta_driver.o" "$out/obj/fasta_reader.o" "$out/obj/genotyper.o" \
     "$out/obj/faidx_compat.o" "$out/obj/ltr_fasta_reader.o" "$out/obj/mathops.o" "$out/obj/error.o" "$out/obj/stringops.o" \
     -Wl,--no-undefined -lm -lpthread
echo "built $out/libltr_ref_fasta.so"

# ---- libltr_ref_em.so: the reference's EMStutterGenotyper (length-based EM of the stutter model) behind oracle/em_driver.cpp ----
$CXX $FLAGS -I"$here/shim" -c "$ref/src/em_stutter_genotyper.cpp" -o "$out/obj/em_stutter_genotyper.o"
$CXX -O2 -g -std=c++11 -fPIC -w -fno-access-control -I"$here/shim" -I"$ref/src" -c "$here/em_driver.cpp" -o "$out/obj/em_driver.o"
$LINKXX -shared -o "$out/libltr_ref_em.so" "$out/obj/em_driver.o" "$out/obj/em_stutter_genotyper.o" "$out/obj/genotyper.o" \
     "$out/obj/stutter_model.o" "$out/obj/mathops.o" "$out/obj/error.o" "$out/obj/stringops.o" "$out/obj/region.o" \
     "$out/obj/fasta_reader.o" "$out/obj/hts_stubs.o" \
     -Wl,--no-undefined -lm -lpthread
echo "built $out/libltr_ref_em.so"

# ---- IO-less per-locus genotyper (SeqStutterGenotyper ctor -> genotype -> write_vcf_record), twice ------------
#   ltr_ref_full : every object is the reference's own (golden VCF records)
#   ltr_ref_gpu  : HapAligner::process_reads and Genotyper::calc_log_sample_posteriors are taken from
#                        integration/reference_binding.cpp (C ABI -> liblongtr_b200.so); the reference's own
#                        definitions are weakened in COPIES of its objects under oracle/_ref/obj.  Haplotype::aln_haps_to_ref
#                        is elided as in ltr_ref_lazy (integration/lazy_haplotype_alignment.cpp).
FULL_TUS="seq_stutter_genotyper SeqAlignment/HaplotypeGenerator SeqAlignment/AlignmentOps vcf_writer vcf_input
          debruijn_graph directed_graph extract_indels zalgorithm"
fobjs=""
for tu in $FULL_TUS; do
  o="$out/obj/$(basename "$tu").o"
  # the generator of the full binaries is the build whose spoa names are served by oracle/poa_restatement.hpp (above)
  if [ "$tu" = "SeqAlignment/HaplotypeGenerator" ]; then fobjs="$fobjs $out/obj/HaplotypeGenerator_poa.o"; continue; fi
  if [ ! -f "$o" ] || [ "$ref/src/$tu.cpp" -nt "$o" ]; then
    $CXX $FLAGS -I"$here/shim" -c "$ref/src/$tu.cpp" -o "$o"
  fi
  fobjs="$fobjs $o"
done
$CXX -O2 -g -std=c++11 -fPIC -w -DLTR_FULL_MAIN -I"$here/shim" -I"$ref/src" -I"$here" -c "$here/full_driver.cpp" -o "$out/obj/full_driver.o"
base_objs=""
for o in $objs; do case "$o" in *fasta_reader.o) ;; *) base_objs="$base_objs $o";; esac; done
$CXX -o "$out/ltr_ref_full" "$out/obj/full_driver.o" "$out/obj/hts_stubs.o" "$out/obj/fasta_reader.o" \
     $base_objs $fobjs -lm -lpthread
echo "built $out/ltr_ref_full"
# ---- ltr_ref_trace: all-CPU reference with a recording wrapper around Genotyper::calc_log_sample_posteriors ---------
objcopy --redefine-sym _ZN9Genotyper26calc_log_sample_posteriorsERSt6vectorIiSaIiEE=ltr_orig_calc_log_sample_posteriors \
        "$out/obj/genotyper.o" "$out/obj/genotyper_trace.o"
$CXX -O2 -g -std=c++11 -fPIC -w -DLTR_FULL_MAIN -DLTR_TRACE_POSTERIORS -I"$here/shim" -I"$ref/src" -I"$here" \
     -c "$here/full_driver.cpp" -o "$out/obj/full_driver_trace.o"
trace_objs=""
for o in $base_objs; do
  case "$o" in
    *genotyper.o) trace_objs="$trace_objs $out/obj/genotyper_trace.o";;
    *) trace_objs="$trace_objs $o";;
  esac
done
$CXX -o "$out/ltr_ref_trace" "$out/obj/full_driver_trace.o" "$out/obj/hts_stubs.o" "$out/obj/fasta_reader.o" \
     $trace_objs $fobjs -lm -lpthread
echo "built $out/ltr_ref_trace"
# ---- ltr_ref_lazy: all-CPU reference with Haplotype::aln_haps_to_ref elided (integration/lazy_haplotype_alignment.cpp);
#      the reference's own definition is renamed in a copy of its object and stays reachable through an env switch.
objcopy --redefine-sym _ZN9Haplotype15aln_haps_to_refEv=ltr_orig_aln_haps_to_ref \
        "$out/obj/Haplotype.o" "$out/obj/Haplotype_lazy.o"
$CXX -O2 -g -std=c++11 -fPIC -w -I"$here/shim" -I"$ref/src" \
     -c "$here/../integration/lazy_haplotype_alignment.cpp" -o "$out/obj/lazy_haplotype_alignment.o"
lazy_objs=""
for o in $base_objs; do
  case "$o" in
    *obj/Haplotype.o) lazy_objs="$lazy_objs $out/obj/Haplotype_lazy.o";;
    *) lazy_objs="$lazy_objs $o";;
  esac
done
$CXX -o "$out/ltr_ref_lazy" "$out/obj/full_driver.o" "$out/obj/lazy_haplotype_alignment.o" "$out/obj/hts_stubs.o" \
     "$out/obj/fasta_reader.o" $lazy_objs $fobjs -lm -lpthread
echo "built $out/ltr_ref_lazy"
lib="$here/../longtr_b200/csrc"
if [ -f "$lib/liblongtr_b200.so" ]; then
  objcopy --weaken-symbol=_ZN10HapAligner13process_readsERKSt6vectorI9AlignmentSaIS1_EEiPK11BaseQualityRKS0_IbSaIbEEPdPi \
          "$out/obj/HapAligner.o" "$out/obj/HapAligner_weak.o"
  objcopy --weaken-symbol=_ZN9Genotyper26calc_log_sample_posteriorsERSt6vectorIiSaIiEE \
          "$out/obj/genotyper.o" "$out/obj/genotyper_weak.o"
  $CXX -O2 -g -std=c++11 -fPIC -w -I"$here/shim" -I"$ref/src" -I"$here/../include" \
       -c "$here/../integration/reference_binding.cpp" -o "$out/obj/reference_binding.o"
  gpu_objs=""
  for o in $base_objs; do
    case "$o" in
      *HapAligner.o) gpu_objs="$gpu_objs $out/obj/HapAligner_weak.o";;
      *obj/Haplotype.o) gpu_objs="$gpu_objs $out/obj/Haplotype_lazy.o";;
      *genotyper.o) gpu_objs="$gpu_objs $out/obj/genotyper_weak.o";;
      *) gpu_objs="$gpu_objs $o";;
    esac
  done
  $CXX -o "$out/ltr_ref_gpu" "$out/obj/reference_binding.o" "$out/obj/lazy_haplotype_alignment.o" "$out/obj/full_driver.o" "$out/obj/hts_stubs.o" \
       "$out/obj/fasta_reader.o" $gpu_objs $fobjs -L"$lib" -llongtr_b200 \
       -Wl,-rpath,'$ORIGIN/../../longtr_b200/csrc' -lm -lpthread
  echo "built $out/ltr_ref_gpu"
fi
