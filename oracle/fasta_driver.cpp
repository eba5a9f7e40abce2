// TEST INFRASTRUCTURE ONLY -- the reference's FastaReader (src/fasta_reader.{h,cpp}, compiled IN PLACE) on top of
// integration/faidx_compat.cpp (the faidx names it binds, served by the library's own FASTA reader), and
// Genotyper::get_vcf_header (src/genotyper.cpp:258-336), whose contig lines come through the same reader.
#include <stdlib.h>
#include <string.h>

#include <sstream>
#include <string>
#include <vector>

#include "fasta_reader.h"
#include "genotyper.h"

static char* dup(const std::string& s) {
  char* r = (char*)malloc(s.size() + 1);
  memcpy(r, s.c_str(), s.size() + 1);
  return r;
}

// FastaReader(path).get_sequence(chrom) / get_sequence_length
extern "C" char* ltr_ref_fasta_sequence(const char* path, const char* chrom, long long* length) {
  FastaReader reader(path);
  *length = reader.get_sequence_length(chrom);
  std::string seq;
  if (*length >= 0) reader.get_sequence(chrom, seq);
  return dup(seq);
}

// samples: newline separated
extern "C" char* ltr_ref_vcf_header(const char* fasta_path, const char* command, const char* samples) {
  std::vector<std::string> names, chroms;
  std::stringstream ss(samples);
  std::string item;
  while (std::getline(ss, item, '\n'))
    if (!item.empty()) names.push_back(item);
  return dup(Genotyper::get_vcf_header(fasta_path, command, chroms, names));
}

// The same under the output switches `mask` (the library's LTR_VCF_* bits: 1 ALLREADS, 2 MALLREADS, 4 GL, 8 PL, 16 PHASEDGL,
// 32 FILTER -> Genotyper::OUTPUT_*, what src/hipstr_main.cpp:178-183 sets); the defaults are restored afterwards.
extern "C" char* ltr_ref_vcf_header_switches(const char* fasta_path, const char* command, const char* samples, unsigned mask) {
  int* sw[6] = {&Genotyper::OUTPUT_ALLREADS, &Genotyper::OUTPUT_MALLREADS,  &Genotyper::OUTPUT_GLS,
                &Genotyper::OUTPUT_PLS,      &Genotyper::OUTPUT_PHASED_GLS, &Genotyper::OUTPUT_FILTERS};
  int saved[6];
  for (int i = 0; i < 6; ++i) {
    saved[i] = *sw[i];
    *sw[i] = (mask >> i) & 1;
  }
  char* text = ltr_ref_vcf_header(fasta_path, command, samples);
  for (int i = 0; i < 6; ++i) *sw[i] = saved[i];
  return text;
}
