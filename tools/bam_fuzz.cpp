// Robustness check of the BAM reader and the region preparation on damaged files (not part of the test suite: run by hand).
// tools/bam_fuzz_make.py writes 300 BAM files whose UNCOMPRESSED content was damaged (random bytes, extreme 32-bit values) and
// re-compressed into well-formed BGZF blocks; this program opens, indexes and queries every one of them and runs
// ltr_region_collect + ltr_candidate_alleles on top.  Build with -fsanitize=address,undefined:
//   g++ -O1 -g -std=c++17 -fsanitize=address,undefined -Iinclude -Ilongtr_b200/csrc tools/bam_fuzz.cpp \
//       longtr_b200/csrc/host/{bam_reader,region_loader,candidate_alleles,poa}.cpp -lz -pthread -o /tmp/bam_fuzz
//   python tools/bam_fuzz_make.py <seed> && /tmp/bam_fuzz
// Round 2: 1 500 damaged files, no sanitizer report (two findings fixed on the way: a record whose CIGAR does not fit its
// sequence made the trimming throw across the ABI; an end position overflowed 32 bits).  After the reader's decoding loops
// were vectorised: seeds 101-108 (2 400 files), one more finding fixed -- a read name without its terminating NUL let strlen
// run past the name buffer (tests/test_bam_reader.py::test_read_name_without_terminator) -- then clean.
#include <cstdio>
#include <string>
#include "longtr_b200.h"
int main(int argc, char** argv) {
  int opened = 0, fetched = 0, idx = 0;
  for (int i = 0; i < 300; ++i) {
    const std::string path = "/tmp/mut/m" + std::to_string(i) + ".bam";
    ltr_bam* bam = nullptr;
    if (ltr_bam_open(path.c_str(), nullptr, &bam) != LTR_OK) continue;
    ++opened;
    if (ltr_bam_build_index(bam) == LTR_OK) ++idx;
    for (int q = 0; q < 6; ++q) {
      ltr_bam_reads* out = nullptr;
      const long beg = (q * 37 % 20) * 4000;
      if (ltr_bam_fetch(bam, q == 5 ? -1 : 0, beg, beg + 6000, q & 1, &out) == LTR_OK) { ++fetched; ltr_bam_reads_free(out); }
    }
    // the region loop on top of it
    ltr_region_params rp; ltr_region_params_default(&rp);
    const ltr_bam* bams[1] = {bam};
    static std::string ref(100000, 'A');
    for (int l = 0; l < 20; l += 3) {
      ltr_region_reads* rr = nullptr;
      if (ltr_region_collect(bams, 1, "chrS", l * 4000 + 600, l * 4000 + 700, (const uint8_t*)ref.data(), 0, (long)ref.size(), &rp, &rr) == LTR_OK) {
        ltr_candidates* c = nullptr;
        ltr_candidate_alleles(rr, l * 4000 + 600, l * 4000 + 700, 2, (const uint8_t*)ref.data(), 0, (long)ref.size(), 5, &c);
        ltr_candidates_free(c);
      }
      ltr_region_reads_free(rr);
    }
    ltr_bam_close(bam);
  }
  printf("opened %d indexed %d fetched %d\n", opened, idx, fetched);
  return 0;
}
