// TEST INFRASTRUCTURE ONLY -- the reference's own alignment input (BamCramReader / BamAlignment, src/bam_io.{h,cpp},
// compiled IN PLACE from /root/reference) running on top of integration/hts_compat.cpp + the library's BAM reader, and its
// read trimming (BamAlignment::TrimAlignment, bam_io.cpp:267-372) as GenotyperBamProcessor::left_align_reads calls it
// (genotyper_bam_processor.cpp:55-62).  Prints one line per read so that tests can compare with ltr_bam_* / ltr_region_*.
#include <stdint.h>

#include <stdlib.h>
#include <string.h>

#include <sstream>
#include <string>

#include "bam_io.h"

// Every alignment BamCramReader yields for chrom:[start, end) -- "name pos end flag mapq cigar bases quals hp" per line --
// and, with trim_lo <= trim_hi, the same reads after TrimAlignment(trim_lo, trim_hi) (only reads with pos <= span_lo and
// end >= span_hi, as left_align_reads filters them), with the `deleted` flag appended.
extern "C" char* ltr_ref_io_region(const char* path, const char* chrom, int32_t start, int32_t end, int32_t span_lo,
                                   int32_t span_hi, int32_t trim_lo, int32_t trim_hi) {
  BamCramReader reader(path, "");
  std::ostringstream out;
  if (reader.SetRegion(chrom, start, end)) {
    BamAlignment next;
    while (reader.GetNextAlignment(next)) {
      BamAlignment aln(next);  // the region loop trims copies (vector elements): `deleted_` starts out false for each read
      const bool trim = trim_lo <= trim_hi;
      if (trim) {
        if (aln.Position() > span_lo || aln.GetEndPosition() < span_hi) continue;
        aln.TrimAlignment(trim_lo, trim_hi);
      }
      int64_t hp = 0;
      if (aln.HasTag("HP")) aln.GetIntTag("HP", hp);
      out << aln.Name() << ' ' << aln.Position() << ' ' << aln.GetEndPosition() << ' ' << (aln.IsReverseStrand() ? 1 : 0)
          << ' ' << aln.MapQuality() << ' ';
      const std::vector<CigarOp>& cig = aln.CigarData();
      if (cig.empty()) out << '*';
      for (size_t k = 0; k < cig.size(); ++k) out << cig[k].Length << cig[k].Type;
      const std::string bases = aln.QueryBases(), quals = aln.Qualities();
      out << ' ' << (bases.empty() ? "*" : bases) << ' ' << (quals.empty() ? "*" : quals) << ' ' << hp;
      if (trim) out << ' ' << (aln.GetDeleted() ? 1 : 0);
      out << '\n';
    }
  }
  const std::string s = out.str();
  char* r = (char*)malloc(s.size() + 1);
  memcpy(r, s.c_str(), s.size() + 1);
  return r;
}
extern "C" void ltr_ref_io_free(char* p) { free(p); }
