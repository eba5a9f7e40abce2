// TEST INFRASTRUCTURE ONLY -- the reference's own HaplotypeGenerator (add_haplotype_block + fuse_haplotype_blocks,
// src/SeqAlignment/HaplotypeGenerator.cpp:521-607, compiled IN PLACE) on the flat reads ltr_region_collect produces.  Built
// twice (oracle/build_ref.sh): libltr_ref_hapgen.so with the spoa stand-in throwing instead of aborting ("needs assembly" is
// an answer), libltr_ref_hapgen_poa.so with the spoa names served by oracle/poa_restatement.hpp, so that the reference's own
// clustering / merging / support logic around the consensus (:397-471) runs to completion; that build also exports the
// restated consensus alone (ltr_oracle_poa_consensus).  Private members are reached with -fno-access-control; the reference's
// sources are not touched.  Output: one text record, see tests/test_candidate_alleles.py.
#include <limits.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <sstream>
#include <string>
#include <vector>

#include "SeqAlignment/AlignmentData.h"
#include "SeqAlignment/HapBlock.h"
#include "SeqAlignment/HaplotypeGenerator.h"
#include "region.h"
#include "stutter_model.h"

// reads: sample-major flat arrays (ltr_region_reads layout); chrom_seq: the chromosome (or a stand-in whose position 0 is
// reference position 0).  Returns "status=<msg>" or "ok block=<s>,<e> lflank=<..> rflank=<..> alleles=<a0>,<a1>,..".
extern "C" char* ltr_ref_candidate_alleles(uint32_t n_samples, uint32_t n_reads, const int32_t* read_sample,
                                           const int32_t* read_start, const int32_t* read_stop, const uint32_t* read_off,
                                           const uint8_t* read_bytes, const uint32_t* cigar_off, const uint32_t* cigar_ops,
                                           const uint8_t* hap_gen_ok, const uint8_t* deleted, int32_t region_start,
                                           int32_t region_stop, const char* motif, const char* chrom_seq,
                                           int32_t indel_flank_len) {
  std::vector<std::vector<Alignment> > alns(n_samples);
  int32_t all_min = INT_MAX, all_max = INT_MIN;
  for (uint32_t i = 0; i < n_reads; ++i) {
    all_min = std::min(all_min, read_start[i]);
    all_max = std::max(all_max, read_stop[i]);
    if (!hap_gen_ok[i]) continue;
    const std::string seq((const char*)read_bytes + read_off[i], read_off[i + 1] - read_off[i]);
    std::string aln_str;  // the gapped alignment string left_align_reads builds (genotyper_bam_processor.cpp:79-129)
    size_t si = 0;
    Alignment a(read_start[i], read_stop[i], false, deleted[i] != 0, "r", std::string(seq.size(), 'I'), seq, "");
    for (uint32_t k = cigar_off[i]; k < cigar_off[i + 1]; ++k) {
      const char t = "MIDNSHP=X"[cigar_ops[k] & 15];
      const int32_t n = (int32_t)(cigar_ops[k] >> 4);
      a.add_cigar_element(CigarElement(t, n));
      if (t == 'D') aln_str += std::string((size_t)n, '-');
      else {
        aln_str += seq.substr(si, (size_t)n);
        si += (size_t)n;
      }
    }
    a.set_alignment(aln_str);
    a.set_hap_gen_info(std::vector<bool>(1, true));
    alns[(size_t)read_sample[i]].push_back(a);
  }
  std::ostringstream out;
  try {
    HaplotypeGenerator gen(all_min, all_max, indel_flank_len);
    Region region("chr", region_start, region_stop, motif);
    StutterModel model(0.9, 0.01, 0.01, 0.9, 0.001, 0.001, motif);
    const std::string chrom(chrom_seq);
    if (!gen.add_haplotype_block(region, chrom, alns, std::vector<std::string>(), &model)) {
      out << "status=" << gen.failure_msg();
    } else if (!gen.fuse_haplotype_blocks(chrom)) {
      out << "status=" << gen.failure_msg();
    } else {
      const std::vector<HapBlock*> blocks = gen.get_haplotype_blocks();
      out << "ok block=" << blocks[1]->start() << ',' << blocks[1]->end() << " lstart=" << blocks[0]->start()
          << " lflank=" << blocks[0]->get_seq(0) << " rflank=" << blocks[2]->get_seq(0) << " alleles=";
      for (int k = 0; k < blocks[1]->num_options(); ++k) out << (k ? "," : "") << '[' << blocks[1]->get_seq(k) << ']';
      out << " inexact=";
      for (int k = 0; k < blocks[1]->num_options(); ++k) out << (k == 0 ? 0 : (int)blocks[1]->get_inexact(k));
    }
  } catch (const int&) {
    out << "status=needs assembly";
  }
  const std::string s = out.str();
  char* r = (char*)malloc(s.size() + 1);
  memcpy(r, s.c_str(), s.size() + 1);
  return r;
}

// readRegions + orderRegions (src/region.cpp:26-75) on a well-formed region file (a malformed one makes the reference exit).
extern "C" char* ltr_ref_read_regions(const char* path, uint32_t max_regions, const char* chrom_limit) {
  std::vector<Region> regions;
  std::ostringstream log, out;
  readRegions(path, max_regions, chrom_limit ? chrom_limit : "", regions, log);
  orderRegions(regions);
  for (const Region& r : regions)
    out << r.chrom() << ' ' << r.start() << ' ' << r.stop() << ' ' << r.period() << ' ' << (r.name().empty() ? "." : r.name())
        << ' ' << r.motif() << '\n';
  const std::string s = out.str();
  char* res = (char*)malloc(s.size() + 1);
  memcpy(res, s.c_str(), s.size() + 1);
  return res;
}

#ifdef LTR_SPOA_RESTATEMENT
// HaplotypeGenerator::poa (:167-199) itself on a list of sequences (fewer than 30: no random sampling).
extern "C" char* ltr_ref_poa(uint32_t n_seqs, const uint32_t* seq_off, const uint8_t* seq_bytes) {
  std::vector<std::string> seqs;
  for (uint32_t i = 0; i < n_seqs; ++i) seqs.push_back(std::string((const char*)seq_bytes + seq_off[i], seq_off[i + 1] - seq_off[i]));
  HaplotypeGenerator gen(0, 1, 5);
  std::string consensus;
  gen.poa(seqs, consensus);
  char* r = (char*)malloc(consensus.size() + 1);
  memcpy(r, consensus.c_str(), consensus.size() + 1);
  return r;
}
#endif
