// stutter_host.cpp -- homopolymer / --stutter-align-len path of HapAligner::process_reads
// (reference src/SeqAlignment/HapAligner.cpp:567-579, 855-975): host preparation for kernel 2.
//
// Per pooled read: calc_seed_base on the host (integer CIGAR work); reads without a seed get the reference's
// all-zero row (:570-574).  Everything else -- both flank alignments per (read, haplotype), the stutter row
// and the seed join -- runs on the GPU through ltr_stutter_ll (stutter_pair_kernel).
#include "longtr_host.h"

namespace ltr {

void HapAligner::process_reads_short(const std::vector<Alignment>& alns, int init_read_index, const BaseQuality* bq,
                                     const std::vector<bool>& realign_read, double* aln_probs, int* seed_positions) {
  (void)bq;  // the device uses the same BaseQuality tables, built by ltr_stutter_ll with the host's libm
  if (fw_haplotype_->num_blocks() != 3 || fw_haplotype_->get_block(0)->num_options() != 1 ||
      fw_haplotype_->get_block(2)->num_options() != 1 || fw_haplotype_->get_block(0)->get_repeat_info() != NULL ||
      fw_haplotype_->get_block(2)->get_repeat_info() != NULL) {
    status_ = LTR_ERR_UNSUPPORTED;  // flank / repeat / flank is the only layout LongTR builds (HaplotypeGenerator.cpp:580-607)
    return;
  }
  const HapBlock* left = fw_haplotype_->get_block(0);
  const HapBlock* rep = fw_haplotype_->get_block(1);
  const HapBlock* right = fw_haplotype_->get_block(2);
  const int H = rep->num_options();  // == num_combs, column = allele index (Haplotype.cpp:157-196 with one free block)
  const std::string& lf = left->get_seq(0);
  const std::string& rf = right->get_seq(0);

  std::string allele_bytes, read_bytes, qual_bytes;
  std::vector<uint32_t> allele_off(1, 0), read_off(1, 0);
  std::vector<uint8_t> realign_allele((size_t)H), realign_row(alns.size());
  for (int a = 0; a < H; ++a) {
    allele_bytes += rep->get_seq(a);
    allele_off.push_back((uint32_t)allele_bytes.size());
    realign_allele[a] = realign_to_hap_[a] ? 1 : 0;
  }
  std::vector<int32_t> seeds(alns.size(), -1);
  for (size_t i = 0; i < alns.size(); ++i) {
    const Alignment& aln = alns[i];
    realign_row[i] = realign_read[i] ? 1 : 0;
    if (realign_read[i]) {
      if (aln.get_sequence().size() != aln.get_base_qualities().size()) {
        status_ = LTR_ERR_INVALID;  // the reference asserts (HapAligner.cpp:857)
        return;
      }
      const int seed = calc_seed_base(aln);
      if (seed == -2) {
        status_ = LTR_ERR_INVALID;
        return;
      }
      seeds[i] = seed;
      seed_positions[init_read_index + i] = seed;
    }
    read_bytes += aln.get_sequence();
    qual_bytes += realign_read[i] ? aln.get_base_qualities() : std::string(aln.get_sequence().size(), '!');
    read_off.push_back((uint32_t)read_bytes.size());
  }
  const uint32_t lab[2] = {0u, (uint32_t)H}, lrb[2] = {0u, (uint32_t)alns.size()};
  const uint32_t lfo[2] = {0u, (uint32_t)lf.size()}, rfo[2] = {0u, (uint32_t)rf.size()};
  double stutter[6];
  const StutterModel& model = rep->get_repeat_info()->get_stutter_model();
  model.get_parameters(stutter);
  const int32_t motif_len = model.period();
  ltr_stutter_batch b;
  b.n_loci = 1;
  b.locus_allele_begin = lab;
  b.locus_read_begin = lrb;
  b.lflank_off = lfo;
  b.lflank_bytes = reinterpret_cast<const uint8_t*>(lf.data());
  b.rflank_off = rfo;
  b.rflank_bytes = reinterpret_cast<const uint8_t*>(rf.data());
  b.allele_off = allele_off.data();
  b.allele_bytes = reinterpret_cast<const uint8_t*>(allele_bytes.data());
  b.stutter = stutter;
  b.motif_len = &motif_len;
  b.read_off = read_off.data();
  b.read_bytes = reinterpret_cast<const uint8_t*>(read_bytes.data());
  b.qual_bytes = reinterpret_cast<const uint8_t*>(qual_bytes.data());
  b.read_seed = seeds.data();
  b.realign_allele = realign_allele.data();
  b.realign_read = realign_row.data();
  ltr_params p;
  p.ins_ins = model_.LOG_INS_TO_INS;
  p.ins_match = model_.LOG_INS_TO_MATCH;
  p.del_del = model_.LOG_DEL_TO_DEL;
  p.del_match = model_.LOG_DEL_TO_MATCH;
  p.match_match = model_.LOG_MATCH_TO_MATCH;
  p.match_ins = model_.LOG_MATCH_TO_INS;
  p.match_del = model_.LOG_MATCH_TO_DEL;
  p.indel_flank_len = INDEL_FLANK_LEN_;
  status_ = ltr_stutter_ll(ctx_, &p, &b, aln_probs + (size_t)init_read_index * H, NULL);
}

}  // namespace ltr
