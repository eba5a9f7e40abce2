// hap_aligner.cpp -- host mirror of LongTR's HapAligner (reference src/SeqAlignment/HapAligner.{h,cpp}).
//
// What the reference does per (read, haplotype) pair on one CPU thread is done here as: trim every
// pooled read once on the host (integer CIGAR work, HapAligner.cpp:346-465), enumerate the candidate
// haplotypes once in gray-code order (Haplotype.cpp:157-196), flatten both into one ltr_viterbi_batch and
// let the GPU evaluate all pairs (ltr_viterbi_ll -> viterbi_stream_kernel).  Loci whose repeat has period 1
// take the homopolymer path when SWITCH_OLD_ALIGN_LEN != 0 (HapAligner.cpp:552), see process_reads_short.
#include <string.h>

#include <algorithm>

#include "longtr_host.h"

namespace ltr {

static const int32_t MIN_SEED_DIST = 5;  // HapAligner.cpp:17

HapAligner::HapAligner(Haplotype* haplotype, std::vector<bool>& realign_to_haplotype, int INDEL_FLANK_LEN,
                       int SWITCH_OLD_ALIGN_LEN, std::vector<float>& alignment_model_params, ltr_ctx* ctx)
    : fw_haplotype_(haplotype), realign_to_hap_(realign_to_haplotype), INDEL_FLANK_LEN_(INDEL_FLANK_LEN),
      SWITCH_OLD_ALIGN_LEN_(SWITCH_OLD_ALIGN_LEN), ctx_(ctx), status_(LTR_OK) {
  if ((int)realign_to_hap_.size() != haplotype->num_combs()) status_ = LTR_ERR_INVALID;
  for (int i = 0; i < fw_haplotype_->num_blocks(); ++i) {
    const HapBlock* block = fw_haplotype_->get_block(i);
    if (block->get_repeat_info() != NULL) {
      repeat_starts_.push_back(block->start());
      repeat_ends_.push_back(block->end());
    }
  }
  if (repeat_starts_.empty() || fw_haplotype_->num_blocks() < 2) status_ = LTR_ERR_INVALID;
  ltr_params p;
  ltr_params_default(&p);  // Dindel defaults, HapAligner.h:118
  model_.MAX_HOMOP_LEN = 10;
  if (alignment_model_params.size() >= 7) {
    p.ins_ins = alignment_model_params[0];
    p.ins_match = alignment_model_params[1];
    p.del_del = alignment_model_params[2];
    p.del_match = alignment_model_params[3];
    p.match_match = alignment_model_params[4];
    p.match_ins = alignment_model_params[5];
    p.match_del = alignment_model_params[6];
  } else if (!alignment_model_params.empty()) {
    status_ = LTR_ERR_INVALID;  // the reference would read past the end of the vector
  }
  model_.LOG_INS_TO_INS = p.ins_ins;
  model_.LOG_INS_TO_MATCH = p.ins_match;
  model_.LOG_DEL_TO_DEL = p.del_del;
  model_.LOG_DEL_TO_MATCH = p.del_match;
  model_.LOG_MATCH_TO_MATCH = p.match_match;
  model_.LOG_MATCH_TO_INS = p.match_ins;
  model_.LOG_MATCH_TO_DEL = p.match_del;
}

// ---------------------------------------------------------------------------------------------------
// Cursor that hands out the CIGAR one base-unit at a time from either end (the reference edits a copy
// of the element vector in place, HapAligner.cpp:356-460).
namespace {
struct CigarWindow {
  std::vector<CigarElement> ops;
  size_t f, b;  // live elements are [f, b)
  explicit CigarWindow(const std::vector<CigarElement>& c) : ops(c), f(0), b(c.size()) {}
  bool live() const { return f < b; }
  char front() const { return ops[f].get_type(); }
  char back() const { return ops[b - 1].get_type(); }
  void pop_front() {
    if (ops[f].get_num() == 1) ++f;
    else ops[f].set_num(ops[f].get_num() - 1);
  }
  void pop_back() {
    if (ops[b - 1].get_num() == 1) --b;
    else ops[b - 1].set_num(ops[b - 1].get_num() - 1);
  }
};
enum OpClass { OP_ALIGNED, OP_DEL, OP_READ_ONLY, OP_NONE, OP_BAD };
inline OpClass classify(char op) {
  switch (op) {
    case 'M': case '=': case 'X': return OP_ALIGNED;
    case 'D': return OP_DEL;
    case 'I': case 'S': return OP_READ_ONLY;
    case 'H': return OP_NONE;
    default: return OP_BAD;
  }
}
}  // namespace

bool HapAligner::trim_alignment(const Alignment& aln, std::string& trimmed_seq) const {
  const int32_t pad = INDEL_FLANK_LEN_;
  const int32_t lo = repeat_starts_[0] - pad, hi = repeat_ends_[0] + pad;
  int32_t start_pos = aln.get_start() + 1, end_pos = aln.get_stop() + 1;
  int32_t ltrim = 0, rtrim = 0;
  CigarWindow w(aln.get_cigar_list());
  // bases left of the window
  while (start_pos <= lo && w.live()) {
    switch (classify(w.front())) {
      case OP_ALIGNED: ltrim++; start_pos++; break;
      case OP_DEL: start_pos++; break;
      case OP_READ_ONLY: ltrim++; break;
      case OP_NONE: break;
      default: return false;
    }
    w.pop_front();
  }
  // inside the left pad: a deletion pulls one upstream base back in, insertions stay
  for (int32_t mid = start_pos; mid > lo && mid <= lo + pad && w.live(); w.pop_front()) {
    switch (classify(w.front())) {
      case OP_ALIGNED: mid++; break;
      case OP_DEL: ltrim--; mid++; break;
      case OP_READ_ONLY: case OP_NONE: break;
      default: return false;
    }
  }
  // bases right of the window
  while (end_pos > hi && w.live()) {
    switch (classify(w.back())) {
      case OP_ALIGNED: rtrim++; end_pos--; break;
      case OP_DEL: end_pos--; break;
      case OP_READ_ONLY: rtrim++; break;
      case OP_NONE: break;
      default: return false;
    }
    w.pop_back();
  }
  // inside the right pad
  for (int32_t mid = end_pos; mid > hi - pad && mid <= hi && w.live(); w.pop_back()) {
    switch (classify(w.back())) {
      case OP_ALIGNED: mid--; break;
      case OP_DEL: rtrim--; mid--; break;
      case OP_READ_ONLY: case OP_NONE: break;
      default: return false;
    }
  }
  ltrim = std::max(ltrim, 0);
  rtrim = std::max(rtrim, 0);
  const int32_t len = (int32_t)aln.get_sequence().size();
  if (ltrim + rtrim > len) return false;  // the reference asserts (HapAligner.cpp:463)
  trimmed_seq = aln.get_sequence().substr(ltrim, len - ltrim - rtrim);
  return true;
}

// ---------------------------------------------------------------------------------------------------
// HapAligner.cpp:467-491: position inside [region_start, region_end] that is farthest from any repeat block
void HapAligner::calc_best_seed_position(int32_t region_start, int32_t region_end, int32_t& best_dist,
                                         int32_t& best_pos) const {
  best_dist = best_pos = -1;
  int32_t pos = region_start;
  size_t k = 0;
  while (k < repeat_starts_.size() && pos <= region_end) {
    if (pos < repeat_starts_[k]) {
      const int32_t dist = 1 + (std::min(region_end, repeat_starts_[k] - 1) - pos) / 2;
      if (dist >= best_dist) {
        best_dist = dist;
        best_pos = dist - 1 + pos;
      }
      pos = repeat_ends_[k++];
    } else if (pos < repeat_ends_[k]) {
      pos = repeat_ends_[k++];
    } else {
      k++;
    }
  }
  if (pos <= region_end) {
    const int32_t dist = 1 + (region_end - pos) / 2;
    if (dist >= best_dist) {
      best_dist = dist;
      best_pos = dist - 1 + pos;
    }
  }
}

// HapAligner.cpp:493-542.  Returns -2 for a CIGAR operation the reference dies on.
int HapAligner::calc_seed_base(const Alignment& aln) const {
  int32_t pos = aln.get_start();
  int best_seed = -1, cur_base = 0, max_dist = MIN_SEED_DIST;
  const std::vector<CigarElement>& cigar = aln.get_cigar_list();
  for (size_t c = 0; c < cigar.size(); ++c) {
    const int num = cigar[c].get_num();
    switch (cigar[c].get_type()) {
      case '=': {
        int32_t min_region = std::max(pos, fw_haplotype_->get_first_block()->start());
        int32_t max_region = std::min(pos + num - 1, fw_haplotype_->get_last_block()->end() - 1);
        if (min_region <= max_region) {
          int32_t distance, dist_pos;
          calc_best_seed_position(min_region, max_region, distance, dist_pos);
          if (distance >= max_dist) {
            max_dist = distance;
            best_seed = cur_base + (dist_pos - pos);
          }
        }
        pos += num;
        cur_base += num;
        break;
      }
      case 'I': cur_base += num; break;
      case 'X': pos += num; cur_base += num; break;
      case 'D': pos += num; break;
      default: return -2;
    }
  }
  if (best_seed < -1 || best_seed == 0 || best_seed >= ((int)aln.get_sequence().size()) - 1) return -1;
  return best_seed;
}

// ---------------------------------------------------------------------------------------------------
void HapAligner::process_reads(const std::vector<Alignment>& alignments, int init_read_index,
                               const BaseQuality* base_quality, const std::vector<bool>& realign_read,
                               double* aln_probs, int* seed_positions) {
  if (status_ != LTR_OK) return;
  if (ctx_ == NULL || alignments.size() != realign_read.size()) {
    status_ = LTR_ERR_INVALID;
    return;
  }
  // HapAligner.cpp:552: the homopolymer path is chosen per locus, by the period of block 1
  if (uses_short_path())
    process_reads_short(alignments, init_read_index, base_quality, realign_read, aln_probs, seed_positions);
  else
    process_reads_long(alignments, init_read_index, realign_read, aln_probs, seed_positions);
}

bool HapAligner::uses_short_path() const {
  const RepeatStutterInfo* info = fw_haplotype_->get_block(1)->get_repeat_info();
  return info != NULL && info->get_period() == 1 && SWITCH_OLD_ALIGN_LEN_ != 0;
}

void HapAligner::fill_params(ltr_params& p) const {
  p.ins_ins = model_.LOG_INS_TO_INS;
  p.ins_match = model_.LOG_INS_TO_MATCH;
  p.del_del = model_.LOG_DEL_TO_DEL;
  p.del_match = model_.LOG_DEL_TO_MATCH;
  p.match_match = model_.LOG_MATCH_TO_MATCH;
  p.match_ins = model_.LOG_MATCH_TO_INS;
  p.match_del = model_.LOG_MATCH_TO_DEL;
  p.indel_flank_len = INDEL_FLANK_LEN_;
}

// Long path, first half (HapAligner.cpp:556-566, 812-854): the locus' haplotypes and trimmed reads, flattened.
bool HapAligner::prepare_long(const std::vector<Alignment>& alns, int init_read_index, const std::vector<bool>& realign_read,
                              int* seed_positions, LongPart& part, std::string& hap_bytes, std::vector<uint32_t>& hap_off,
                              std::string& read_bytes, std::vector<uint32_t>& read_off) {
  if (status_ != LTR_OK) return false;
  if (alns.size() != realign_read.size()) {
    status_ = LTR_ERR_INVALID;
    return false;
  }
  // haplotypes in column order; only the ones flagged for realignment travel to the GPU
  fw_haplotype_->reset();
  do {
    const int c = fw_haplotype_->cur_index();
    if (!realign_to_hap_[c]) continue;  // HapAligner.cpp:841-845: slot left untouched
    part.hap_cols.push_back(c);
    hap_bytes += fw_haplotype_->get_seq();
    hap_off.push_back((uint32_t)hap_bytes.size());
  } while (fw_haplotype_->next());
  fw_haplotype_->reset();

  std::string trimmed;
  for (size_t i = 0; i < alns.size(); ++i) {
    if (!realign_read[i]) continue;
    if (alns[i].get_sequence().size() != alns[i].get_base_qualities().size() && !alns[i].get_base_qualities().empty()) {
      status_ = LTR_ERR_INVALID;  // the reference asserts (HapAligner.cpp:816)
      return false;
    }
    seed_positions[init_read_index + i] = (int)alns[i].get_sequence().size() - 1;  // HapAligner.cpp:562-563
    if (!trim_alignment(alns[i], trimmed)) {
      status_ = LTR_ERR_INVALID;
      return false;
    }
    if (trimmed.empty()) {  // HapAligner.cpp:820-823: 10 bp pseudo read from the flank blocks
      const std::string& lf = fw_haplotype_->get_first_block()->get_seq(0);
      const std::string& rf = fw_haplotype_->get_last_block()->get_seq(0);
      if (lf.size() < 5 || rf.size() < 5) {
        status_ = LTR_ERR_INVALID;
        return false;
      }
      trimmed = lf.substr(lf.size() - 5, 5) + rf.substr(0, 5);
    }
    const size_t nul = trimmed.find('\0');  // std::string(const char*) in the reference stops at a NUL
    if (nul != std::string::npos) trimmed.resize(nul);
    if (trimmed.empty()) {
      status_ = LTR_ERR_INVALID;
      return false;
    }
    part.read_rows.push_back((int)i);
    read_bytes += trimmed;
    read_off.push_back((uint32_t)read_bytes.size());
  }
  return true;
}

// Long path, second half: aln_probs[(init_read_index + read) * num_combs + column] (HapAligner.cpp:549).
void HapAligner::scatter_long(const LongPart& part, const double* ll, int init_read_index, double* aln_probs) const {
  const int ncombs = fw_haplotype_->num_combs();
  for (size_t r = 0; r < part.read_rows.size(); ++r) {
    double* row = aln_probs + (size_t)(init_read_index + part.read_rows[r]) * ncombs;
    for (size_t h = 0; h < part.hap_cols.size(); ++h) row[part.hap_cols[h]] = ll[r * part.hap_cols.size() + h];
  }
}

// Long path for one locus: one GPU batch for the whole locus.
void HapAligner::process_reads_long(const std::vector<Alignment>& alns, int init_read_index,
                                    const std::vector<bool>& realign_read, double* aln_probs, int* seed_positions) {
  LongPart part;
  std::vector<uint32_t> hap_off(1, 0), read_off(1, 0);
  std::string hap_bytes, read_bytes;
  if (!prepare_long(alns, init_read_index, realign_read, seed_positions, part, hap_bytes, hap_off, read_bytes, read_off))
    return;
  if (part.hap_cols.empty() || part.read_rows.empty()) return;
  const uint32_t lhb[2] = {0u, (uint32_t)part.hap_cols.size()}, lrb[2] = {0u, (uint32_t)part.read_rows.size()};
  ltr_viterbi_batch b;
  b.n_loci = 1;
  b.locus_hap_begin = lhb;
  b.locus_read_begin = lrb;
  b.hap_off = hap_off.data();
  b.hap_bytes = reinterpret_cast<const uint8_t*>(hap_bytes.data());
  b.read_off = read_off.data();
  b.read_bytes = reinterpret_cast<const uint8_t*>(read_bytes.data());
  ltr_params p;
  fill_params(p);
  std::vector<double> ll(part.hap_cols.size() * part.read_rows.size());
  status_ = ltr_viterbi_ll(ctx_, &p, &b, ll.data(), NULL);
  if (status_ != LTR_OK) return;
  scatter_long(part, ll.data(), init_read_index, aln_probs);
}

}  // namespace ltr
