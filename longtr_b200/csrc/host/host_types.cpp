// host_types.cpp -- the small value classes of the host mirror (see longtr_host.h for the reference
// file:line each one follows): Alignment, BaseQuality, StutterModel, RepeatStutterInfo, HapBlock,
// Haplotype (gray-code iterator), ReadPooler and the mathops helpers the genotype extraction uses.
#include <math.h>
#include <string.h>

#include <algorithm>

#include "longtr_host.h"

namespace ltr {

// ---------------------------------------------------------------------------------------------------
std::string Alignment::getCigarString() const {
  std::string s;
  for (size_t i = 0; i < cigar_list_.size(); ++i) {
    s += std::to_string(cigar_list_[i].get_num());
    s += cigar_list_[i].get_type();
  }
  return s;
}

bool Alignment::set_cigar_string(const char* cigar) {
  cigar_list_.clear();
  if (cigar == NULL) return false;
  long num = 0;
  bool have_num = false;
  for (const char* p = cigar; *p; ++p) {
    if (*p >= '0' && *p <= '9') {
      num = num * 10 + (*p - '0');
      have_num = true;
      if (num > 0x7fffffffL) return false;
    } else {
      if (!have_num) return false;
      cigar_list_.push_back(CigarElement(*p, (int)num));
      num = 0;
      have_num = false;
    }
  }
  return !have_num;
}

// ---------------------------------------------------------------------------------------------------
// base_quality.h:29-38: q = 1..41 -> log(1-10^(-q/10)), log(10^(-q/50)); q = 0 -> (-100, 0)
BaseQuality::BaseQuality() {
  memset(log_correct_, 0, sizeof(log_correct_));
  memset(log_error_, 0, sizeof(log_error_));
  log_correct_[0] = -100;
  log_error_[0] = 0;
  const int max_index = MAX_BASE_QUALITY - MIN_BASE_QUALITY;
  for (int i = 1; i <= max_index; ++i) {
    log_correct_[i] = log(1.0 - pow(10.0, i / (-10.0)));
    log_error_[i] = log(pow(10.0, i / (-10.0) / 5.0));
  }
}

void BaseQuality::byte_tables(double* log_correct256, double* log_error256) const {
  for (int b = 0; b < 256; ++b) {
    const char q = (char)b;  // the reference indexes with a (signed) char
    log_correct256[b] = log_prob_correct(q);
    log_error256[b] = log_prob_error(q);
  }
}

std::string BaseQuality::median_base_qualities(const std::vector<const std::string*>& qualities) const {
  if (qualities.empty()) return std::string();
  const size_t n = qualities[0]->size();
  for (size_t i = 0; i < qualities.size(); ++i)
    if (qualities[i]->size() != n) return std::string();
  std::string out(n, 'N');
  std::vector<char> column(qualities.size());
  for (size_t i = 0; i < n; ++i) {
    for (size_t j = 0; j < qualities.size(); ++j) column[j] = (*qualities[j])[i];
    std::sort(column.begin(), column.end());
    out[i] = column[column.size() / 2];  // upper median, base_quality.cpp:25
  }
  return out;
}

// ---------------------------------------------------------------------------------------------------
StutterModel::StutterModel(double inframe_geom, double inframe_up, double inframe_down, double outframe_geom,
                           double outframe_up, double outframe_down, const std::string& motif)
    : in_geom_(inframe_geom), in_up_(inframe_up), in_down_(inframe_down), out_geom_(outframe_geom),
      out_up_(outframe_up), out_down_(outframe_down), motif_len_((int)motif.size()), motif_(motif) {
  valid_ = inframe_geom > 0.0 && inframe_geom < 1.0 && outframe_geom > 0.0 && outframe_geom < 1.0 &&
           inframe_up > 0.0 && inframe_down > 0.0 && outframe_up > 0.0 && outframe_down > 0.0 &&
           inframe_up + inframe_down + outframe_up + outframe_down < 1.0 && motif_len_ >= 1;
  in_log_step_ = log(1 - inframe_geom);
  in_log_nostep_ = log(inframe_geom);
  in_log_up_ = log(inframe_up);
  in_log_down_ = log(inframe_down);
  out_log_step_ = log(1 - outframe_geom);
  out_log_nostep_ = log(outframe_geom);
  out_log_up_ = log(outframe_up);
  out_log_down_ = log(outframe_down);
  log_equal_ = log(1 - inframe_up - inframe_down - outframe_up - outframe_down);
}

// stutter_model.cpp:29-53
double StutterModel::log_stutter_pmf(int sample_bps, int read_bps) const {
  const int bp_diff = read_bps - sample_bps;
  if (bp_diff % motif_len_ != 0) {
    const int eff_diff = bp_diff - (bp_diff / motif_len_);
    if (eff_diff < 0) return out_log_down_ + out_log_nostep_ + out_log_step_ * (-eff_diff - 1);
    return out_log_up_ + out_log_nostep_ + out_log_step_ * (eff_diff - 1);
  }
  const int rep_diff = bp_diff / motif_len_;
  if (rep_diff == 0) return log_equal_;
  if (rep_diff < 0) return in_log_down_ + in_log_nostep_ + in_log_step_ * (-rep_diff - 1);
  return in_log_up_ + in_log_nostep_ + in_log_step_ * (rep_diff - 1);
}

// RepeatStutterInfo.h:53-61
double RepeatStutterInfo::log_prob_pcr_artifact(int seq_index, int artifact_size) const {
  const double LARGE_NEGATIVE = -10e6;
  const int allele = allele_sizes_.at(seq_index);
  const int read_size = allele + artifact_size;
  if (artifact_size == 0) return model_.log_stutter_pmf(allele, read_size);
  if (artifact_size > 0) return artifact_size > max_ins_ ? LARGE_NEGATIVE : model_.log_stutter_pmf(allele, read_size);
  return (artifact_size < max_del_ || read_size < 0) ? LARGE_NEGATIVE : model_.log_stutter_pmf(allele, read_size);
}

// ---------------------------------------------------------------------------------------------------
int HapBlock::max_size() const {
  int m = 0;
  for (size_t i = 0; i < seqs_.size(); ++i) m = std::max(m, (int)seqs_[i].size());
  return m;
}

Haplotype::Haplotype(std::vector<HapBlock*>& blocks) : blocks_(blocks), max_size_(0), fixed_(false) {
  for (size_t i = 0; i < blocks_.size(); ++i) {
    nopts_.push_back(blocks_[i]->num_options());
    max_size_ += blocks_[i]->max_size();
  }
  dirs_.resize(blocks_.size());
  factors_.resize(blocks_.size());
  counts_.resize(blocks_.size());
  init();
}

// Haplotype.cpp:123-155 (forward order: block 0 is the fastest digit)
void Haplotype::init() {
  ncombs_ = 1;
  cur_size_ = 0;
  for (size_t i = 0; i < blocks_.size(); ++i) {
    factors_[i] = ncombs_;
    ncombs_ *= nopts_[i];
    dirs_[i] = 1;
    counts_[i] = 0;
    cur_size_ += blocks_[i]->size(0);
  }
  counter_ = 0;
  last_changed_ = -1;
}

void Haplotype::reset() { init(); }

// Haplotype.cpp:157-196: mixed-radix reflected gray code
bool Haplotype::next() {
  if (fixed_ || counter_ == ncombs_ - 1) return false;
  int index = -1;
  int t = counter_ + 1;
  for (int j = (int)blocks_.size() - 1; j >= 0; --j) {
    t %= factors_[j];
    if (t == 0) {
      index = j;
      break;
    }
  }
  if (index < 0) return false;
  last_changed_ = index;
  cur_size_ -= blocks_[index]->size(counts_[index]);
  counts_[index] += dirs_[index];
  cur_size_ += blocks_[index]->size(counts_[index]);
  if (counts_[index] == 0 || counts_[index] == nopts_[index] - 1) dirs_[index] *= -1;
  counter_++;
  return true;
}

bool Haplotype::go_to(int hap_index) {
  if (hap_index < 0 || hap_index >= ncombs_) return false;
  if (hap_index < counter_) reset();
  while (counter_ < hap_index) next();
  last_changed_ = -1;
  return true;
}

std::string Haplotype::get_seq() const {
  std::string s;
  for (int i = 0; i < num_blocks(); ++i) s += get_seq(i);
  return s;
}

// ---------------------------------------------------------------------------------------------------
// read_pooler.cpp:3-20: reads with identical sequence share a pool; the first read's coordinates and
// CIGAR represent the pool, qualities become the per-position median.
int32_t ReadPooler::add_alignment(const Alignment& aln) {
  if (pooled_) return -1;
  std::map<std::string, int32_t>::iterator it = seq_to_pool_.find(aln.get_sequence());
  if (it == seq_to_pool_.end()) {
    const int32_t idx = (int32_t)pooled_alns_.size();
    seq_to_pool_[aln.get_sequence()] = idx;
    pooled_alns_.push_back(Alignment(aln.get_start(), aln.get_stop(), false, aln.get_deleted(), "READPOOL", "",
                                     aln.get_sequence(), aln.get_alignment()));
    pooled_alns_.back().set_cigar_list(aln.get_cigar_list());
    qualities_by_pool_.push_back(std::vector<std::string>(1, aln.get_base_qualities()));
    return idx;
  }
  qualities_by_pool_[it->second].push_back(aln.get_base_qualities());
  return it->second;
}

bool ReadPooler::pool(const BaseQuality& base_quality) {
  for (size_t i = 0; i < pooled_alns_.size(); ++i) {
    std::vector<const std::string*> ptrs;
    for (size_t j = 0; j < qualities_by_pool_[i].size(); ++j) ptrs.push_back(&qualities_by_pool_[i][j]);
    const std::string med = base_quality.median_base_qualities(ptrs);
    if (med.size() != pooled_alns_[i].get_sequence().size()) return false;
    pooled_alns_[i].set_base_qualities(med);
  }
  pooled_ = true;
  return true;
}

// ---------------------------------------------------------------------------------------------------
// mathops.cpp:14-22: the reference tabulates log(i) for i < 1e6 with INT_LOGS[0] = -1000
double int_log(int val) { return val == 0 ? -1000.0 : log((double)val); }

double log_sum_exp(double log_v1, double log_v2) {  // mathops.cpp:53-58
  if (log_v1 > log_v2) return log_v1 + log(1 + exp(log_v2 - log_v1));
  return log_v2 + log(1 + exp(log_v1 - log_v2));
}

namespace {
// fastonebigheader.h:188-204 (Mineiro's fastpow2/fastexp) and :320-337 (fastlog2/fastlog), single precision
inline float fastpow2(float p) {
  const float offset = (p < 0) ? 1.0f : 0.0f;
  const float clipp = (p < -126) ? -126.0f : p;
  const int w = (int)clipp;
  const float z = clipp - w + offset;
  union { uint32_t i; float f; } v;
  v.i = (uint32_t)((1 << 23) * (clipp + 121.2740575f + 27.7280233f / (4.84252568f - z) - 1.49012907f * z));
  return v.f;
}
inline float fastexp(float p) { return fastpow2(1.442695040f * p); }
inline float fastlog2(float x) {
  union { float f; uint32_t i; } vx;
  vx.f = x;
  union { uint32_t i; float f; } mx;
  mx.i = (vx.i & 0x007FFFFF) | 0x3f000000;
  float y = (float)vx.i;
  y *= 1.1920928955078125e-7f;
  return y - 124.22551499f - 1.498030302f * mx.f - 1.72587999f / (0.3520887068f + mx.f);
}
inline float fastlog(float x) { return 0.69314718f * fastlog2(x); }
}  // namespace

double fast_log_sum_exp(double log_v1, double log_v2) {  // mathops.cpp:87-96
  static const double LOG_THRESH = log(0.001);
  if (log_v1 > log_v2) {
    const double diff = log_v2 - log_v1;
    return diff < LOG_THRESH ? log_v1 : log_v1 + fastlog(1 + fastexp(diff));
  }
  const double diff = log_v1 - log_v2;
  return diff < LOG_THRESH ? log_v2 : log_v2 + fastlog(1 + fastexp(diff));
}

void update_streaming_log_sum_exp(double log_val, double& max_val, double& total) {  // mathops.cpp:73-81
  if (log_val <= max_val)
    total += exp(log_val - max_val);
  else {
    total *= exp(max_val - log_val);
    total += 1.0;
    max_val = log_val;
  }
}

double finish_streaming_log_sum_exp(double max_val, double total) { return max_val + log(total); }

}  // namespace ltr
