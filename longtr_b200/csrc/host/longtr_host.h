// longtr_host.h -- C++ host side of the drop-in boundary.
//
// Mirrors, for the hot path only, the classes LongTR's callers use -- same class and method names,
// same argument meaning and error behaviour -- so that code written against the reference's
// interfaces (reference: src/SeqAlignment/{AlignmentData,HapBlock,RepeatBlock,Haplotype,HapAligner}.h,
// src/{base_quality,stutter_model,read_pooler,genotyper}.h) reads the same here.  All arithmetic of the
// path runs on the GPU through the C ABI (include/longtr_b200.h); what stays on the host is integer /
// string work: CIGAR trimming, seed selection, haplotype enumeration, flattening, genotype extraction.
// Everything lives in namespace ltr to stay link-compatible with a process that also holds the reference.
#pragma once
#include <stdint.h>

#include <map>
#include <string>
#include <utility>
#include <vector>

#include "longtr_b200.h"

namespace ltr {

// ---- reference: src/SeqAlignment/AlignmentData.h:12-29 -------------------------------------------
class CigarElement {
 public:
  CigarElement(char type, int num) : type_(type), num_(num) {}
  void set_type(char type) { type_ = type; }
  void set_num(int num) { num_ = num; }
  char get_type() const { return type_; }
  int get_num() const { return num_; }

 private:
  char type_;
  int num_;
};

// ---- reference: src/SeqAlignment/AlignmentData.h:32-140 ------------------------------------------
class Alignment {
 public:
  Alignment(int32_t start, int32_t stop, bool rev_strand, bool deleted, const std::string& name,
            const std::string& base_qualities, const std::string& sequence, const std::string& alignment)
      : start_(start), stop_(stop), rev_strand_(rev_strand), deleted_(deleted), name_(name),
        base_qualities_(base_qualities), sequence_(sequence), alignment_(alignment) {}
  const std::string& get_name() const { return name_; }
  int32_t get_start() const { return start_; }
  int32_t get_stop() const { return stop_; }
  bool get_deleted() const { return deleted_; }
  bool is_from_reverse_strand() const { return rev_strand_; }
  void set_base_qualities(const std::string& q) { base_qualities_ = q; }
  void add_cigar_element(CigarElement e) { cigar_list_.push_back(e); }
  void set_cigar_list(const std::vector<CigarElement>& c) { cigar_list_ = c; }
  const std::string& get_base_qualities() const { return base_qualities_; }
  const std::string& get_sequence() const { return sequence_; }
  const std::string& get_alignment() const { return alignment_; }
  const std::vector<CigarElement>& get_cigar_list() const { return cigar_list_; }
  std::string getCigarString() const;
  // "110=4I90=" -> cigar list; returns false on a malformed string
  bool set_cigar_string(const char* cigar);

 private:
  int32_t start_, stop_;
  bool rev_strand_, deleted_;
  std::string name_, base_qualities_, sequence_, alignment_;
  std::vector<CigarElement> cigar_list_;
};

// ---- reference: src/base_quality.h:14-89, base_quality.cpp:11-28 ---------------------------------
class BaseQuality {
 public:
  static const char MIN_BASE_QUALITY = '!';
  static const char MAX_BASE_QUALITY = 'J';
  BaseQuality();
  double log_prob_error(char quality) const { return log_error_[index(quality)]; }
  double log_prob_correct(char quality) const { return log_correct_[index(quality)]; }
  // per-position median over the reads of a pool (all strings must have the same length; returns "" otherwise)
  std::string median_base_qualities(const std::vector<const std::string*>& qualities) const;
  // 256-entry tables indexed by the raw quality byte (what the GPU consumes)
  void byte_tables(double* log_correct256, double* log_error256) const;

 private:
  static int index(char quality) {
    if (quality < MIN_BASE_QUALITY) return 0;
    if (quality > MAX_BASE_QUALITY) return MAX_BASE_QUALITY - MIN_BASE_QUALITY;
    return quality - MIN_BASE_QUALITY;
  }
  double log_correct_[256], log_error_[256];
};

// ---- reference: src/stutter_model.h:16-88, stutter_model.cpp:29-53 --------------------------------
class StutterModel {
 public:
  StutterModel(double inframe_geom, double inframe_up, double inframe_down, double outframe_geom,
               double outframe_up, double outframe_down, const std::string& motif);
  double log_stutter_pmf(int sample_bps, int read_bps) const;
  int period() const { return motif_len_; }
  void set_period(int period) { motif_len_ = period; }
  const std::string& motif() const { return motif_; }
  // the reference's copy() re-derives the period from the motif string (stutter_model.h:72)
  StutterModel copy() const {
    return StutterModel(in_geom_, in_up_, in_down_, out_geom_, out_up_, out_down_, motif_);
  }
  bool valid() const { return valid_; }
  void get_parameters(double out[6]) const {  // constructor order
    out[0] = in_geom_; out[1] = in_up_; out[2] = in_down_; out[3] = out_geom_; out[4] = out_up_; out[5] = out_down_;
  }

 private:
  double in_geom_, in_up_, in_down_, out_geom_, out_up_, out_down_;
  double in_log_nostep_, in_log_step_, in_log_up_, in_log_down_, log_equal_;
  double out_log_nostep_, out_log_step_, out_log_up_, out_log_down_;
  int motif_len_;
  std::string motif_;
  bool valid_;
};

// ---- reference: src/SeqAlignment/RepeatStutterInfo.h:9-62 -----------------------------------------
const int MAX_STUTTER_REPEAT_INS = 6;
const int MAX_STUTTER_REPEAT_DEL = -6;
class RepeatStutterInfo {
 public:
  RepeatStutterInfo(int period, const std::string& ref_allele, const StutterModel& model)
      : period_(period), max_ins_(MAX_STUTTER_REPEAT_INS * period), max_del_(MAX_STUTTER_REPEAT_DEL * period),
        model_(model.copy()) {
    allele_sizes_.push_back((int)ref_allele.size());
  }
  int get_period() const { return period_; }
  int max_insertion() const { return max_ins_; }
  int max_deletion() const { return max_del_; }
  const StutterModel& get_stutter_model() const { return model_; }
  void add_alternate_allele(const std::string& alt) { allele_sizes_.push_back((int)alt.size()); }
  double log_prob_pcr_artifact(int seq_index, int artifact_size) const;

 private:
  int period_, max_ins_, max_del_;
  StutterModel model_;
  std::vector<int> allele_sizes_;
};

// ---- reference: src/SeqAlignment/HapBlock.h:18-160 (a block with >1 option carries a RepeatStutterInfo:
//      RepeatBlock.h:15-70) ---------------------------------------------------------------------------
class HapBlock {
 public:
  HapBlock(int32_t start, int32_t end, const std::string& ref_seq) : start_(start), end_(end) {
    seqs_.push_back(ref_seq);
  }
  virtual ~HapBlock() {}
  virtual const RepeatStutterInfo* get_repeat_info() const { return NULL; }
  virtual void add_alternate(const std::pair<std::string, bool>& alt_seq) {
    seqs_.push_back(alt_seq.first);
    inexact_.push_back(alt_seq.second);
  }
  int32_t start() const { return start_; }
  int32_t end() const { return end_; }
  int num_options() const { return (int)seqs_.size(); }
  int size(int index) const { return (int)seqs_.at(index).size(); }
  int max_size() const;
  const std::string& get_seq(unsigned int index) const { return seqs_.at(index); }

 protected:
  int32_t start_, end_;
  std::vector<std::string> seqs_;
  std::vector<bool> inexact_;
};

class RepeatBlock : public HapBlock {
 public:
  RepeatBlock(int32_t start, int32_t end, const std::string& ref_seq, int period, const StutterModel* stutter_model)
      : HapBlock(start, end, ref_seq), info_(period, ref_seq, *stutter_model) {}
  const RepeatStutterInfo* get_repeat_info() const { return &info_; }
  void add_alternate(const std::pair<std::string, bool>& alt) {
    HapBlock::add_alternate(alt);
    info_.add_alternate_allele(alt.first);
  }

 private:
  RepeatStutterInfo info_;
};

// ---- reference: src/SeqAlignment/Haplotype.h:12-130, Haplotype.cpp:123-206 (gray-code iterator) ----
class Haplotype {
 public:
  explicit Haplotype(std::vector<HapBlock*>& blocks);
  const std::string& get_seq(int block_index) const { return blocks_[block_index]->get_seq(counts_[block_index]); }
  std::string get_seq() const;
  char get_first_char() const { return get_seq(0)[0]; }
  char get_last_char() const { const std::string& s = get_seq((int)blocks_.size() - 1); return s[s.size() - 1]; }
  HapBlock* get_block(int block_index) const { return blocks_[block_index]; }
  HapBlock* get_first_block() const { return blocks_.front(); }
  HapBlock* get_last_block() const { return blocks_.back(); }
  int num_blocks() const { return (int)blocks_.size(); }
  int num_combs() const { return ncombs_; }
  int last_changed() const { return last_changed_; }
  int max_size() const { return max_size_; }
  int cur_size() const { return cur_size_; }
  int cur_index() const { return counter_; }
  int cur_index(int block_index) const { return counts_[block_index]; }
  void fix() { fixed_ = true; }
  void unfix() { fixed_ = false; }
  void reset();
  bool next();
  bool go_to(int hap_index);  // false (instead of exit(1)) for an invalid index

 private:
  void init();
  std::vector<HapBlock*> blocks_;
  std::vector<int> nopts_, dirs_, factors_, counts_;
  int ncombs_, cur_size_, counter_, last_changed_, max_size_;
  bool fixed_;
};

// ---- reference: src/SeqAlignment/HapAligner.h:12-37 --------------------------------------------------
struct AlignmentModel {
  unsigned int MAX_HOMOP_LEN;
  float LOG_INS_TO_INS, LOG_INS_TO_MATCH, LOG_DEL_TO_DEL, LOG_DEL_TO_MATCH, LOG_MATCH_TO_MATCH, LOG_MATCH_TO_INS,
      LOG_MATCH_TO_DEL;
};

// ---- reference: src/SeqAlignment/HapAligner.h:39-147 -------------------------------------------------
// Same constructor and process_reads signature as the reference plus the GPU context the work runs on.
// status() reports the last C-ABI error instead of the reference's exit(1).
class HapAligner {
 public:
  HapAligner(Haplotype* haplotype, std::vector<bool>& realign_to_haplotype, int INDEL_FLANK_LEN,
             int SWITCH_OLD_ALIGN_LEN, std::vector<float>& alignment_model_params, ltr_ctx* ctx);
  // 0-based index of the seed base in the read, or -1 (HapAligner.cpp:493-542)
  int calc_seed_base(const Alignment& alignment) const;
  // read bases aligned to [repeat_start - INDEL_FLANK_LEN, repeat_end + INDEL_FLANK_LEN) (HapAligner.cpp:346-465),
  // false for a CIGAR with an unknown operation
  bool trim_alignment(const Alignment& aln, std::string& trimmed_seq) const;
  // aln_probs[(init_read_index+i)*num_combs + hap], seed_positions[init_read_index+i] (HapAligner.cpp:545-581)
  void process_reads(const std::vector<Alignment>& alignments, int init_read_index, const BaseQuality* base_quality,
                     const std::vector<bool>& realign_read, double* aln_probs, int* seed_positions);
  int status() const { return status_; }
  const AlignmentModel& model() const { return model_; }

  // The long path in two halves, so that many loci can share ONE GPU job (ltr_process_reads_flat_batch):
  // prepare_long appends the locus' haplotypes (column order, only the ones flagged for realignment) and trimmed
  // reads to the flattened strings and sets seed_positions like process_reads; scatter_long writes the job's
  // P x H log-likelihood matrix of the locus into aln_probs.  uses_short_path(): the locus takes the homopolymer
  // path instead (HapAligner.cpp:552).
  struct LongPart {
    std::vector<int> hap_cols, read_rows;
  };
  bool uses_short_path() const;
  bool prepare_long(const std::vector<Alignment>& alns, int init_read_index, const std::vector<bool>& realign_read,
                    int* seed_positions, LongPart& part, std::string& hap_bytes, std::vector<uint32_t>& hap_off,
                    std::string& read_bytes, std::vector<uint32_t>& read_off);
  void scatter_long(const LongPart& part, const double* ll, int init_read_index, double* aln_probs) const;
  void fill_params(ltr_params& p) const;

 private:
  void calc_best_seed_position(int32_t region_start, int32_t region_end, int32_t& best_dist, int32_t& best_pos) const;
  void process_reads_long(const std::vector<Alignment>& alns, int init_read_index, const std::vector<bool>& realign_read,
                          double* aln_probs, int* seed_positions);
  void process_reads_short(const std::vector<Alignment>& alns, int init_read_index, const BaseQuality* bq,
                           const std::vector<bool>& realign_read, double* aln_probs, int* seed_positions);
  Haplotype* fw_haplotype_;
  std::vector<bool> realign_to_hap_;
  std::vector<int32_t> repeat_starts_, repeat_ends_;
  int INDEL_FLANK_LEN_, SWITCH_OLD_ALIGN_LEN_;
  AlignmentModel model_;
  ltr_ctx* ctx_;
  int status_;
};

// ---- reference: src/read_pooler.h:12-53, read_pooler.cpp:3-20 ----------------------------------------
class ReadPooler {
 public:
  ReadPooler() : pooled_(false) {}
  int32_t num_pools() const { return (int32_t)pooled_alns_.size(); }
  int32_t add_alignment(const Alignment& aln);  // -1 once pool() has run
  bool pool(const BaseQuality& base_quality);
  std::vector<Alignment>& get_alignments() { return pooled_alns_; }

 private:
  std::vector<Alignment> pooled_alns_;
  std::vector<std::vector<std::string> > qualities_by_pool_;
  std::map<std::string, int32_t> seq_to_pool_;
  bool pooled_;
};

// ---- reference: src/genotyper.h:14-161, genotyper.cpp:21-256 -----------------------------------------
class Genotyper {
 public:
  Genotyper(bool haploid, const std::vector<std::string>& sample_names,
            const std::vector<std::vector<double> >& log_p1, const std::vector<std::vector<double> >& log_p2,
            ltr_ctx* ctx);
  virtual ~Genotyper() {}
  void calc_PLs(const std::vector<double>& gls, std::vector<int>& pls) const;
  double calc_gl_diff(const std::vector<double>& gls, int gt_a, int gt_b) const;
  void extract_genotypes_and_likelihoods(
      int num_variants, std::vector<int>& hap_to_allele, std::vector<std::pair<int, int> >& best_haplotypes,
      std::vector<std::pair<int, int> >& best_gts, std::vector<double>& log_phased_posteriors,
      std::vector<double>& log_unphased_posteriors, std::vector<double>& hap_log_phased_posteriors,
      std::vector<double>& hap_log_unphased_posteriors, bool calc_gls, std::vector<std::vector<double> >& gls,
      std::vector<double>& gl_diffs, bool calc_pls, std::vector<std::vector<int> >& pls, bool calc_phased_gls,
      std::vector<std::vector<double> >& phased_gls);
  // the reference keeps these protected and reaches them from SeqStutterGenotyper; public here for the flat API
  void set_num_alleles(int num_alleles);
  double* log_aln_probs() { return log_aln_probs_.data(); }
  const double* log_sample_posteriors() const { return log_sample_posteriors_.data(); }
  const double* sample_total_LLs() const { return sample_total_LLs_.data(); }
  void load_posteriors(const double* post, const double* totals) {
    log_sample_posteriors_.assign(post, post + (size_t)num_samples_ * num_alleles_ * num_alleles_);
    sample_total_LLs_.assign(totals, totals + num_samples_);
  }
  double calc_log_sample_posteriors();  // GPU (ltr_posteriors); clamps log_aln_probs in place like genotyper.cpp:57-58
  void get_optimal_haplotypes(std::vector<std::pair<int, int> >& gts) const;
  int num_reads() const { return (int)num_reads_; }
  int num_samples() const { return num_samples_; }
  int status() const { return status_; }

 protected:
  double log_homozygous_prior() const;
  double log_heterozygous_prior() const;
  unsigned int num_reads_;
  int num_samples_, num_alleles_;
  bool haploid_;
  std::vector<double> log_p1_, log_p2_;
  std::vector<int32_t> sample_label_;
  std::vector<std::string> sample_names_;
  std::vector<double> log_sample_posteriors_, log_aln_probs_, sample_total_LLs_;
  ltr_ctx* ctx_;
  int status_;
};

// mathops (reference src/mathops.cpp:14-107, src/fastonebigheader.h:188-218, 320-357)
double int_log(int val);
double log_sum_exp(double log_v1, double log_v2);
double fast_log_sum_exp(double log_v1, double log_v2);
void update_streaming_log_sum_exp(double log_val, double& max_val, double& total);
double finish_streaming_log_sum_exp(double max_val, double total);

}  // namespace ltr
