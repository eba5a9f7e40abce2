// reference_binding.cpp -- the LongTR-side binding of liblongtr_b200.so.
//
// This file is compiled against LongTR's OWN headers and replaces the bodies of exactly two member functions,
// keeping their signatures, so that everything else in LongTR (SeqStutterGenotyper, HaplotypeGenerator,
// VCF writing, ...) runs unmodified on top of the GPU path:
//   HapAligner::process_reads               reference src/SeqAlignment/HapAligner.cpp:545-581
//   Genotyper::calc_log_sample_posteriors   reference src/genotyper.cpp:45-83
// Both flatten their inputs into the plain C structs of include/longtr_b200*.h and call the C ABI.  There is no
// CPU fallback: without a B200 the process exits through LongTR's own printErrorAndDie.
// tests/test_gpu_dropin.py links this file into the reference's per-locus genotyper (oracle/build_ref.sh,
// libltr_ref_gpu.so) and checks that the VCF records it writes are identical to the all-CPU reference's.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "SeqAlignment/HapAligner.h"
#include "SeqAlignment/RepeatBlock.h"
#include "error.h"
#include "genotyper.h"
#include "stutter_model.h"

#include "longtr_b200.h"

// LONGTR_B200_TIMING=1: time spent inside the two replaced functions, printed when the process exits.
namespace {
struct BindingTimer {
  double process_reads_s, posteriors_s;
  long process_reads_n, posteriors_n;
  bool on;
  BindingTimer() : process_reads_s(0), posteriors_s(0), process_reads_n(0), posteriors_n(0),
                   on(std::getenv("LONGTR_B200_TIMING") != NULL) {}
  ~BindingTimer() {
    if (on)
      std::fprintf(stderr, "[longtr_b200 binding] process_reads: %ld calls, %.3f ms each; calc_log_sample_posteriors: %ld calls, "
                           "%.3f ms each\n", process_reads_n, 1e3 * process_reads_s / (process_reads_n ? process_reads_n : 1),
                   posteriors_n, 1e3 * posteriors_s / (posteriors_n ? posteriors_n : 1));
  }
};
BindingTimer g_timer;
struct Scope {
  double* acc; long* n;
  std::chrono::steady_clock::time_point t0;
  Scope(double* a, long* c) : acc(a), n(c), t0(std::chrono::steady_clock::now()) {}
  ~Scope() { *acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); ++*n; }
};
}  // namespace

static ltr_ctx* longtr_b200_ctx() {
  static ltr_ctx* ctx = NULL;  // LongTR is single-threaded: one context per process
  if (ctx == NULL) {
    const int rc = ltr_ctx_create(0, &ctx);
    if (rc != LTR_OK) printErrorAndDie(std::string("longtr_b200: ") + ltr_strerror(rc));
  }
  return ctx;
}

void HapAligner::process_reads(const std::vector<Alignment>& alignments, int init_read_index,
                               const BaseQuality* base_quality, const std::vector<bool>& realign_read,
                               double* aln_probs, int* seed_positions) {
  assert(alignments.size() == realign_read.size());
  ltr_ctx* ctx_first = longtr_b200_ctx();  // (context creation stays outside the timer)
  (void)ctx_first;
  Scope timer(&g_timer.process_reads_s, &g_timer.process_reads_n);
  if (fw_haplotype_->num_blocks() != 3) printErrorAndDie("longtr_b200: expected flank / repeat / flank haplotype blocks");
  HapBlock* left = fw_haplotype_->get_block(0);
  HapBlock* rep = fw_haplotype_->get_block(1);
  HapBlock* right = fw_haplotype_->get_block(2);
  RepeatStutterInfo* info = rep->get_repeat_info();
  if (info == NULL || left->num_options() != 1 || right->num_options() != 1)
    printErrorAndDie("longtr_b200: expected a single multi-allele repeat block");

  ltr_flat_locus L;
  std::vector<const char*> alleles;
  for (int a = 0; a < rep->num_options(); ++a) alleles.push_back(rep->get_seq(a).c_str());
  L.lflank = left->get_seq(0).c_str();
  L.rflank = right->get_seq(0).c_str();
  L.repeat_start = rep->start();
  L.repeat_end = rep->end();
  L.period = info->get_period();
  L.n_alleles = (int32_t)alleles.size();
  L.alleles = alleles.data();
  StutterModel* sm = info->get_stutter_model();
  L.stutter[0] = sm->get_parameter(true, 'P');
  L.stutter[1] = sm->get_parameter(true, 'U');
  L.stutter[2] = sm->get_parameter(true, 'D');
  L.stutter[3] = sm->get_parameter(false, 'P');
  L.stutter[4] = sm->get_parameter(false, 'U');
  L.stutter[5] = sm->get_parameter(false, 'D');
  const std::string motif = sm->motif();
  L.motif = motif.c_str();
  std::vector<ltr_flat_read> reads(alignments.size());
  std::vector<std::string> cigars(alignments.size());
  for (size_t i = 0; i < alignments.size(); ++i) {
    cigars[i] = alignments[i].getCigarString();
    reads[i].start = alignments[i].get_start();
    reads[i].stop = alignments[i].get_stop();
    reads[i].seq = alignments[i].get_sequence().c_str();
    reads[i].qual = alignments[i].get_base_qualities().c_str();
    reads[i].cigar = cigars[i].c_str();
  }
  L.n_reads = (int32_t)reads.size();
  L.reads = reads.data();
  L.indel_flank_len = INDEL_FLANK_LEN;
  L.switch_old_align_len = SWITCH_OLD_ALIGN_LEN;
  L.n_aln_params = 7;
  L.aln_params[0] = AlnModel->LOG_INS_TO_INS;
  L.aln_params[1] = AlnModel->LOG_INS_TO_MATCH;
  L.aln_params[2] = AlnModel->LOG_DEL_TO_DEL;
  L.aln_params[3] = AlnModel->LOG_DEL_TO_MATCH;
  L.aln_params[4] = AlnModel->LOG_MATCH_TO_MATCH;
  L.aln_params[5] = AlnModel->LOG_MATCH_TO_INS;
  L.aln_params[6] = AlnModel->LOG_MATCH_TO_DEL;
  std::vector<uint8_t> hap_mask(realign_to_hap_.size()), read_mask(realign_read.size());
  for (size_t a = 0; a < hap_mask.size(); ++a) hap_mask[a] = realign_to_hap_[a] ? 1 : 0;
  for (size_t i = 0; i < read_mask.size(); ++i) read_mask[i] = realign_read[i] ? 1 : 0;
  L.realign_to_hap = hap_mask.data();
  L.realign_read = read_mask.data();
  (void)base_quality;  // the library builds the same BaseQuality tables (src/base_quality.h:29-38)

  std::vector<int32_t> seeds(alignments.size());
  for (size_t i = 0; i < seeds.size(); ++i) seeds[i] = seed_positions[init_read_index + i];
  const int rc = ltr_process_reads_flat(longtr_b200_ctx(), &L, aln_probs + init_read_index * fw_haplotype_->num_combs(),
                                        seeds.data());
  if (rc != LTR_OK) printErrorAndDie(std::string("longtr_b200: ") + ltr_strerror(rc));
  for (size_t i = 0; i < seeds.size(); ++i) seed_positions[init_read_index + i] = seeds[i];
}

double Genotyper::calc_log_sample_posteriors(std::vector<int>& read_weights) {
  assert(read_weights.size() == num_reads_);  // accepted but unused, as in the reference (genotyper.cpp:45-83)
  ltr_ctx* ctx_first = longtr_b200_ctx();
  (void)ctx_first;
  Scope timer(&g_timer.posteriors_s, &g_timer.posteriors_n);
  double total_LL = 0.0;
  const int rc = ltr_posteriors(longtr_b200_ctx(), haploid_ ? 1 : 0, num_samples_, (int32_t)num_reads_, num_alleles_,
                                log_aln_probs_ /* clamped in place like :57-58 */, log_p1_, log_p2_, sample_label_,
                                log_sample_posteriors_, sample_total_LLs_, &total_LL);
  if (rc != LTR_OK) printErrorAndDie(std::string("longtr_b200: ") + ltr_strerror(rc));
  return total_LL;
}
