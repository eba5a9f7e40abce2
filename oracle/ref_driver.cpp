// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product.
//
// Thin extern "C" driver around the UNMODIFIED reference translation units that
// oracle/build_ref.sh compiles in place from /root/reference/src (outputs only in
// oracle/_ref/).  It feeds a flat locus (include/longtr_b200_locus.h) through the
// reference's own public API:
//   Haplotype / HapBlock / RepeatBlock  -> HapAligner::process_reads
//                                          (src/SeqAlignment/HapAligner.h:94-138)
//   Genotyper::calc_log_sample_posteriors  (src/genotyper.cpp:45-83, protected:
//                                           reached through a derived probe class)
// and reports the time spent inside process_reads only (the reference's
// "Haplotype alignment" stage minus object construction).
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "SeqAlignment/AlignmentData.h"
#include "SeqAlignment/HapAligner.h"
#include "SeqAlignment/HapBlock.h"
#include "SeqAlignment/Haplotype.h"
#include "SeqAlignment/RepeatBlock.h"
#include "base_quality.h"
#include "genotyper.h"
#include "mathops.h"
#include "stutter_model.h"

#include "longtr_b200_locus.h"

namespace {

bool g_logs_ready = false;
void ensure_tables() {
  if (!g_logs_ready) {
    precompute_integer_logs();
    g_logs_ready = true;
  }
}

void parse_cigar(const char* cigar, Alignment& aln) {
  int num = 0;
  for (const char* p = cigar; *p; ++p) {
    if (*p >= '0' && *p <= '9')
      num = num * 10 + (*p - '0');
    else {
      aln.add_cigar_element(CigarElement(*p, num));
      num = 0;
    }
  }
}

class PosteriorProbe : public Genotyper {
 public:
  PosteriorProbe(bool haploid, const std::vector<std::string>& names,
                 const std::vector<std::vector<double> >& p1,
                 const std::vector<std::vector<double> >& p2, int num_alleles)
      : Genotyper(haploid, names, p1, p2) {
    num_alleles_           = num_alleles;
    log_sample_posteriors_ = new double[num_samples_ * num_alleles * num_alleles];
    log_aln_probs_         = new double[num_reads_ * num_alleles];
  }
  double run(const double* ll_in, double* ll_out, double* post, double* totals,
             int32_t* best_pairs) {
    std::memcpy(log_aln_probs_, ll_in, sizeof(double) * num_reads_ * num_alleles_);
    double total = calc_log_sample_posteriors();
    std::memcpy(ll_out, log_aln_probs_, sizeof(double) * num_reads_ * num_alleles_);
    std::memcpy(post, log_sample_posteriors_,
                sizeof(double) * num_samples_ * num_alleles_ * num_alleles_);
    std::memcpy(totals, sample_total_LLs_, sizeof(double) * num_samples_);
    if (best_pairs != NULL) {
      std::vector<std::pair<int, int> > gts;
      get_optimal_haplotypes(gts);
      for (int s = 0; s < num_samples_; ++s) {
        best_pairs[2 * s]     = gts[s].first;
        best_pairs[2 * s + 1] = gts[s].second;
      }
    }
    return total;
  }
};

}  // namespace

extern "C" {

// Returns 0 on success. out_ll is [n_reads * n_alleles] (caller pre-fills; slots
// of non-realigned reads / haplotypes stay untouched, as in the reference),
// out_seeds is [n_reads]. *seconds (optional) accumulates the wall time spent
// inside HapAligner::process_reads.
int ltr_ref_process_reads(const ltr_flat_locus* L, double* out_ll, int32_t* out_seeds,
                          double* seconds) {
  ensure_tables();
  StutterModel model(L->stutter[0], L->stutter[1], L->stutter[2], L->stutter[3],
                     L->stutter[4], L->stutter[5], std::string(L->motif));
  model.set_period(L->period);
  std::string lflank(L->lflank), rflank(L->rflank);
  std::vector<HapBlock*> blocks;
  blocks.push_back(new HapBlock(L->repeat_start - (int32_t)lflank.size(), L->repeat_start, lflank));
  RepeatBlock* rep = new RepeatBlock(L->repeat_start, L->repeat_end, std::string(L->alleles[0]),
                                     L->period, &model);
  for (int a = 1; a < L->n_alleles; ++a)
    rep->add_alternate(std::pair<std::string, bool>(std::string(L->alleles[a]), false));
  blocks.push_back(rep);
  blocks.push_back(new HapBlock(L->repeat_end, L->repeat_end + (int32_t)rflank.size(), rflank));
  Haplotype* hap = new Haplotype(blocks);

  std::vector<Alignment> alns;
  for (int r = 0; r < L->n_reads; ++r) {
    const ltr_flat_read& fr = L->reads[r];
    Alignment aln(fr.start, fr.stop, false, false, "read", std::string(fr.qual),
                  std::string(fr.seq), std::string(fr.seq));
    parse_cigar(fr.cigar, aln);
    alns.push_back(aln);
  }
  std::vector<bool> realign_hap(L->n_alleles, true), realign_read(L->n_reads, true);
  if (L->realign_to_hap != NULL)
    for (int a = 0; a < L->n_alleles; ++a) realign_hap[a] = L->realign_to_hap[a] != 0;
  if (L->realign_read != NULL)
    for (int r = 0; r < L->n_reads; ++r) realign_read[r] = L->realign_read[r] != 0;
  std::vector<float> params(L->aln_params, L->aln_params + L->n_aln_params);

  BaseQuality bq;
  {
    HapAligner aligner(hap, realign_hap, L->indel_flank_len, L->switch_old_align_len, params);
    std::vector<int> seeds(L->n_reads, 0);
    for (int r = 0; r < L->n_reads; ++r) seeds[r] = out_seeds[r];
    auto t0 = std::chrono::steady_clock::now();
    aligner.process_reads(alns, 0, &bq, realign_read, out_ll, seeds.data());
    auto t1 = std::chrono::steady_clock::now();
    if (seconds != NULL) *seconds += std::chrono::duration<double>(t1 - t0).count();
    for (int r = 0; r < L->n_reads; ++r) out_seeds[r] = seeds[r];
  }
  delete hap;
  for (size_t i = 0; i < blocks.size(); ++i) delete blocks[i];
  return 0;
}

// reads are sample-major: sample s owns reads_per_sample[s] consecutive rows of
// ll_in / log_p1 / log_p2.  ll_out receives the (clamped in place) copy of LL,
// post is [S*H*H], totals [S], best_pairs (optional) [2*S] from
// Genotyper::get_optimal_haplotypes (src/genotyper.cpp:85-100).
double ltr_ref_log_sample_posteriors(int haploid, int n_samples, const int32_t* reads_per_sample,
                                     int n_alleles, const double* ll_in, const double* log_p1,
                                     const double* log_p2, double* ll_out, double* post,
                                     double* totals, int32_t* best_pairs) {
  ensure_tables();
  std::vector<std::string> names;
  std::vector<std::vector<double> > p1(n_samples), p2(n_samples);
  int idx = 0;
  for (int s = 0; s < n_samples; ++s) {
    names.push_back("S" + std::to_string(s));
    for (int r = 0; r < reads_per_sample[s]; ++r, ++idx) {
      p1[s].push_back(log_p1[idx]);
      p2[s].push_back(log_p2[idx]);
    }
  }
  PosteriorProbe probe(haploid != 0, names, p1, p2, n_alleles);
  return probe.run(ll_in, ll_out, post, totals, best_pairs);
}

const char* ltr_ref_version(void) { return "LongTR reference sources, compiled in place (oracle/_ref)"; }

}  // extern "C"
