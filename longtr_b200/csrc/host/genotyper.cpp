// genotyper.cpp -- host mirror of LongTR's Genotyper base class (reference src/genotyper.{h,cpp}).
//
// calc_log_sample_posteriors runs on the GPU (ltr_posteriors -> posterior_kernel); what remains on the host
// is the integer / small-vector work that turns the S x H x H posterior array into calls:
// get_optimal_haplotypes (genotyper.cpp:85-100), extract_genotypes_and_likelihoods (:132-256),
// calc_PLs (:102-107), calc_gl_diff (:109-130).  Arithmetic follows the reference operation by operation
// (streaming log-sum-exp, the approximate two-argument fast_log_sum_exp, LOG_E_BASE_10 = 0.4342944819).
#include <float.h>
#include <math.h>
#include <string.h>

#include <algorithm>

#include "longtr_host.h"

namespace ltr {

static const double TOLERANCE = 1e-10;             // mathops.cpp:11
static const double LOG_E_BASE_10 = 0.4342944819;  // mathops.cpp:12

Genotyper::Genotyper(bool haploid, const std::vector<std::string>& sample_names,
                     const std::vector<std::vector<double> >& log_p1, const std::vector<std::vector<double> >& log_p2,
                     ltr_ctx* ctx)
    : num_reads_(0), num_samples_((int)log_p1.size()), num_alleles_(-1), haploid_(haploid),
      sample_names_(sample_names), ctx_(ctx), status_(LTR_OK) {
  if (log_p1.size() != log_p2.size() || log_p1.size() != sample_names.size()) status_ = LTR_ERR_INVALID;
  for (size_t s = 0; s < log_p1.size() && status_ == LTR_OK; ++s) {
    if (log_p1[s].size() != log_p2[s].size()) {
      status_ = LTR_ERR_INVALID;
      break;
    }
    for (size_t r = 0; r < log_p1[s].size(); ++r) {
      if (!(log_p1[s][r] <= 0.0 && log_p2[s][r] <= 0.0)) status_ = LTR_ERR_INVALID;  // the reference asserts
      log_p1_.push_back(log_p1[s][r]);
      log_p2_.push_back(log_p2[s][r]);
      sample_label_.push_back((int32_t)s);
    }
  }
  num_reads_ = (unsigned int)log_p1_.size();
  sample_total_LLs_.assign((size_t)std::max(num_samples_, 0), 0.0);
}

void Genotyper::set_num_alleles(int num_alleles) {
  num_alleles_ = num_alleles;
  log_sample_posteriors_.assign((size_t)num_samples_ * num_alleles * num_alleles, 0.0);
  log_aln_probs_.assign((size_t)num_reads_ * num_alleles, 0.0);
}

double Genotyper::log_homozygous_prior() const {  // genotyper.cpp:21-26
  if (haploid_) return -int_log(num_alleles_);
  return int_log(2) - int_log(num_alleles_) - int_log(num_alleles_ + 1);
}

double Genotyper::log_heterozygous_prior() const {  // genotyper.cpp:28-33
  if (haploid_) return -DBL_MAX / 2;
  return -int_log(num_alleles_) - int_log(num_alleles_ + 1);
}

double Genotyper::calc_log_sample_posteriors() {
  if (status_ != LTR_OK) return 0.0;
  if (ctx_ == NULL || num_alleles_ < 1) {
    status_ = LTR_ERR_INVALID;
    return 0.0;
  }
  double total = 0.0;
  status_ = ltr_posteriors(ctx_, haploid_ ? 1 : 0, num_samples_, (int32_t)num_reads_, num_alleles_, log_aln_probs_.data(),
                           log_p1_.data(), log_p2_.data(), sample_label_.data(), log_sample_posteriors_.data(),
                           sample_total_LLs_.data(), &total);
  return total;
}

void Genotyper::get_optimal_haplotypes(std::vector<std::pair<int, int> >& gts) const {
  gts.assign((size_t)num_samples_, std::pair<int, int>(-1, -1));
  const double* p = log_sample_posteriors_.data();
  for (int s = 0; s < num_samples_; ++s) {
    double best = -DBL_MAX;
    for (int a = 0; a < num_alleles_; ++a)
      for (int b = 0; b < num_alleles_; ++b, ++p)
        if (*p > best) {  // first strict maximum in row-major order
          best = *p;
          gts[s] = std::pair<int, int>(a, b);
        }
  }
}

void Genotyper::calc_PLs(const std::vector<double>& gls, std::vector<int>& pls) const {
  const double max_gl = *std::max_element(gls.begin(), gls.end());
  for (size_t i = 0; i < gls.size(); ++i) pls.push_back(std::min(999, (int)(-10 * (gls[i] - max_gl))));
}

double Genotyper::calc_gl_diff(const std::vector<double>& gls, int gt_a, int gt_b) const {
  if (num_alleles_ == 1) return -1000;
  const double max_gl = *std::max_element(gls.begin(), gls.end());
  double second_gl = -DBL_MAX;
  for (size_t i = 0; i < gls.size(); ++i)
    if (gls[i] < max_gl) second_gl = std::max(second_gl, gls[i]);
  if (second_gl == -DBL_MAX) second_gl = max_gl;
  int gl_index;
  if (haploid_)
    gl_index = gt_a;
  else {
    const int lo = std::min(gt_a, gt_b), hi = std::max(gt_a, gt_b);
    gl_index = hi * (hi + 1) / 2 + lo;
  }
  return (fabs(max_gl - gls[gl_index]) < TOLERANCE) ? (max_gl - second_gl) : gls[gl_index] - max_gl;
}

void Genotyper::extract_genotypes_and_likelihoods(
    int num_variants, std::vector<int>& hap_to_allele, std::vector<std::pair<int, int> >& best_haplotypes,
    std::vector<std::pair<int, int> >& best_gts, std::vector<double>& log_phased_posteriors,
    std::vector<double>& log_unphased_posteriors, std::vector<double>& hap_log_phased_posteriors,
    std::vector<double>& hap_log_unphased_posteriors, bool calc_gls, std::vector<std::vector<double> >& gls,
    std::vector<double>& gl_diffs, bool calc_pls, std::vector<std::vector<int> >& pls, bool calc_phased_gls,
    std::vector<std::vector<double> >& phased_gls) {
  const int S = num_samples_, H = num_alleles_, V = num_variants;
  get_optimal_haplotypes(best_haplotypes);
  for (int s = 0; s < S; ++s)
    best_gts.push_back(std::pair<int, int>(hap_to_allele[best_haplotypes[s].first], hap_to_allele[best_haplotypes[s].second]));

  // haplotype pairs -> allele pairs, streaming log-sum-exp in storage order (genotyper.cpp:157-176)
  std::vector<std::vector<double> > run_max((size_t)S, std::vector<double>((size_t)V * V, -DBL_MAX / 2));
  std::vector<std::vector<double> > total((size_t)S, std::vector<double>((size_t)V * V, 0.0));
  const double* p = log_sample_posteriors_.data();
  for (int s = 0; s < S; ++s)
    for (int a = 0; a < H; ++a)
      for (int b = 0; b < H; ++b, ++p) {
        const int gt = V * hap_to_allele[a] + hap_to_allele[b];
        update_streaming_log_sum_exp(*p, run_max[s][gt], total[s][gt]);
      }
  for (int gt = 0; gt < V * V; ++gt)
    for (int s = 0; s < S; ++s) total[s][gt] = finish_streaming_log_sum_exp(run_max[s][gt], total[s][gt]);

  // posteriors of the optimal haplotype pair, phased and unphased (genotyper.cpp:178-190)
  p = log_sample_posteriors_.data();
  for (int s = 0; s < S; ++s, p += (size_t)H * H) {
    const int ia = best_haplotypes[s].first * H + best_haplotypes[s].second;
    const int ib = best_haplotypes[s].second * H + best_haplotypes[s].first;
    hap_log_phased_posteriors.push_back(p[ia]);
    hap_log_unphased_posteriors.push_back(ia != ib ? fast_log_sum_exp(p[ia], p[ib]) : p[ia]);
  }
  // ... and of the optimal genotype (genotyper.cpp:192-203)
  for (int s = 0; s < S; ++s) {
    const int ga = best_gts[s].first, gb = best_gts[s].second;
    const double phased = total[s][V * ga + gb];
    log_phased_posteriors.push_back(phased);
    log_unphased_posteriors.push_back(ga == gb ? phased : log_sum_exp(phased, total[s][V * gb + ga]));
  }

  if (!(calc_gls || calc_phased_gls || calc_pls)) return;
  // likelihoods = posteriors with the priors taken out again (genotyper.cpp:207-245)
  gls.assign((size_t)S, std::vector<double>());
  if (calc_phased_gls) phased_gls.assign((size_t)S, std::vector<double>());
  const double hom_corr = log_homozygous_prior();
  const double het_corr = haploid_ ? 0 : log_heterozygous_prior();
  double gl_nconfig, pgl_nconfig;
  if (haploid_) {
    gl_nconfig = int_log(2) + int_log(H) - int_log(V);
    pgl_nconfig = int_log(H) - int_log(V);
  } else {
    gl_nconfig = int_log(2) + 2 * (int_log(H) - int_log(V));
    pgl_nconfig = 2 * (int_log(H) - int_log(V));
  }
  int gt = 0;
  for (int a = 0; a < V; ++a)
    for (int b = 0; b < V; ++b, ++gt) {
      const int alt_gt = b * V + a;
      const double gl_corr = (a == b ? hom_corr : het_corr) + gl_nconfig;
      const double pgl_corr = (a == b ? hom_corr : het_corr) + pgl_nconfig;
      for (int s = 0; s < S; ++s) {
        if (b <= a && (!haploid_ || a == b)) {
          const double gl_e = sample_total_LLs_[s] - gl_corr + fast_log_sum_exp(total[s][gt], total[s][alt_gt]);
          gls[s].push_back(gl_e * LOG_E_BASE_10);
        }
        if (calc_phased_gls && (!haploid_ || a == b))
          phased_gls[s].push_back((sample_total_LLs_[s] - pgl_corr + total[s][gt]) * LOG_E_BASE_10);
      }
    }
  for (int s = 0; s < S; ++s) gl_diffs.push_back(calc_gl_diff(gls[s], best_gts[s].first, best_gts[s].second));
  if (calc_pls) {
    pls.assign((size_t)S, std::vector<int>());
    for (int s = 0; s < S; ++s) calc_PLs(gls[s], pls[s]);
  }
  if (!calc_gls) gls.clear();
}

}  // namespace ltr

// ---- C ABI: genotype calls for one locus ----------------------------------------------------------------
namespace {
int fill_calls(ltr::Genotyper& g, int haploid, int32_t n_samples, int32_t n_alleles, double total, ltr_locus_calls* out) {
  // one haplotype block with options -> haplotype index == allele index (Appendix D of SURVEY.md)
  std::vector<int> hap_to_allele((size_t)n_alleles);
  for (int a = 0; a < n_alleles; ++a) hap_to_allele[a] = a;
  std::vector<std::pair<int, int> > best_haps, best_gts;
  std::vector<double> lpp, lup, hlpp, hlup, gl_diffs;
  std::vector<std::vector<double> > gls, pgls;
  std::vector<std::vector<int> > pls;
  g.extract_genotypes_and_likelihoods(n_alleles, hap_to_allele, best_haps, best_gts, lpp, lup, hlpp, hlup, true, gls,
                                      gl_diffs, true, pls, true, pgls);
  const size_t S = (size_t)n_samples, H = (size_t)n_alleles;
  const size_t n_gl = haploid ? H : H * (H + 1) / 2, n_pgl = haploid ? H : H * H;
  out->total_ll = total;
  for (size_t s = 0; s < S; ++s) {
    if (out->best_gts) { out->best_gts[2 * s] = best_gts[s].first; out->best_gts[2 * s + 1] = best_gts[s].second; }
    if (out->log_phased_posteriors) out->log_phased_posteriors[s] = lpp[s];
    if (out->log_unphased_posteriors) out->log_unphased_posteriors[s] = lup[s];
    if (out->hap_log_phased_posteriors) out->hap_log_phased_posteriors[s] = hlpp[s];
    if (out->hap_log_unphased_posteriors) out->hap_log_unphased_posteriors[s] = hlup[s];
    if (out->gl_diffs) out->gl_diffs[s] = gl_diffs[s];
    if (out->sample_total_lls) out->sample_total_lls[s] = g.sample_total_LLs()[s];
    if (gls[s].size() != n_gl || pls[s].size() != n_gl || pgls[s].size() != n_pgl) return LTR_ERR_INVALID;
    if (out->gls) std::copy(gls[s].begin(), gls[s].end(), out->gls + s * n_gl);
    if (out->pls) std::copy(pls[s].begin(), pls[s].end(), out->pls + s * n_gl);
    if (out->phased_gls) std::copy(pgls[s].begin(), pgls[s].end(), out->phased_gls + s * n_pgl);
  }
  if (out->log_sample_posteriors)
    std::copy(g.log_sample_posteriors(), g.log_sample_posteriors() + S * H * H, out->log_sample_posteriors);
  return LTR_OK;
}
}  // namespace

extern "C" int ltr_extract_calls(int haploid, int32_t n_samples, int32_t n_alleles, const double* post,
                                 const double* totals, ltr_locus_calls* out) {
  using namespace ltr;
  if (!post || !totals || !out || n_samples < 1 || n_alleles < 1) return LTR_ERR_INVALID;
  std::vector<std::string> names((size_t)n_samples, "S");
  std::vector<std::vector<double> > none((size_t)n_samples);
  Genotyper g(haploid != 0, names, none, none, NULL);
  g.set_num_alleles(n_alleles);
  g.load_posteriors(post, totals);
  double total = 0.0;
  for (int s = 0; s < n_samples; ++s) total += totals[s];
  return fill_calls(g, haploid, n_samples, n_alleles, total, out);
}

extern "C" int ltr_genotype_locus(ltr_ctx* ctx, int haploid, int32_t n_samples, const int32_t* reads_per_sample,
                                  int32_t n_alleles, double* ll, const double* log_p1, const double* log_p2,
                                  ltr_locus_calls* out) {
  using namespace ltr;
  if (!ctx || !reads_per_sample || !ll || !log_p1 || !log_p2 || !out || n_samples < 1 || n_alleles < 1)
    return LTR_ERR_INVALID;
  std::vector<std::string> names;
  std::vector<std::vector<double> > p1((size_t)n_samples), p2((size_t)n_samples);
  size_t idx = 0;
  for (int s = 0; s < n_samples; ++s) {
    names.push_back("S" + std::to_string(s));
    if (reads_per_sample[s] < 0) return LTR_ERR_INVALID;
    for (int r = 0; r < reads_per_sample[s]; ++r, ++idx) {
      p1[s].push_back(log_p1[idx]);
      p2[s].push_back(log_p2[idx]);
    }
  }
  Genotyper g(haploid != 0, names, p1, p2, ctx);
  if (g.status() != LTR_OK) return g.status();
  g.set_num_alleles(n_alleles);
  const size_t nll = (size_t)g.num_reads() * n_alleles;
  std::copy(ll, ll + nll, g.log_aln_probs());
  const double total = g.calc_log_sample_posteriors();
  if (g.status() != LTR_OK) return g.status();
  std::copy(g.log_aln_probs(), g.log_aln_probs() + nll, ll);  // clamped in place, genotyper.cpp:57-58
  return fill_calls(g, haploid, n_samples, n_alleles, total, out);
}

// SeqStutterGenotyper::genotype after the first posterior pass (reference src/seq_stutter_genotyper.cpp:636-645):
// alleles that are in no sample's optimal haplotype pair are dropped (get_unused_alleles with check_called, :250-311;
// only samples with at least one aligned read vote, :262-266; the reference allele always stays), the LL columns of
// the kept alleles are carried over (add_and_remove_alleles, :317-409 -- no realignment when nothing is added) and
// the posteriors are recomputed on the reduced allele set.
extern "C" int ltr_genotype_locus_pruned(ltr_ctx* ctx, int haploid, int32_t n_samples, const int32_t* reads_per_sample,
                                         int32_t n_alleles, double* ll, const double* log_p1, const double* log_p2,
                                         const int32_t* seed_positions, int32_t* kept_alleles, int32_t* n_kept,
                                         ltr_locus_calls* out) {
  using namespace ltr;
  if (!ctx || !reads_per_sample || !ll || !log_p1 || !log_p2 || !out || !kept_alleles || !n_kept || n_samples < 1 ||
      n_alleles < 1)
    return LTR_ERR_INVALID;
  std::vector<std::string> names;
  std::vector<std::vector<double> > p1((size_t)n_samples), p2((size_t)n_samples);
  std::vector<bool> aligned_read((size_t)n_samples, false);
  size_t idx = 0;
  for (int s = 0; s < n_samples; ++s) {
    names.push_back("S" + std::to_string(s));
    if (reads_per_sample[s] < 0) return LTR_ERR_INVALID;
    for (int r = 0; r < reads_per_sample[s]; ++r, ++idx) {
      p1[s].push_back(log_p1[idx]);
      p2[s].push_back(log_p2[idx]);
      if (seed_positions == NULL || seed_positions[idx] >= 0) aligned_read[s] = true;
    }
  }
  const size_t R = idx;
  // first pass on the full allele set
  Genotyper g(haploid != 0, names, p1, p2, ctx);
  if (g.status() != LTR_OK) return g.status();
  g.set_num_alleles(n_alleles);
  std::copy(ll, ll + R * n_alleles, g.log_aln_probs());
  g.calc_log_sample_posteriors();
  if (g.status() != LTR_OK) return g.status();
  std::copy(g.log_aln_probs(), g.log_aln_probs() + R * n_alleles, ll);  // clamped in place
  std::vector<std::pair<int, int> > haps;
  g.get_optimal_haplotypes(haps);
  std::vector<bool> called((size_t)n_alleles, false);
  for (int s = 0; s < n_samples; ++s)
    if (aligned_read[s]) {
      called[haps[s].first] = true;
      called[haps[s].second] = true;
    }
  int kept = 0;
  for (int a = 0; a < n_alleles; ++a)
    if (a == 0 || called[a]) kept_alleles[kept++] = a;
  *n_kept = kept;
  if (kept == n_alleles) {  // nothing to remove: the first pass stands (seq_stutter_genotyper.cpp:641)
    double total = 0.0;
    for (int s = 0; s < n_samples; ++s) total += g.sample_total_LLs()[s];
    return fill_calls(g, haploid, n_samples, n_alleles, total, out);
  }
  Genotyper g2(haploid != 0, names, p1, p2, ctx);
  g2.set_num_alleles(kept);
  for (size_t r = 0; r < R; ++r)
    for (int k = 0; k < kept; ++k) g2.log_aln_probs()[r * kept + k] = ll[r * n_alleles + kept_alleles[k]];
  const double total = g2.calc_log_sample_posteriors();
  if (g2.status() != LTR_OK) return g2.status();
  return fill_calls(g2, haploid, n_samples, kept, total, out);
}

// ---- C ABI: host-only helpers of HapAligner (no GPU needed) ------------------------------------------------
namespace {
struct FlatView {  // the three blocks + one read of a flat locus as host-mirror objects
  ltr::StutterModel model;
  ltr::HapBlock left, right;
  ltr::RepeatBlock repeat;
  std::vector<ltr::HapBlock*> blocks;
  explicit FlatView(const ltr_flat_locus* L)
      : model(L->stutter[0], L->stutter[1], L->stutter[2], L->stutter[3], L->stutter[4], L->stutter[5], L->motif),
        left(L->repeat_start - (int32_t)strlen(L->lflank), L->repeat_start, L->lflank),
        right(L->repeat_end, L->repeat_end + (int32_t)strlen(L->rflank), L->rflank),
        repeat(L->repeat_start, L->repeat_end, L->alleles[0], L->period, &model) {
    for (int a = 1; a < L->n_alleles; ++a) repeat.add_alternate(std::make_pair(std::string(L->alleles[a]), false));
    blocks.push_back(&left);
    blocks.push_back(&repeat);
    blocks.push_back(&right);
  }
};
bool flat_ok(const ltr_flat_locus* L, int32_t read_index) {
  return L && L->lflank && L->rflank && L->alleles && L->n_alleles >= 1 && L->motif && L->reads && read_index >= 0 &&
         read_index < L->n_reads && L->reads[read_index].seq && L->reads[read_index].cigar && L->period >= 1;
}
}  // namespace

extern "C" int32_t ltr_trim_read_flat(const ltr_flat_locus* L, int32_t read_index, char* out, int32_t cap) {
  using namespace ltr;
  if (!flat_ok(L, read_index) || !out) return LTR_ERR_INVALID;
  FlatView v(L);
  Haplotype hap(v.blocks);
  std::vector<bool> all((size_t)hap.num_combs(), true);
  std::vector<float> params;
  HapAligner aligner(&hap, all, L->indel_flank_len, L->switch_old_align_len, params, NULL);
  const ltr_flat_read& fr = L->reads[read_index];
  Alignment aln(fr.start, fr.stop, false, false, "read", fr.qual ? fr.qual : "", fr.seq, fr.seq);
  if (!aln.set_cigar_string(fr.cigar)) return LTR_ERR_INVALID;
  std::string trimmed;
  if (!aligner.trim_alignment(aln, trimmed)) return LTR_ERR_INVALID;
  if ((int32_t)trimmed.size() + 1 > cap) return LTR_ERR_INVALID;
  memcpy(out, trimmed.c_str(), trimmed.size() + 1);
  return (int32_t)trimmed.size();
}

extern "C" int32_t ltr_seed_base_flat(const ltr_flat_locus* L, int32_t read_index) {
  using namespace ltr;
  if (!flat_ok(L, read_index)) return LTR_ERR_INVALID;
  FlatView v(L);
  Haplotype hap(v.blocks);
  std::vector<bool> all((size_t)hap.num_combs(), true);
  std::vector<float> params;
  HapAligner aligner(&hap, all, L->indel_flank_len, L->switch_old_align_len, params, NULL);
  const ltr_flat_read& fr = L->reads[read_index];
  Alignment aln(fr.start, fr.stop, false, false, "read", fr.qual ? fr.qual : "", fr.seq, fr.seq);
  if (!aln.set_cigar_string(fr.cigar)) return LTR_ERR_INVALID;
  const int seed = aligner.calc_seed_base(aln);
  return seed == -2 ? LTR_ERR_INVALID : seed;
}
