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
#include <thread>
#include <vector>

#include "SeqAlignment/AlignmentData.h"
#include "SeqAlignment/HapAligner.h"
#include "SeqAlignment/HapBlock.h"
#include "SeqAlignment/Haplotype.h"
#include "SeqAlignment/RepeatBlock.h"
#include "base_quality.h"
#include "genotyper.h"
#include "mathops.h"
#include "read_pooler.h"
#include "stutter_model.h"

#include "longtr_b200.h"
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
  // Genotyper::extract_genotypes_and_likelihoods (src/genotyper.cpp:132-256) with hap_to_allele = identity
  int calls(int haploid, double total, ltr_locus_calls* out) {
    const int S = num_samples_, H = num_alleles_;
    std::vector<int> hap_to_allele(H);
    for (int a = 0; a < H; ++a) hap_to_allele[a] = a;
    std::vector<std::pair<int, int> > best_haps, best_gts;
    std::vector<double> lpp, lup, hlpp, hlup, gl_diffs;
    std::vector<std::vector<double> > gls, pgls;
    std::vector<std::vector<int> > pls;
    extract_genotypes_and_likelihoods(H, hap_to_allele, best_haps, best_gts, lpp, lup, hlpp, hlup, true, gls, gl_diffs,
                                      true, pls, true, pgls);
    const size_t n_gl = haploid ? H : H * (H + 1) / 2, n_pgl = haploid ? H : H * H;
    out->total_ll = total;
    for (int s = 0; s < S; ++s) {
      out->best_gts[2 * s] = best_gts[s].first;
      out->best_gts[2 * s + 1] = best_gts[s].second;
      out->log_phased_posteriors[s] = lpp[s];
      out->log_unphased_posteriors[s] = lup[s];
      out->hap_log_phased_posteriors[s] = hlpp[s];
      out->hap_log_unphased_posteriors[s] = hlup[s];
      out->gl_diffs[s] = gl_diffs[s];
      out->sample_total_lls[s] = sample_total_LLs_[s];
      if (gls[s].size() != n_gl || pls[s].size() != n_gl || pgls[s].size() != n_pgl) return -1;
      std::copy(gls[s].begin(), gls[s].end(), out->gls + s * n_gl);
      std::copy(pls[s].begin(), pls[s].end(), out->pls + s * n_gl);
      std::copy(pgls[s].begin(), pgls[s].end(), out->phased_gls + s * n_pgl);
    }
    std::memcpy(out->log_sample_posteriors, log_sample_posteriors_, sizeof(double) * S * H * H);
    return 0;
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

// Flattened batch (layout of ltr_viterbi_batch) through the reference classes: per locus a
// three-block Haplotype is rebuilt from the full haplotype strings (35 bp flanks) and every
// trimmed read becomes an Alignment that exactly covers [repeat_start-5, repeat_end+5), so
// that HapAligner::trim_alignment leaves it untouched.  Loci are sharded over n_threads
// std::threads (the reference itself is single-threaded; README.md:78-82 recommends splitting
// the BED).  *seconds = max over threads of the time spent inside process_reads.
int ltr_ref_viterbi_batch(uint32_t n_loci, const uint32_t* lhb, const uint32_t* lrb,
                          const uint32_t* hap_off, const uint8_t* hap_bytes, const uint32_t* read_off,
                          const uint8_t* read_bytes, int n_aln_params, const float* aln_params,
                          int n_threads, double* out_ll, double* seconds,
                          /* optional posterior stage (all NULL to skip): one sample per locus */
                          const uint32_t* lsb, const uint32_t* pool_index, const double* log_p1,
                          const double* log_p2, double* out_post) {
  ensure_tables();
  std::vector<uint64_t> post_off((size_t)n_loci + 1, 0);
  for (uint32_t l = 0; l < n_loci; ++l)
    post_off[l + 1] = post_off[l] + (uint64_t)(lhb[l + 1] - lhb[l]) * (lhb[l + 1] - lhb[l]);
  std::vector<uint64_t> ll_off((size_t)n_loci + 1, 0);
  for (uint32_t l = 0; l < n_loci; ++l)
    ll_off[l + 1] = ll_off[l] + (uint64_t)(lhb[l + 1] - lhb[l]) * (lrb[l + 1] - lrb[l]);
  if (n_threads < 1) n_threads = 1;
  std::vector<double> tsec((size_t)n_threads, 0.0);
  std::vector<int> trc((size_t)n_threads, 0);
  auto work = [&](int t) {
    for (uint32_t l = (uint32_t)t; l < n_loci; l += (uint32_t)n_threads) {
      const uint32_t H = lhb[l + 1] - lhb[l], P = lrb[l + 1] - lrb[l];
      if (H == 0 || P == 0) continue;
      std::vector<std::string> haps, reads, quals, cigars;
      for (uint32_t h = lhb[l]; h < lhb[l + 1]; ++h)
        haps.push_back(std::string((const char*)hap_bytes + hap_off[h], hap_off[h + 1] - hap_off[h]));
      bool ok = true;
      for (const std::string& s : haps) ok = ok && s.size() >= 70 && s.compare(0, 35, haps[0], 0, 35) == 0;
      if (!ok) { trc[t] = -3; continue; }
      const std::string lflank = haps[0].substr(0, 35), rflank = haps[0].substr(haps[0].size() - 35);
      std::vector<std::string> alleles;
      std::vector<const char*> allele_ptrs;
      for (const std::string& s : haps) alleles.push_back(s.substr(35, s.size() - 70));
      for (const std::string& s : alleles) allele_ptrs.push_back(s.c_str());
      const int32_t rs = 1000, re = rs + (int32_t)alleles[0].size();
      std::vector<ltr_flat_read> fr(P);
      for (uint32_t r = 0; r < P; ++r) {
        const uint32_t g = lrb[l] + r;
        reads.push_back(std::string((const char*)read_bytes + read_off[g], read_off[g + 1] - read_off[g]));
        quals.push_back(std::string(reads.back().size(), 'I'));
        const int m = (int)reads.back().size();
        if (m < 11) { trc[t] = -3; ok = false; break; }
        cigars.push_back("5=" + std::to_string(m - 10) + "I5=");
      }
      if (!ok) continue;
      for (uint32_t r = 0; r < P; ++r) {
        fr[r].start = rs - 5; fr[r].stop = re + 5 - 1;
        fr[r].seq = reads[r].c_str(); fr[r].qual = quals[r].c_str(); fr[r].cigar = cigars[r].c_str();
      }
      ltr_flat_locus L;
      std::memset(&L, 0, sizeof(L));
      L.lflank = lflank.c_str(); L.rflank = rflank.c_str();
      L.repeat_start = rs; L.repeat_end = re; L.period = 2; L.n_alleles = (int32_t)H;
      L.alleles = allele_ptrs.data();
      const double st[6] = {0.95, 0.05, 0.05, 0.95, 0.01, 0.01};
      std::memcpy(L.stutter, st, sizeof(st));
      L.motif = "AC";
      L.n_reads = (int32_t)P; L.reads = fr.data();
      L.indel_flank_len = 5; L.switch_old_align_len = 0;
      L.n_aln_params = n_aln_params;
      for (int i = 0; i < n_aln_params && i < 7; ++i) L.aln_params[i] = aln_params[i];
      std::vector<int32_t> seeds(P, 0);
      int rc = ltr_ref_process_reads(&L, out_ll + ll_off[l], seeds.data(), &tsec[t]);
      if (rc != 0) trc[t] = rc;
      if (lsb != NULL && out_post != NULL) {
        // per-read LL rows scattered from the pools (seq_stutter_genotyper.cpp:526-538), then
        // Genotyper::calc_log_sample_posteriors
        const uint32_t r0 = lsb[l], r1 = lsb[l + 1];
        const int32_t R = (int32_t)(r1 - r0);
        std::vector<double> rows((size_t)R * H), clamped((size_t)R * H);
        for (int32_t r = 0; r < R; ++r)
          std::memcpy(&rows[(size_t)r * H], out_ll + ll_off[l] + (size_t)pool_index[r0 + r] * H, sizeof(double) * H);
        double tot = 0.0;
        auto p0 = std::chrono::steady_clock::now();
        ltr_ref_log_sample_posteriors(0, 1, &R, (int)H, rows.data(), log_p1 + r0, log_p2 + r0, clamped.data(),
                                      out_post + post_off[l], &tot, NULL);
        tsec[t] += std::chrono::duration<double>(std::chrono::steady_clock::now() - p0).count();
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
  double mx = 0.0;
  int rc = 0;
  for (int t = 0; t < n_threads; ++t) { if (tsec[t] > mx) mx = tsec[t]; if (trc[t]) rc = trc[t]; }
  if (seconds) *seconds = mx;
  return rc;
}

// calc_log_sample_posteriors + extract_genotypes_and_likelihoods for one locus (all outputs of
// ltr_locus_calls must be non-NULL here).
int ltr_ref_genotype_locus(int haploid, int n_samples, const int32_t* reads_per_sample, int n_alleles,
                           const double* ll_in, const double* log_p1, const double* log_p2, double* ll_out,
                           ltr_locus_calls* out) {
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
  std::vector<double> post((size_t)n_samples * n_alleles * n_alleles), totals(n_samples);
  const double total = probe.run(ll_in, ll_out, post.data(), totals.data(), NULL);
  return probe.calls(haploid, total, out);
}

// HapAligner::calc_seed_base (src/SeqAlignment/HapAligner.cpp:493-542) for every read of a flat locus
int ltr_ref_seed_bases(const ltr_flat_locus* L, int32_t* out_seeds) {
  ensure_tables();
  StutterModel model(L->stutter[0], L->stutter[1], L->stutter[2], L->stutter[3], L->stutter[4], L->stutter[5],
                     std::string(L->motif));
  std::string lflank(L->lflank), rflank(L->rflank);
  std::vector<HapBlock*> blocks;
  blocks.push_back(new HapBlock(L->repeat_start - (int32_t)lflank.size(), L->repeat_start, lflank));
  RepeatBlock* rep = new RepeatBlock(L->repeat_start, L->repeat_end, std::string(L->alleles[0]), L->period, &model);
  for (int a = 1; a < L->n_alleles; ++a)
    rep->add_alternate(std::pair<std::string, bool>(std::string(L->alleles[a]), false));
  blocks.push_back(rep);
  blocks.push_back(new HapBlock(L->repeat_end, L->repeat_end + (int32_t)rflank.size(), rflank));
  Haplotype* hap = new Haplotype(blocks);
  std::vector<bool> realign_hap(L->n_alleles, true);
  std::vector<float> params;
  {
    HapAligner aligner(hap, realign_hap, L->indel_flank_len, L->switch_old_align_len, params);
    for (int r = 0; r < L->n_reads; ++r) {
      const ltr_flat_read& fr = L->reads[r];
      Alignment aln(fr.start, fr.stop, false, false, "read", std::string(fr.qual), std::string(fr.seq), std::string(fr.seq));
      parse_cigar(fr.cigar, aln);
      out_seeds[r] = aligner.calc_seed_base(aln);
    }
  }
  delete hap;
  for (size_t i = 0; i < blocks.size(); ++i) delete blocks[i];
  return 0;
}

const char* ltr_ref_version(void) { return "LongTR reference sources, compiled in place (oracle/_ref)"; }


// ReadPooler (src/read_pooler.cpp:3-20, read_pooler.h:42-48) on n reads given as sequence / quality strings: pool index of
// every read, number of pools and, back to back in pool order, the median qualities BaseQuality::median_base_qualities
// (src/base_quality.cpp:11-28) leaves on the pooled alignments.  Returns the bytes written to pooled_quals, < 0 on error.
int64_t ltr_ref_pool_reads(int32_t n_reads, const char* const* seqs, const char* const* quals, int32_t* pool_index,
                           int32_t* n_pools, char* pooled_quals, int64_t cap) {
  ReadPooler pooler;
  BaseQuality bq;
  for (int32_t r = 0; r < n_reads; ++r) {
    Alignment aln(100, 100 + (int32_t)strlen(seqs[r]) - 1, false, false, "read", std::string(quals[r]), std::string(seqs[r]),
                  std::string(seqs[r]));
    pool_index[r] = pooler.add_alignment(aln);
  }
  pooler.pool(bq);
  *n_pools = pooler.num_pools();
  int64_t off = 0;
  std::vector<Alignment>& alns = pooler.get_alignments();
  for (size_t i = 0; i < alns.size(); ++i) {
    const std::string& q = alns[i].get_base_qualities();
    if (off + (int64_t)q.size() > cap) return -1;
    memcpy(pooled_quals + off, q.data(), q.size());
    off += (int64_t)q.size();
  }
  return off;
}
}  // extern "C"
