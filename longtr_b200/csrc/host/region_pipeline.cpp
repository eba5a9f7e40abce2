// region_pipeline.cpp -- ltr_regions_run: BAM files + regions of one chromosome -> genotype calls.
// The region loop of LongTR (reference BamProcessor::process_regions, src/bam_processor.cpp:536-628, and
// GenotyperBamProcessor::analyze_reads_and_phasing, src/genotyper_bam_processor.cpp:227-351) re-cut for a GPU: instead of
// one locus at a time through seek -> filter -> trim -> candidate alleles -> align -> posteriors, the host threads prepare the
// reads (ltr_region_collect) and candidate alleles (ltr_candidate_alleles) of MANY regions, the survivors are laid out as one
// ltr_locus_batch, and ltr_genotyper_run aligns and genotypes them as a few asynchronous GPU jobs.
// Regions the reference skips (too long, too near the contig ends, too few reads, no spanning alignments) are reported with
// their reason and do not enter the batch.  Regions whose reads are not explained by exact candidates get consensus alleles
// from the assembly branch of ltr_candidate_alleles (clustering + partial-order consensus on the preparing host thread);
// opts->no_assembly reports them as LTR_REGION_NEEDS_ASSEMBLY instead.
#include <limits.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <memory>
#include <new>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "longtr_b200.h"

namespace {

struct RegionWork {  // what the preparation leaves behind for one region; freed with it
  ltr_region_reads* reads = nullptr;
  ltr_candidates* cand = nullptr;
  int32_t status = LTR_REGION_OK;
  RegionWork() {}
  RegionWork(const RegionWork&) = delete;
  RegionWork& operator=(const RegionWork&) = delete;
  void release() {
    ltr_candidates_free(cand);
    ltr_region_reads_free(reads);
    cand = nullptr;
    reads = nullptr;
  }
  ~RegionWork() { release(); }
};

struct Owner {
  ltr_regions_result pub;
  std::vector<int32_t> status, locus_index, block_start, block_end;
  std::vector<uint32_t> region_allele_begin, allele_off, region_sample_begin, sample_file;
  std::vector<uint8_t> allele_bytes, allele_inexact;
  std::vector<uint32_t> record_off;
  std::string records;
};

// fn(i) for every i in [0, n) on n_threads host threads (the caller's among them), `grab` consecutive items at a time.
// false when an item threw (exhausted memory): no exception leaves a worker thread or the library.
template <typename F>
bool parallel_items(uint32_t n, uint32_t grab, int n_threads, F fn) {
  std::atomic<uint32_t> next(0);
  std::atomic<bool> failed(false);
  auto loop = [&]() {
    for (uint32_t i0 = next.fetch_add(grab); i0 < n; i0 = next.fetch_add(grab))
      for (uint32_t i = i0; i < std::min(n, i0 + grab); ++i) {
        try {
          fn(i);
        } catch (...) {
          failed.store(true);
        }
      }
  };
  std::vector<std::thread> th;
  try {
    for (int t = 1; t < n_threads; ++t) th.emplace_back(loop);
  } catch (...) {  // no more threads to be had: the ones that started (and this one) do the work
  }
  loop();
  for (std::thread& t : th) t.join();
  return !failed.load();
}

}  // namespace

static int regions_run_impl(ltr_genotyper* g, const ltr_params* params, const ltr_bam* const* bams, int32_t n_bams,
                            const char* chrom, const ltr_region* regions, uint32_t n_regions, const uint8_t* ref_seq,
                            int64_t ref_seq_start, int64_t ref_seq_len, const ltr_region_params* rp,
                            const ltr_regions_opts* opts, ltr_regions_result** out);

// C ABI boundary: no exception leaves the library (exhausted memory becomes an error code; what was built is freed)
extern "C" int ltr_regions_run(ltr_genotyper* g, const ltr_params* params, const ltr_bam* const* bams, int32_t n_bams,
                               const char* chrom, const ltr_region* regions, uint32_t n_regions, const uint8_t* ref_seq,
                               int64_t ref_seq_start, int64_t ref_seq_len, const ltr_region_params* rp,
                               const ltr_regions_opts* opts, ltr_regions_result** out) {
  try {
    return regions_run_impl(g, params, bams, n_bams, chrom, regions, n_regions, ref_seq, ref_seq_start, ref_seq_len, rp, opts, out);
  } catch (const std::bad_alloc&) {
    return LTR_ERR_OOM;
  } catch (...) {
    return LTR_ERR_INVALID;
  }
}

static int regions_run_impl(ltr_genotyper* g, const ltr_params* params, const ltr_bam* const* bams, int32_t n_bams,
                            const char* chrom, const ltr_region* regions, uint32_t n_regions, const uint8_t* ref_seq,
                            int64_t ref_seq_start, int64_t ref_seq_len, const ltr_region_params* rp,
                            const ltr_regions_opts* opts, ltr_regions_result** out) {
  if (!g || !params || !bams || n_bams < 1 || !chrom || (n_regions && !regions) || !ref_seq || !rp || !opts || !out)
    return LTR_ERR_INVALID;
  if (opts->vcf_records && n_regions && !opts->region_motifs) return LTR_ERR_INVALID;
  *out = nullptr;
  std::vector<RegionWork> work(n_regions);
  int n_threads = opts->host_threads > 0 ? opts->host_threads : (int)std::thread::hardware_concurrency();
  if (n_threads < 1) n_threads = 1;
  if ((uint32_t)n_threads > n_regions) n_threads = n_regions ? (int)n_regions : 1;
  typedef std::chrono::steady_clock Clock;
  auto ms_since = [](Clock::time_point t) { return std::chrono::duration<double, std::milli>(Clock::now() - t).count(); };
  const Clock::time_point t_begin = Clock::now();
  // a few consecutive regions per grab: neighbours of a sorted region list share BGZF blocks (the reader keeps the last two
  // blocks a thread inflated)
  const bool prepared = parallel_items(n_regions, 4, n_threads, [&](uint32_t r) {
    RegionWork& W = work[r];
    const ltr_region& R = regions[r];
    if (R.stop <= R.start || R.period < 1) { W.status = LTR_REGION_INVALID; return; }
    if (R.stop - R.start > opts->max_tr_len) { W.status = LTR_REGION_TOO_LONG; return; }  // bam_processor.cpp:568-574
    if (R.start < 50 || (int64_t)R.stop + 50 >= ref_seq_start + ref_seq_len) {               // :583-586
      W.status = LTR_REGION_NEAR_CONTIG_END;
      return;
    }
    int rc = ltr_region_collect(bams, n_bams, chrom, R.start, R.stop, ref_seq, ref_seq_start, ref_seq_len, rp, &W.reads);
    if (rc == LTR_ERR_UNSUPPORTED) { W.status = LTR_REGION_PAIRED_READS; return; }
    if (rc != LTR_OK) { W.status = LTR_REGION_INVALID; return; }
    if ((int32_t)W.reads->n_passed < opts->min_total_reads) {  // genotyper_bam_processor.cpp:231-238
      W.status = LTR_REGION_TOO_FEW_READS;
      return;
    }
    if (W.reads->n_reads == 0) { W.status = LTR_REGION_NO_SPANNING; return; }
    for (uint32_t i = 0; i < W.reads->n_reads && W.status == LTR_REGION_OK; ++i)
      if (W.reads->read_off[i + 1] == W.reads->read_off[i]) W.status = LTR_REGION_DELETED_READ;  // empty (deleted) reads: not batched
    if (W.status != LTR_REGION_OK) return;
    rc = ltr_candidate_alleles_flags(W.reads, R.start, R.stop, R.period, ref_seq, ref_seq_start, ref_seq_len,
                                     params->indel_flank_len, opts->no_assembly ? LTR_CAND_FLAG_NO_ASSEMBLY : 0u, &W.cand);
    if (rc != LTR_OK) { W.status = LTR_REGION_INVALID; return; }
    switch (W.cand->status) {
      case LTR_CAND_OK: break;
      case LTR_CAND_NEAR_CHROM_END: W.status = LTR_REGION_NEAR_CONTIG_END; break;
      case LTR_CAND_NO_SPANNING: W.status = LTR_REGION_NO_SPANNING; break;
      default: W.status = LTR_REGION_NEEDS_ASSEMBLY; break;
    }
  });
  if (!prepared) return LTR_ERR_OOM;
  const double prepare_ms = ms_since(t_begin);
  const Clock::time_point t_layout = Clock::now();
  // ---- lay the surviving regions out as one ltr_locus_batch ------------------------------------------------------------
  // Sizes and offsets first (serial, a few integers per region); the bytes are then copied -- and the per-region work freed --
  // by the host threads, each region into its own slice.
  std::unique_ptr<Owner> owner(new Owner());  // handed to the result at the end
  Owner* O = owner.get();
  memset(&O->pub, 0, sizeof(O->pub));
  O->status.resize(n_regions);
  O->locus_index.assign(n_regions, -1);
  O->block_start.assign(n_regions, 0);
  O->block_end.assign(n_regions, 0);
  O->region_allele_begin.assign(1, 0);
  O->region_sample_begin.assign(1, 0);
  std::vector<uint32_t> lflank_off(1, 0), rflank_off(1, 0), lab(1, 0), aoff, lrb(1, 0), roff, coff, cops, lns;
  std::vector<uint8_t> lfb, rfb, ab, rb;
  std::vector<int32_t> rstart, rend, read_start, read_stop, read_sample;
  std::vector<double> p1, p2;
  std::vector<int32_t> read_bp_diff, n_hp1, n_hp2;   // for the records: ExtractCigar per read, HP counts per (locus, sample)
  std::vector<uint32_t> locus_sample_begin(1, 0), locus_region;
  std::vector<uint64_t> o_ab_off(1, 0), ab_off(1, 0), rb_off(1, 0), cop_off(1, 0);  // byte / operation offsets per region / locus
  const bool want_records = opts->vcf_records != 0;
  const bool want_pgl = want_records && !opts->haploid && (opts->vcf_switches & LTR_VCF_PHASED_GLS);
  uint32_t n_loci = 0;
  for (uint32_t r = 0; r < n_regions; ++r) {
    const RegionWork& W = work[r];
    O->status[r] = W.status;
    O->region_sample_begin.push_back(O->region_sample_begin.back() + (W.reads ? W.reads->n_samples : 0u));
    const bool has_alleles = W.cand && W.cand->n_alleles > 0;
    const uint32_t na = has_alleles ? (uint32_t)W.cand->n_alleles : 0u;
    const uint64_t nab = has_alleles ? (uint64_t)(W.cand->allele_off[na] - W.cand->allele_off[0]) : 0u;
    if (has_alleles) {
      O->block_start[r] = W.cand->block_start;
      O->block_end[r] = W.cand->block_end;
      if (W.cand->status == LTR_CAND_OK && W.cand->n_cluster_samples > 0) ++O->pub.n_assembled;
    }
    O->region_allele_begin.push_back(O->region_allele_begin.back() + na);
    o_ab_off.push_back(o_ab_off.back() + nab);
    if (W.status != LTR_REGION_OK) continue;
    const ltr_region_reads& RR = *W.reads;
    const ltr_candidates& C = *W.cand;
    O->locus_index[r] = (int32_t)n_loci++;
    locus_region.push_back(r);
    lflank_off.push_back(lflank_off.back() + (uint32_t)strlen(C.lflank));
    rflank_off.push_back(rflank_off.back() + (uint32_t)strlen(C.rflank));
    lab.push_back(lab.back() + na);
    ab_off.push_back(ab_off.back() + nab);
    rstart.push_back(C.block_start);
    rend.push_back(C.block_end);
    lrb.push_back(lrb.back() + RR.n_reads);
    rb_off.push_back(rb_off.back() + (RR.read_off[RR.n_reads] - RR.read_off[0]));
    cop_off.push_back(cop_off.back() + (RR.cigar_off[RR.n_reads] - RR.cigar_off[0]));
    lns.push_back(RR.n_samples);
    locus_sample_begin.push_back(locus_sample_begin.back() + RR.n_samples);
  }
  if (o_ab_off.back() > 0xFFFFFFF0ull || ab_off.back() > 0xFFFFFFF0ull || rb_off.back() > 0xFFFFFFF0ull || cop_off.back() > 0xFFFFFFF0ull) {
    return LTR_ERR_INVALID;  // more than 4 GB of reads in one call: the batch's offsets are 32 bits wide
  }
  {
    const size_t n_reads_all = lrb.back();
    O->sample_file.resize(O->region_sample_begin.back());
    O->allele_off.assign((size_t)O->region_allele_begin.back() + 1, 0);
    O->allele_bytes.resize(o_ab_off.back());
    O->allele_inexact.resize(O->region_allele_begin.back());
    // (one spare element each: the batch's byte arrays get a terminator appended below without growing)
    lfb.reserve((size_t)lflank_off.back() + 1); lfb.resize(lflank_off.back());
    rfb.reserve((size_t)rflank_off.back() + 1); rfb.resize(rflank_off.back());
    aoff.assign((size_t)lab.back() + 1, 0);
    ab.reserve(ab_off.back() + 1); ab.resize(ab_off.back());
    roff.assign(n_reads_all + 1, 0);
    coff.assign(n_reads_all + 1, 0);
    rb.reserve(rb_off.back() + 1); rb.resize(rb_off.back());
    cops.reserve(cop_off.back() + 1); cops.resize(cop_off.back());
    read_start.resize(n_reads_all); read_stop.resize(n_reads_all); read_sample.resize(n_reads_all);
    p1.resize(n_reads_all); p2.resize(n_reads_all);
    if (want_records) {
      read_bp_diff.resize(n_reads_all);
      n_hp1.assign(locus_sample_begin.back(), 0);
      n_hp2.assign(locus_sample_begin.back(), 0);
    }
  }
  parallel_items(n_regions, 8, n_threads, [&](uint32_t r) {
    RegionWork& W = work[r];
    if (W.reads)
      for (uint32_t s = 0; s < W.reads->n_samples; ++s) O->sample_file[O->region_sample_begin[r] + s] = W.reads->sample_file[s];
    if (W.cand && W.cand->n_alleles > 0) {
      const ltr_candidates& C = *W.cand;
      const uint32_t a0 = O->region_allele_begin[r], c0 = C.allele_off[0];
      memcpy(O->allele_bytes.data() + o_ab_off[r], C.allele_bytes + c0, C.allele_off[C.n_alleles] - c0);
      for (int32_t a = 0; a < C.n_alleles; ++a) {
        O->allele_off[a0 + (uint32_t)a + 1] = (uint32_t)o_ab_off[r] + (C.allele_off[a + 1] - c0);
        O->allele_inexact[a0 + (uint32_t)a] = C.allele_inexact[a];
      }
    }
    if (W.status == LTR_REGION_OK) {
      const ltr_region_reads& RR = *W.reads;
      const ltr_candidates& C = *W.cand;
      const uint32_t l = (uint32_t)O->locus_index[r];
      memcpy(lfb.data() + lflank_off[l], C.lflank, lflank_off[l + 1] - lflank_off[l]);
      memcpy(rfb.data() + rflank_off[l], C.rflank, rflank_off[l + 1] - rflank_off[l]);
      const uint32_t c0 = C.allele_off[0];
      memcpy(ab.data() + ab_off[l], C.allele_bytes + c0, C.allele_off[C.n_alleles] - c0);
      for (int32_t a = 0; a < C.n_alleles; ++a) aoff[lab[l] + (uint32_t)a + 1] = (uint32_t)ab_off[l] + (C.allele_off[a + 1] - c0);
      const uint32_t i0 = lrb[l], n = RR.n_reads, b0 = RR.read_off[0], g0 = RR.cigar_off[0];
      memcpy(rb.data() + rb_off[l], RR.read_bytes + b0, RR.read_off[n] - b0);
      if (RR.cigar_off[n] > g0) memcpy(cops.data() + cop_off[l], RR.cigar_ops + g0, sizeof(uint32_t) * (RR.cigar_off[n] - g0));
      for (uint32_t i = 0; i < n; ++i) {
        read_start[i0 + i] = RR.read_start[i];
        read_stop[i0 + i] = RR.read_stop[i];
        roff[i0 + i + 1] = (uint32_t)rb_off[l] + (RR.read_off[i + 1] - b0);
        coff[i0 + i + 1] = (uint32_t)cop_off[l] + (RR.cigar_off[i + 1] - g0);
        read_sample[i0 + i] = RR.read_sample[i];
        p1[i0 + i] = RR.log_p1[i];
        p2[i0 + i] = RR.log_p2[i];
      }
      if (want_records) {
        const size_t h0 = locus_sample_begin[l];
        for (uint32_t i = 0; i < n; ++i) {
          int32_t d = 0;  // write_vcf_record (:1017-1023): ExtractCigar over the region +- 5 bp
          const int got = ltr_extract_cigar_bp_diff(RR.cigar_ops + RR.cigar_off[i], RR.cigar_off[i + 1] - RR.cigar_off[i],
                                                    RR.read_start[i], regions[r].start - 5, regions[r].stop + 5, &d);
          read_bp_diff[i0 + i] = got ? d : INT32_MIN;
          if (RR.read_hp[i] == 1) ++n_hp1[h0 + (size_t)RR.read_sample[i]];
          if (RR.read_hp[i] == 2) ++n_hp2[h0 + (size_t)RR.read_sample[i]];
        }
      }
    }
    W.release();
  });
  int rc = LTR_OK;
  ltr_batch_calls* calls = nullptr;
  std::unique_ptr<ltr_batch_calls, void (*)(ltr_batch_calls*)> calls_guard(nullptr, ltr_batch_calls_free);
  const double layout_ms = ms_since(t_layout);
  const Clock::time_point t_geno = Clock::now();
  if (n_loci) {
    lfb.push_back(0); rfb.push_back(0); ab.push_back(0); rb.push_back(0); cops.push_back(0);
    ltr_locus_batch B;
    memset(&B, 0, sizeof(B));
    B.n_loci = n_loci;
    B.lflank_off = lflank_off.data(); B.lflank_bytes = lfb.data();
    B.rflank_off = rflank_off.data(); B.rflank_bytes = rfb.data();
    B.locus_allele_begin = lab.data(); B.allele_off = aoff.data(); B.allele_bytes = ab.data();
    B.repeat_start = rstart.data(); B.repeat_end = rend.data();
    B.locus_read_begin = lrb.data(); B.read_start = read_start.data(); B.read_stop = read_stop.data();
    B.read_off = roff.data(); B.read_bytes = rb.data(); B.cigar_off = coff.data(); B.cigar_ops = cops.data();
    B.read_sample = read_sample.data(); B.log_p1 = p1.data(); B.log_p2 = p2.data();
    B.second_mate = nullptr;
    B.locus_n_samples = lns.data();
    std::vector<uint8_t> haploid_flags;
    if (opts->haploid) haploid_flags.assign((size_t)n_loci + 1, 1);
    B.locus_haploid = opts->haploid ? haploid_flags.data() : nullptr;
    if (want_records) ltr_genotyper_set_read_alleles(g, 1);
    if (want_pgl) ltr_genotyper_set_phased_gls(g, 1);
    rc = ltr_genotyper_run(g, params, &B, &calls);
    if (want_records) ltr_genotyper_set_read_alleles(g, 0);
    if (want_pgl) ltr_genotyper_set_phased_gls(g, 0);
    calls_guard.reset(calls);
  }
  const double genotype_ms = ms_since(t_geno);
  const Clock::time_point t_rec = Clock::now();
  if (rc == LTR_OK && want_records) {
    O->record_off.assign((size_t)n_regions + 1, 0);
    std::vector<std::string> text(n_regions);
    std::atomic<int> rec_rc(LTR_OK);
    // the records are independent of each other: composed by the host threads
    const bool composed = parallel_items(n_loci, 8, n_threads, [&](uint32_t l) {
      std::vector<char> buf(4096);
      const uint32_t r = locus_region[l];
      if (calls->status[l] != LTR_OK || rec_rc.load() != LTR_OK) return;
      const uint32_t a0 = O->region_allele_begin[r], na = O->region_allele_begin[r + 1] - a0;
      const uint32_t s0 = calls->locus_sample_begin[l], ns = calls->locus_sample_begin[l + 1] - s0;
      std::vector<uint32_t> rec_aoff(na + 1);  // the region's allele offsets, re-based
      for (uint32_t a = 0; a <= na; ++a) rec_aoff[a] = O->allele_off[a0 + a] - O->allele_off[a0];
      std::vector<int32_t> column((size_t)n_bams, -1);
      for (uint32_t s = 0; s < ns; ++s) {
        const uint32_t f = O->sample_file[O->region_sample_begin[r] + s];
        if (f < (uint32_t)n_bams) column[f] = (int32_t)s;
      }
      ltr_vcf_locus V;
      memset(&V, 0, sizeof(V));
      V.chrom = chrom;
      V.name = opts->region_names ? opts->region_names[r] : "";
      V.motif = opts->region_motifs[r];
      V.region_start = regions[r].start; V.region_stop = regions[r].stop;
      V.chrom_seq = ref_seq; V.chrom_seq_start = ref_seq_start; V.chrom_seq_len = ref_seq_len;
      V.block_start = O->block_start[r]; V.block_end = O->block_end[r];
      V.n_alleles = (int32_t)na; V.allele_off = rec_aoff.data(); V.allele_bytes = O->allele_bytes.data() + O->allele_off[a0];
      V.allele_inexact = O->allele_inexact.data() + a0;
      V.kept_mask = calls->kept_mask + calls->locus_allele_begin[l];
      V.haploid = opts->haploid ? 1 : 0;
      V.n_samples = (int32_t)ns;
      V.gts = calls->gts + 2 * (size_t)s0;
      V.log_unphased_posteriors = calls->log_unphased_posteriors + s0;
      V.log_phased_posteriors = calls->log_phased_posteriors + s0;
      V.gl_diffs = calls->gl_diffs + s0;
      V.n_p1 = n_hp1.data() + locus_sample_begin[l]; V.n_p2 = n_hp2.data() + locus_sample_begin[l];
      V.n_reads = (int32_t)(lrb[l + 1] - lrb[l]);
      V.read_sample = read_sample.data() + lrb[l];
      V.log_p1 = p1.data() + lrb[l]; V.log_p2 = p2.data() + lrb[l];
      V.read_bp_diff = read_bp_diff.data() + lrb[l];
      V.read_allele = calls->read_allele ? calls->read_allele + lrb[l] : nullptr;
      V.n_columns = n_bams; V.column_sample = column.data();
      ltr_vcf_extras X;  // slices re-based to the locus' first sample
      memset(&X, 0, sizeof(X));
      X.switches = opts->vcf_switches;
      std::vector<uint64_t> glb(ns + 1), pglb(ns + 1, 0);
      for (uint32_t s = 0; s <= ns; ++s) glb[s] = calls->gl_begin[s0 + s] - calls->gl_begin[s0];
      X.gl_begin = glb.data(); X.gls = calls->gls + calls->gl_begin[s0]; X.pls = calls->pls + calls->gl_begin[s0];
      if (calls->pgl_begin) {
        for (uint32_t s = 0; s <= ns; ++s) pglb[s] = calls->pgl_begin[s0 + s] - calls->pgl_begin[s0];
        X.pgl_begin = pglb.data(); X.phased_gls = calls->phased_gls + calls->pgl_begin[s0];
      }
      uint32_t len = 0;
      int vrc = ltr_vcf_record_ex(&V, &X, buf.data(), (uint32_t)buf.size(), &len);
      if (vrc == LTR_ERR_INVALID && (size_t)len + 1 > buf.size()) {
        buf.resize((size_t)len + 16);
        vrc = ltr_vcf_record_ex(&V, &X, buf.data(), (uint32_t)buf.size(), &len);
      }
      if (vrc == LTR_OK) text[r].assign(buf.data(), len);
      else if (vrc != LTR_ERR_UNSUPPORTED) rec_rc.store(vrc);
    });
    rc = composed ? rec_rc.load() : LTR_ERR_OOM;
    for (uint32_t r = 0; r < n_regions; ++r) {
      O->records += text[r];
      O->record_off[r + 1] = (uint32_t)O->records.size();
    }
  }
  if (rc != LTR_OK) return rc;
  ltr_regions_result& P = O->pub;
  P.n_regions = n_regions;
  P.status = O->status.data();
  P.locus_index = O->locus_index.data();
  P.n_loci = n_loci;
  P.calls = calls;
  P.block_start = O->block_start.data();
  P.block_end = O->block_end.data();
  P.region_allele_begin = O->region_allele_begin.data();
  P.allele_off = O->allele_off.data();
  P.allele_bytes = O->allele_bytes.data();
  P.allele_inexact = O->allele_inexact.data();
  P.record_off = want_records ? O->record_off.data() : nullptr;
  P.records = want_records ? O->records.c_str() : nullptr;
  P.region_sample_begin = O->region_sample_begin.data();
  P.sample_file = O->sample_file.data();
  P.owner = O;
  P.prepare_ms = prepare_ms; P.layout_ms = layout_ms; P.genotype_ms = genotype_ms; P.records_ms = ms_since(t_rec);
  owner.release();
  calls_guard.release();
  *out = &O->pub;
  return LTR_OK;
}

// ---- the whole run: BAM files + FASTA + region file -> calls (BamProcessor::process_regions, src/bam_processor.cpp:536-628) ----
namespace {
struct RunOwner {
  ltr_bed_run_result pub;
  std::vector<ltr_regions_result*> per_chrom;
  std::vector<uint32_t> begin;
};
}  // namespace

extern "C" int ltr_run_bed(ltr_genotyper* g, const ltr_params* params, const ltr_bam* const* bams, int32_t n_bams,
                           const ltr_fasta* fasta, const ltr_bed* bed, const ltr_region_params* rp, const ltr_regions_opts* opts,
                           ltr_bed_run_result** out) {
  if (!g || !params || !bams || n_bams < 1 || !fasta || !bed || !rp || !opts || !out) return LTR_ERR_INVALID;
  *out = nullptr;
  // verify_chromosomes (:490-531): every chromosome of the region file must exist in the FASTA and in the alignment files
  for (uint32_t c = 0; c < bed->n_chroms; ++c) {
    if (ltr_fasta_seq_len(fasta, bed->chroms[c]) < 0) return LTR_ERR_INVALID;
    for (int32_t b = 0; b < n_bams; ++b)
      if (ltr_bam_ref_id(bams[b], bed->chroms[c]) < 0) return LTR_ERR_INVALID;
  }
  RunOwner* O = new RunOwner();
  memset(&O->pub, 0, sizeof(O->pub));
  O->begin.push_back(0);
  int rc = LTR_OK;
  std::vector<uint8_t> chrom_seq;
  uint32_t r = 0;
  for (uint32_t c = 0; c < bed->n_chroms && rc == LTR_OK; ++c) {
    uint32_t e = r;
    while (e < bed->n_regions && bed->region_chrom[e] == (int32_t)c) ++e;
    const int64_t len = ltr_fasta_seq_len(fasta, bed->chroms[c]);
    chrom_seq.resize((size_t)len + 1);
    rc = ltr_fasta_fetch(fasta, bed->chroms[c], 0, len, chrom_seq.data());
    ltr_regions_result* res = nullptr;
    ltr_regions_opts o = *opts;  // names and motifs of this chromosome's regions for the records
    o.region_names = bed->names + r;
    o.region_motifs = bed->motifs + r;
    if (rc == LTR_OK)
      rc = ltr_regions_run(g, params, bams, n_bams, bed->chroms[c], bed->regions + r, e - r, chrom_seq.data(), 0, len, rp, &o,
                           &res);
    O->per_chrom.push_back(res);
    O->begin.push_back(e);
    r = e;
  }
  if (rc != LTR_OK) {
    for (ltr_regions_result* p : O->per_chrom) ltr_regions_result_free(p);
    delete O;
    return rc;
  }
  O->pub.n_chroms = bed->n_chroms;
  O->pub.per_chrom = O->per_chrom.data();
  O->pub.chrom_region_begin = O->begin.data();
  O->pub.owner = O;
  *out = &O->pub;
  return LTR_OK;
}

// ltr_run_bed in bounded memory: a chromosome's regions go through ltr_regions_run `chunk_regions` at a time and every chunk's
// result is handed to the caller's sink in region order, then freed -- what a host writing a whole-genome VCF needs (the
// reference's VCFWriter receives its records region by region too, src/bam_processor.cpp:563-627).
extern "C" int ltr_run_bed_stream(ltr_genotyper* g, const ltr_params* params, const ltr_bam* const* bams, int32_t n_bams,
                                  const ltr_fasta* fasta, const ltr_bed* bed, const ltr_region_params* rp,
                                  const ltr_regions_opts* opts, int32_t chunk_regions, ltr_regions_sink sink, void* user) {
  if (!g || !params || !bams || n_bams < 1 || !fasta || !bed || !rp || !opts || !sink) return LTR_ERR_INVALID;
  for (uint32_t c = 0; c < bed->n_chroms; ++c) {  // verify_chromosomes (:490-531)
    if (ltr_fasta_seq_len(fasta, bed->chroms[c]) < 0) return LTR_ERR_INVALID;
    for (int32_t b = 0; b < n_bams; ++b)
      if (ltr_bam_ref_id(bams[b], bed->chroms[c]) < 0) return LTR_ERR_INVALID;
  }
  const uint32_t chunk = chunk_regions > 0 ? (uint32_t)chunk_regions : 4096u;
  try {
    std::vector<uint8_t> chrom_seq;
    uint32_t r = 0;
    for (uint32_t c = 0; c < bed->n_chroms; ++c) {
      uint32_t e = r;
      while (e < bed->n_regions && bed->region_chrom[e] == (int32_t)c) ++e;
      const int64_t len = ltr_fasta_seq_len(fasta, bed->chroms[c]);
      chrom_seq.resize((size_t)len + 1);
      int rc = ltr_fasta_fetch(fasta, bed->chroms[c], 0, len, chrom_seq.data());
      for (uint32_t r0 = r; r0 < e && rc == LTR_OK; r0 += chunk) {
        const uint32_t n = std::min(chunk, e - r0);
        ltr_regions_result* res = nullptr;
        ltr_regions_opts o = *opts;  // names and motifs of this chunk's regions for the records
        o.region_names = bed->names + r0;
        o.region_motifs = bed->motifs + r0;
        rc = ltr_regions_run(g, params, bams, n_bams, bed->chroms[c], bed->regions + r0, n, chrom_seq.data(), 0, len, rp, &o, &res);
        if (rc == LTR_OK && sink(user, c, r0, res) != 0) rc = LTR_ERR_INVALID;  // the sink asked to stop
        ltr_regions_result_free(res);
      }
      if (rc != LTR_OK) return rc;
      r = e;
    }
  } catch (const std::bad_alloc&) {
    return LTR_ERR_OOM;
  } catch (...) {
    return LTR_ERR_INVALID;
  }
  return LTR_OK;
}

extern "C" void ltr_bed_run_result_free(ltr_bed_run_result* r) {
  if (!r) return;
  RunOwner* O = static_cast<RunOwner*>(r->owner);
  for (ltr_regions_result* p : O->per_chrom) ltr_regions_result_free(p);
  delete O;
}

extern "C" void ltr_regions_opts_default(ltr_regions_opts* o) {
  if (!o) return;
  o->host_threads = 0;
  o->max_tr_len = 1000;      // MAX_STR_LENGTH (--max-tr-len), bam_processor.h:94
  o->min_total_reads = 10;   // MIN_TOTAL_READS (--min-reads), genotyper_bam_processor.h:110
  o->no_assembly = 0;
  o->vcf_records = 0;
  o->region_names = nullptr;
  o->region_motifs = nullptr;
  o->haploid = 0;
  o->vcf_switches = LTR_VCF_DEFAULT;
}

extern "C" void ltr_regions_result_free(ltr_regions_result* r) {
  if (!r) return;
  ltr_batch_calls_free(r->calls);
  delete static_cast<Owner*>(r->owner);
}
