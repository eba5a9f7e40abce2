// flat_api.cpp -- ltr_process_reads_flat: HapAligner::process_reads on one flat locus
// (include/longtr_b200_locus.h), through the host mirror of the reference classes.
//   blocks  : HapBlock(left flank), RepeatBlock(alleles, period, stutter model), HapBlock(right flank)
//             -- what HaplotypeGenerator::fuse_haplotype_blocks hands to Haplotype (reference
//             src/SeqAlignment/HaplotypeGenerator.cpp:580-607, Haplotype.h:34-50)
//   reads   : Alignment(start, stop, qualities, sequence) + CIGAR (AlignmentData.h:32-60)
//   call    : HapAligner(haplotype, realign_to_hap, INDEL_FLANK_LEN, SWITCH_OLD_ALIGN_LEN, params)
//             .process_reads(alns, 0, &base_quality, realign_read, out_ll, out_seeds)  (HapAligner.h:94-138)
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "longtr_host.h"

namespace {

using namespace ltr;

// The reference objects of one flat locus (addresses are stable: blocks point at the model, the haplotype at the blocks).
struct FlatLocusObjects {
  std::unique_ptr<StutterModel> model;
  std::unique_ptr<HapBlock> left, right;
  std::unique_ptr<RepeatBlock> repeat;
  std::vector<HapBlock*> blocks;
  std::unique_ptr<Haplotype> haplotype;
  std::vector<Alignment> alns;
  std::vector<bool> realign_hap, realign_read;
  std::vector<float> params;
  std::unique_ptr<HapAligner> aligner;
  std::vector<int> seeds;
};

int build_flat_locus(ltr_ctx* ctx, const ltr_flat_locus* L, const int32_t* in_seeds, FlatLocusObjects& o) {
  if (!L || !L->lflank || !L->rflank || !L->alleles || L->n_alleles < 1 || L->n_reads < 0 || !L->motif) return LTR_ERR_INVALID;
  if (L->n_reads > 0 && !L->reads) return LTR_ERR_INVALID;
  if (L->period < 1 || L->indel_flank_len < 0 || L->indel_flank_len > 35) return LTR_ERR_INVALID;
  if (L->indel_flank_len < 5) return LTR_ERR_UNSUPPORTED;  // see job_new (abi.cu)
  if (L->n_aln_params != 0 && L->n_aln_params != 7) return LTR_ERR_INVALID;
  o.model.reset(new StutterModel(L->stutter[0], L->stutter[1], L->stutter[2], L->stutter[3], L->stutter[4], L->stutter[5],
                                 std::string(L->motif)));
  if (!o.model->valid()) return LTR_ERR_INVALID;
  o.model->set_period(L->period);
  const std::string lflank(L->lflank), rflank(L->rflank);
  for (int a = 0; a < L->n_alleles; ++a)
    if (!L->alleles[a]) return LTR_ERR_INVALID;
  o.left.reset(new HapBlock(L->repeat_start - (int32_t)lflank.size(), L->repeat_start, lflank));
  o.repeat.reset(new RepeatBlock(L->repeat_start, L->repeat_end, std::string(L->alleles[0]), L->period, o.model.get()));
  for (int a = 1; a < L->n_alleles; ++a) o.repeat->add_alternate(std::make_pair(std::string(L->alleles[a]), false));
  o.right.reset(new HapBlock(L->repeat_end, L->repeat_end + (int32_t)rflank.size(), rflank));
  o.blocks.push_back(o.left.get());
  o.blocks.push_back(o.repeat.get());
  o.blocks.push_back(o.right.get());
  o.haplotype.reset(new Haplotype(o.blocks));

  o.alns.reserve((size_t)L->n_reads);
  for (int r = 0; r < L->n_reads; ++r) {
    const ltr_flat_read& fr = L->reads[r];
    if (!fr.seq || !fr.qual || !fr.cigar) return LTR_ERR_INVALID;
    o.alns.push_back(Alignment(fr.start, fr.stop, false, false, "read", std::string(fr.qual), std::string(fr.seq),
                               std::string(fr.seq)));
    if (!o.alns.back().set_cigar_string(fr.cigar)) return LTR_ERR_INVALID;
  }
  o.realign_hap.assign((size_t)L->n_alleles, true);
  o.realign_read.assign((size_t)L->n_reads, true);
  if (L->realign_to_hap)
    for (int a = 0; a < L->n_alleles; ++a) o.realign_hap[a] = L->realign_to_hap[a] != 0;
  if (L->realign_read)
    for (int r = 0; r < L->n_reads; ++r) o.realign_read[r] = L->realign_read[r] != 0;
  o.params.assign(L->aln_params, L->aln_params + L->n_aln_params);
  o.aligner.reset(new HapAligner(o.haplotype.get(), o.realign_hap, L->indel_flank_len, L->switch_old_align_len, o.params, ctx));
  o.seeds.resize((size_t)L->n_reads);
  for (int r = 0; r < L->n_reads; ++r) o.seeds[r] = in_seeds[r];
  return LTR_OK;
}

bool same_params(const ltr_params& a, const ltr_params& b) { return memcmp(&a, &b, sizeof(ltr_params)) == 0; }

}  // namespace

extern "C" int ltr_process_reads_flat(ltr_ctx* ctx, const ltr_flat_locus* L, double* out_ll, int32_t* out_seeds) {
  if (!ctx || !L || !out_ll || !out_seeds) return LTR_ERR_INVALID;
  FlatLocusObjects o;
  const int rc = build_flat_locus(ctx, L, out_seeds, o);
  if (rc != LTR_OK) return rc;
  BaseQuality base_quality;
  o.aligner->process_reads(o.alns, 0, &base_quality, o.realign_read, out_ll, o.seeds.data());
  if (o.aligner->status() != LTR_OK) return o.aligner->status();
  for (int r = 0; r < L->n_reads; ++r) out_seeds[r] = o.seeds[r];
  return LTR_OK;
}

// HapAligner::process_reads for many flat loci at once: the long-path loci that share their alignment parameters are
// flattened into ONE ltr_viterbi_batch (one plan, one upload, one set of kernel launches) and their log-likelihoods
// scattered back into each locus' aln_probs; loci on the homopolymer path go through ltr_process_reads_flat one by one.
// This is the form a host that keeps many regions open calls.
extern "C" int ltr_process_reads_flat_batch(ltr_ctx* ctx, int32_t n_loci, const ltr_flat_locus* loci, double* const* out_ll,
                                            int32_t* const* out_seeds) {
  if (!ctx || n_loci < 0 || (n_loci > 0 && (!loci || !out_ll || !out_seeds))) return LTR_ERR_INVALID;
  std::vector<FlatLocusObjects> objs((size_t)n_loci);
  std::vector<HapAligner::LongPart> parts((size_t)n_loci);
  std::vector<std::string> hap_bytes((size_t)n_loci), read_bytes((size_t)n_loci);
  std::vector<std::vector<uint32_t> > hap_off((size_t)n_loci), read_off((size_t)n_loci);
  std::vector<ltr_params> params((size_t)n_loci);
  std::vector<int> state((size_t)n_loci, 0);  // 0 = long path prepared, 1 = short path, 2 = nothing to align, < 0 = error code
  // ---- build the reference objects and flatten every locus (host threads) ---------------------------------------
  {
    unsigned n_threads = std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
    if (const char* env = getenv("LTR_PLAN_THREADS")) n_threads = (unsigned)std::max(1, atoi(env));
    if ((unsigned)n_loci < 4 * n_threads) n_threads = 1;
    auto work = [&](int l0, int l1) {
      for (int l = l0; l < l1; ++l) {
        if (!out_ll[l] || !out_seeds[l]) { state[(size_t)l] = LTR_ERR_INVALID; continue; }
        int rc = build_flat_locus(ctx, &loci[l], out_seeds[l], objs[(size_t)l]);
        if (rc != LTR_OK) { state[(size_t)l] = rc; continue; }
        HapAligner& al = *objs[(size_t)l].aligner;
        if (al.uses_short_path()) { state[(size_t)l] = 1; continue; }
        hap_off[(size_t)l].assign(1, 0);
        read_off[(size_t)l].assign(1, 0);
        if (!al.prepare_long(objs[(size_t)l].alns, 0, objs[(size_t)l].realign_read, objs[(size_t)l].seeds.data(), parts[(size_t)l],
                             hap_bytes[(size_t)l], hap_off[(size_t)l], read_bytes[(size_t)l], read_off[(size_t)l])) {
          state[(size_t)l] = al.status() < 0 ? al.status() : LTR_ERR_INVALID;
          continue;
        }
        al.fill_params(params[(size_t)l]);
        if (parts[(size_t)l].hap_cols.empty() || parts[(size_t)l].read_rows.empty()) state[(size_t)l] = 2;
      }
    };
    if (n_threads <= 1) {
      work(0, n_loci);
    } else {
      std::vector<std::thread> th;
      const int chunk = (n_loci + (int)n_threads - 1) / (int)n_threads;
      for (unsigned t = 0; t < n_threads; ++t) {
        const int l0 = std::min(n_loci, (int)t * chunk), l1 = std::min(n_loci, l0 + chunk);
        if (l0 < l1) th.emplace_back(work, l0, l1);
      }
      for (std::thread& x : th) x.join();
    }
  }
  for (int l = 0; l < n_loci; ++l)
    if (state[(size_t)l] < 0) return state[(size_t)l];  // error codes are negative
  // ---- one job per distinct parameter set ------------------------------------------------------------------------
  std::vector<char> done((size_t)n_loci, 0);
  for (int first = 0; first < n_loci; ++first) {
    if (done[(size_t)first] || state[(size_t)first] != 0) continue;
    std::vector<int> group;
    for (int l = first; l < n_loci; ++l)
      if (!done[(size_t)l] && state[(size_t)l] == 0 && same_params(params[(size_t)l], params[(size_t)first])) {
        group.push_back(l);
        done[(size_t)l] = 1;
      }
    std::vector<uint32_t> lhb(1, 0), lrb(1, 0), hoff(1, 0), roff(1, 0);
    std::string hb, rb;
    size_t n_ll = 0;
    for (int l : group) {
      const uint32_t hbase = (uint32_t)hb.size(), rbase = (uint32_t)rb.size();
      hb += hap_bytes[(size_t)l];
      rb += read_bytes[(size_t)l];
      for (size_t i = 1; i < hap_off[(size_t)l].size(); ++i) hoff.push_back(hbase + hap_off[(size_t)l][i]);
      for (size_t i = 1; i < read_off[(size_t)l].size(); ++i) roff.push_back(rbase + read_off[(size_t)l][i]);
      lhb.push_back((uint32_t)hoff.size() - 1);
      lrb.push_back((uint32_t)roff.size() - 1);
      n_ll += parts[(size_t)l].hap_cols.size() * parts[(size_t)l].read_rows.size();
    }
    if (hb.size() > 0xFFFFFFF0ull || rb.size() > 0xFFFFFFF0ull) return LTR_ERR_INVALID;
    ltr_viterbi_batch b;
    b.n_loci = (uint32_t)group.size();
    b.locus_hap_begin = lhb.data();
    b.locus_read_begin = lrb.data();
    b.hap_off = hoff.data();
    b.hap_bytes = reinterpret_cast<const uint8_t*>(hb.data());
    b.read_off = roff.data();
    b.read_bytes = reinterpret_cast<const uint8_t*>(rb.data());
    std::vector<double> ll(n_ll);
    const int rc = ltr_viterbi_ll(ctx, &params[(size_t)first], &b, ll.data(), NULL);
    if (rc != LTR_OK) return rc;
    size_t pos = 0;
    for (int l : group) {
      objs[(size_t)l].aligner->scatter_long(parts[(size_t)l], ll.data() + pos, 0, out_ll[l]);
      pos += parts[(size_t)l].hap_cols.size() * parts[(size_t)l].read_rows.size();
    }
  }
  // ---- homopolymer-path loci, seeds ---------------------------------------------------------------------------------
  BaseQuality base_quality;
  for (int l = 0; l < n_loci; ++l) {
    FlatLocusObjects& o = objs[(size_t)l];
    if (state[(size_t)l] == 1) {
      o.aligner->process_reads(o.alns, 0, &base_quality, o.realign_read, out_ll[l], o.seeds.data());
      if (o.aligner->status() != LTR_OK) return o.aligner->status();
    }
    for (int r = 0; r < loci[l].n_reads; ++r) out_seeds[l][r] = o.seeds[(size_t)r];
  }
  return LTR_OK;
}

// ---- ltr_flatten_loci: the flattening half of the batch call, on its own (host only, no GPU) ------------------------
namespace {
struct FlatBatchOwner {
  ltr_flat_batch pub;
  std::vector<uint32_t> lhb, lrb, hoff, roff;
  std::string hb, rb;
  std::vector<int32_t> hap_col, read_row;
};
}  // namespace

extern "C" int ltr_flatten_loci(int32_t n_loci, const ltr_flat_locus* loci, ltr_params* params_out, ltr_flat_batch** out) {
  if (!out || n_loci < 0 || (n_loci > 0 && !loci)) return LTR_ERR_INVALID;
  *out = nullptr;
  std::unique_ptr<FlatBatchOwner> B(new FlatBatchOwner());
  B->lhb.assign(1, 0);
  B->lrb.assign(1, 0);
  B->hoff.assign(1, 0);
  B->roff.assign(1, 0);
  ltr_params first_params;
  memset(&first_params, 0, sizeof(first_params));
  for (int l = 0; l < n_loci; ++l) {
    FlatLocusObjects o;
    std::vector<int32_t> no_seeds((size_t)std::max(0, loci[l].n_reads), -1);
    const int rc = build_flat_locus(nullptr, &loci[l], no_seeds.data(), o);
    if (rc != LTR_OK) return rc;
    if (o.aligner->uses_short_path()) return LTR_ERR_UNSUPPORTED;  // homopolymer path: ltr_stutter_ll has its own batch form
    HapAligner::LongPart part;
    std::vector<uint32_t> hap_off(1, 0), read_off(1, 0);
    std::string hap_bytes, read_bytes;
    if (!o.aligner->prepare_long(o.alns, 0, o.realign_read, o.seeds.data(), part, hap_bytes, hap_off, read_bytes, read_off))
      return o.aligner->status() < 0 ? o.aligner->status() : LTR_ERR_INVALID;
    ltr_params p;
    o.aligner->fill_params(p);
    if (l == 0) first_params = p;
    else if (!same_params(p, first_params)) return LTR_ERR_UNSUPPORTED;  // one job = one parameter set
    const uint32_t hbase = (uint32_t)B->hb.size(), rbase = (uint32_t)B->rb.size();
    if ((uint64_t)hbase + hap_bytes.size() > 0xFFFFFFF0ull || (uint64_t)rbase + read_bytes.size() > 0xFFFFFFF0ull) return LTR_ERR_INVALID;
    B->hb += hap_bytes;
    B->rb += read_bytes;
    for (size_t i = 1; i < hap_off.size(); ++i) B->hoff.push_back(hbase + hap_off[i]);
    for (size_t i = 1; i < read_off.size(); ++i) B->roff.push_back(rbase + read_off[i]);
    for (int c : part.hap_cols) B->hap_col.push_back(c);
    for (int r : part.read_rows) B->read_row.push_back(r);
    B->lhb.push_back((uint32_t)B->hoff.size() - 1);
    B->lrb.push_back((uint32_t)B->roff.size() - 1);
  }
  if (params_out) {
    if (n_loci > 0) *params_out = first_params;
    else ltr_params_default(params_out);
  }
  B->pub.vit.n_loci = (uint32_t)n_loci;
  B->pub.vit.locus_hap_begin = B->lhb.data();
  B->pub.vit.locus_read_begin = B->lrb.data();
  B->pub.vit.hap_off = B->hoff.data();
  B->pub.vit.hap_bytes = reinterpret_cast<const uint8_t*>(B->hb.data());
  B->pub.vit.read_off = B->roff.data();
  B->pub.vit.read_bytes = reinterpret_cast<const uint8_t*>(B->rb.data());
  B->pub.hap_col = B->hap_col.data();
  B->pub.read_row = B->read_row.data();
  B->pub.n_haps = (uint32_t)B->hap_col.size();
  B->pub.n_reads = (uint32_t)B->read_row.size();
  *out = &B.release()->pub;  // pub is the first member: the owner is recovered from it in ltr_flat_batch_free
  return LTR_OK;
}

extern "C" void ltr_flat_batch_free(ltr_flat_batch* b) {
  if (b) delete reinterpret_cast<FlatBatchOwner*>(b);
}

// ReadPooler + BaseQuality::median_base_qualities on plain strings (host only): pool index per read, number of pools and
// the pooled reads' median qualities back to back in pool order.  Returns the bytes written, negative = error.
extern "C" int64_t ltr_pool_reads(int32_t n_reads, const char* const* seqs, const char* const* quals, int32_t* pool_index,
                                  int32_t* n_pools, char* pooled_quals, int64_t cap) {
  using namespace ltr;
  if (n_reads < 0 || !n_pools || (n_reads > 0 && (!seqs || !quals || !pool_index))) return LTR_ERR_INVALID;
  ReadPooler pooler;
  BaseQuality bq;
  for (int32_t r = 0; r < n_reads; ++r) {
    if (!seqs[r] || !quals[r] || strlen(seqs[r]) != strlen(quals[r])) return LTR_ERR_INVALID;
    const Alignment aln(100, 100 + (int32_t)strlen(seqs[r]) - 1, false, false, "read", std::string(quals[r]), std::string(seqs[r]),
                        std::string(seqs[r]));
    pool_index[r] = pooler.add_alignment(aln);
  }
  if (!pooler.pool(bq)) return LTR_ERR_INVALID;
  *n_pools = pooler.num_pools();
  int64_t off = 0;
  std::vector<Alignment>& alns = pooler.get_alignments();
  for (size_t i = 0; i < alns.size(); ++i) {
    const std::string& q = alns[i].get_base_qualities();
    if (pooled_quals) {
      if (off + (int64_t)q.size() > cap) return LTR_ERR_INVALID;
      memcpy(pooled_quals + off, q.data(), q.size());
    }
    off += (int64_t)q.size();
  }
  return off;
}
