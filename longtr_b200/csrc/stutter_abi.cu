// stutter_abi.cu -- ltr_stutter_ll: host side of the homopolymer / --stutter-align-len path.
//
// Validates the batch, tabulates what the reference computes with libm on the host (INT_LOGS, BaseQuality
// tables, log_prob_pcr_artifact per allele -- src/mathops.cpp:14-22, src/base_quality.h:29-38,
// RepeatStutterInfo.h:53-61 / stutter_model.cpp:29-53) so that the device sees bit-identical constants, builds
// the (read, allele) task list and launches stutter_pair_kernel.  No CPU fallback.
#include <cuda_runtime.h>
#include <math.h>

#include <algorithm>
#include <vector>

#include "ctx.h"
#include "host/longtr_host.h"
#include "kernels.h"
#include "stutter_core.cuh"

using namespace ltr;

namespace {

template <typename T>
int up(ltr_ctx* ctx, DeviceBuffer& buf, const T* src, size_t count, uint64_t* h2d) {
  const size_t bytes = count * sizeof(T);
  LTR_CUDA(ctx, buf.alloc(bytes + 16));
  if (bytes) LTR_CUDA(ctx, cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, ctx->main_stream));
  if (h2d) *h2d += bytes;
  return LTR_OK;
}

struct Buffers {
  DeviceBuffer tasks, read_off, read_bytes, qual_bytes, seed, lf_off, lf, rf_off, rf, al_off, al, art, out, int_logs,
      qlc, qlw;
  ~Buffers() {
    DeviceBuffer* all[] = {&tasks, &read_off, &read_bytes, &qual_bytes, &seed, &lf_off, &lf, &rf_off, &rf, &al_off, &al,
                           &art, &out, &int_logs, &qlc, &qlw};
    for (DeviceBuffer* b : all) b->free();
  }
};

}  // namespace

extern "C" int ltr_stutter_ll_status(ltr_ctx* ctx, const ltr_params* params, const ltr_stutter_batch* b, double* out_ll,
                                     int32_t* locus_status, ltr_job_stats* stats);

// The whole batch or nothing: the first locus that cannot be processed fails the call (nothing is computed).
extern "C" int ltr_stutter_ll(ltr_ctx* ctx, const ltr_params* params, const ltr_stutter_batch* b, double* out_ll,
                              ltr_job_stats* stats) {
  return ltr_stutter_ll_status(ctx, params, b, out_ll, NULL, stats);
}

// locus_status != NULL: a locus that cannot be processed (empty allele, flanks too long for the shared-memory staging,
// malformed stutter model / seeds) gets its error code there, its rows stay untouched, and the other loci are computed.
extern "C" int ltr_stutter_ll_status(ltr_ctx* ctx, const ltr_params* params, const ltr_stutter_batch* b, double* out_ll,
                                     int32_t* locus_status, ltr_job_stats* stats) {
  if (!ctx || !params || !b || !out_ll) return LTR_ERR_INVALID;
  if (stats) memset(stats, 0, sizeof(*stats));
  if (b->n_loci == 0) return LTR_OK;
  if (!b->locus_allele_begin || !b->locus_read_begin || !b->lflank_off || !b->rflank_off || !b->allele_off ||
      !b->stutter || !b->motif_len || !b->read_off || !b->read_seed)
    return LTR_ERR_INVALID;
  LTR_CUDA(ctx, cudaSetDevice(ctx->device));
  AllocScope alloc_scope(ctx->main_stream);
  const uint32_t n_loci = b->n_loci;
  const uint32_t n_alleles = b->locus_allele_begin[n_loci], n_reads = b->locus_read_begin[n_loci];

  // ---- plan: output offsets, tasks, zero rows, size limits -------------------------------------------------------
  std::vector<unsigned long long> ll_off((size_t)n_loci + 1, 0);
  std::vector<StutterTask> tasks;
  std::vector<double> art((size_t)n_alleles * 13);
  uint32_t max_flank = 1, max_block = 1, max_hap = 2;
  uint64_t n_pairs = 0, n_cells = 0;
  // validation first (offsets must be monotone before anything is sized), then one pass per locus that either commits
  // the locus' tasks or records why it cannot be processed
  for (uint32_t l = 0; l < n_loci; ++l) {
    if (b->locus_allele_begin[l + 1] < b->locus_allele_begin[l] || b->locus_read_begin[l + 1] < b->locus_read_begin[l])
      return LTR_ERR_INVALID;
    const uint32_t H = b->locus_allele_begin[l + 1] - b->locus_allele_begin[l];
    ll_off[l + 1] = ll_off[l] + (unsigned long long)H * (b->locus_read_begin[l + 1] - b->locus_read_begin[l]);
    if (locus_status) locus_status[l] = LTR_OK;
  }
  std::vector<uint32_t> zero_rows;  // reads without a seed: LL 0 for every haplotype (written once validation is through)
  for (uint32_t l = 0; l < n_loci; ++l) {
    const uint32_t a0 = b->locus_allele_begin[l], a1 = b->locus_allele_begin[l + 1];
    const uint32_t r0 = b->locus_read_begin[l], r1 = b->locus_read_begin[l + 1];
    const uint32_t H = a1 - a0;
    const uint32_t n0 = b->lflank_off[l + 1] - b->lflank_off[l], n2 = b->rflank_off[l + 1] - b->rflank_off[l];
    int st = LTR_OK;
    uint32_t l_flank = 1, l_block = 1, l_hap = 2;
    if (b->lflank_off[l + 1] < b->lflank_off[l] || b->rflank_off[l + 1] < b->rflank_off[l] || n0 < 1 || n2 < 1 ||
        b->motif_len[l] < 1)
      st = LTR_ERR_INVALID;
    StutterModel model(b->stutter[6 * l], b->stutter[6 * l + 1], b->stutter[6 * l + 2], b->stutter[6 * l + 3],
                       b->stutter[6 * l + 4], b->stutter[6 * l + 5],
                       std::string((size_t)std::max(1, b->motif_len[l]), 'N'));
    if (st == LTR_OK && !model.valid()) st = LTR_ERR_INVALID;
    for (uint32_t a = a0; a < a1 && st == LTR_OK; ++a) {
      if (b->allele_off[a + 1] < b->allele_off[a]) st = LTR_ERR_INVALID;
      // empty allele (<DEL>): the reference's stutter row would overwrite its predecessor
      else if (b->allele_off[a + 1] == b->allele_off[a]) st = LTR_ERR_UNSUPPORTED;
      else {
        const uint32_t B = b->allele_off[a + 1] - b->allele_off[a];
        l_block = std::max(l_block, B);
        l_hap = std::max(l_hap, n0 + B + n2);
      }
    }
    for (uint32_t r = r0; r < r1 && st == LTR_OK; ++r) {
      if (b->realign_read && !b->realign_read[r]) continue;
      if (b->read_off[r + 1] < b->read_off[r]) { st = LTR_ERR_INVALID; break; }
      const int32_t N = (int32_t)(b->read_off[r + 1] - b->read_off[r]);
      const int32_t seed = b->read_seed[r];
      if (seed < 0) continue;
      if (seed < 1 || seed >= N - 1) { st = LTR_ERR_INVALID; break; }
      l_flank = std::max<uint32_t>(l_flank, (uint32_t)std::max(seed, N - seed - 1));
    }
    // read flank / allele too long for the shared-memory staging of the kernel
    if (st == LTR_OK && stutter_block_smem_bytes(l_flank, l_block, l_hap) > 220 * 1024) st = LTR_ERR_UNSUPPORTED;
    if (st != LTR_OK) {
      if (!locus_status) return st;
      locus_status[l] = st;
      continue;
    }
    max_flank = std::max(max_flank, l_flank);
    max_block = std::max(max_block, l_block);
    max_hap = std::max(max_hap, l_hap);
    for (uint32_t a = a0; a < a1; ++a) {
      const uint32_t B = b->allele_off[a + 1] - b->allele_off[a];
      // RepeatStutterInfo(period = 1, ...): artifacts of -6..+6 bases (RepeatStutterInfo.h:10-11, 53-61)
      RepeatStutterInfo info(1, std::string((size_t)B, 'N'), model);
      for (int D = -6; D <= 6; ++D) art[(size_t)a * 13 + (D + 6)] = info.log_prob_pcr_artifact(0, D);
    }
    for (uint32_t r = r0; r < r1; ++r) {
      if (b->realign_read && !b->realign_read[r]) continue;
      const int32_t N = (int32_t)(b->read_off[r + 1] - b->read_off[r]);
      if (b->read_seed[r] < 0) {  // HapAligner.cpp:570-574: no seed -> LL 0 for every haplotype of the read
        zero_rows.push_back(r);
        continue;
      }
      for (uint32_t a = a0; a < a1; ++a) {
        if (b->realign_allele && !b->realign_allele[a]) continue;
        StutterTask t;
        t.locus = l;
        t.read = r;
        t.allele = a;
        t.out_index = ll_off[l] + (unsigned long long)(r - r0) * H + (a - a0);
        tasks.push_back(t);
        n_pairs++;
        n_cells += stutter_pair_cells((int32_t)(n0 + n2), (int32_t)(b->allele_off[a + 1] - b->allele_off[a]), N);
      }
    }
  }
  // nothing of the caller's array was touched up to here (a call that fails as a whole leaves out_ll as it was)
  {
    std::vector<uint32_t> read_locus;
    if (!zero_rows.empty()) {
      read_locus.resize(n_reads);
      for (uint32_t l = 0; l < n_loci; ++l)
        for (uint32_t r = b->locus_read_begin[l]; r < b->locus_read_begin[l + 1]; ++r) read_locus[r] = l;
    }
    for (uint32_t r : zero_rows) {
      const uint32_t l = read_locus[r];
      const uint32_t H = b->locus_allele_begin[l + 1] - b->locus_allele_begin[l];
      double* row = out_ll + ll_off[l] + (size_t)(r - b->locus_read_begin[l]) * H;
      for (uint32_t h = 0; h < H; ++h) row[h] = 0.0;
    }
  }
  if (stats) {
    stats->n_pairs = n_pairs;
    stats->n_cells = n_cells;
  }
  if (tasks.empty()) return LTR_OK;

  // ---- constants from the host's libm ---------------------------------------------------------------------------------
  std::vector<double> int_logs((size_t)std::max(max_block, max_hap) + 16);
  for (size_t i = 0; i < int_logs.size(); ++i) int_logs[i] = int_log((int)i);
  double qlc[256], qlw[256];
  BaseQuality().byte_tables(qlc, qlw);

  Buffers d;
  uint64_t h2d = 0;
  int rc = LTR_OK;
#define UP(buf, ptr, n) if (rc == LTR_OK) rc = up(ctx, buf, ptr, n, &h2d)
  UP(d.tasks, tasks.data(), tasks.size());
  UP(d.read_off, b->read_off, (size_t)n_reads + 1);
  UP(d.read_bytes, b->read_bytes, (size_t)b->read_off[n_reads]);
  UP(d.qual_bytes, b->qual_bytes, (size_t)b->read_off[n_reads]);
  UP(d.seed, b->read_seed, (size_t)n_reads);
  UP(d.lf_off, b->lflank_off, (size_t)n_loci + 1);
  UP(d.lf, b->lflank_bytes, (size_t)b->lflank_off[n_loci]);
  UP(d.rf_off, b->rflank_off, (size_t)n_loci + 1);
  UP(d.rf, b->rflank_bytes, (size_t)b->rflank_off[n_loci]);
  UP(d.al_off, b->allele_off, (size_t)n_alleles + 1);
  UP(d.al, b->allele_bytes, (size_t)b->allele_off[n_alleles]);
  UP(d.art, art.data(), art.size());
  UP(d.int_logs, int_logs.data(), int_logs.size());
  UP(d.qlc, qlc, 256);
  UP(d.qlw, qlw, 256);
#undef UP
  if (rc != LTR_OK) return rc;
  const size_t n_ll = (size_t)ll_off[n_loci];
  LTR_CUDA(ctx, d.out.alloc(n_ll * sizeof(double)));
  // rows / columns that are not realigned must stay untouched: start from the caller's array
  LTR_CUDA(ctx, cudaMemcpyAsync(d.out.p, out_ll, n_ll * sizeof(double), cudaMemcpyHostToDevice, ctx->main_stream));
  h2d += n_ll * sizeof(double);

  StutConsts C;
  C.i2i = (double)params->ins_ins;
  C.i2m = (double)params->ins_match;
  C.d2d = (double)params->del_del;
  C.d2m = (double)params->del_match;
  C.m2m = (double)params->match_match;
  C.m2i = (double)params->match_ins;
  C.m2d = (double)params->match_del;
  C.log_thresh = log(0.001);
  C.int_logs = d.int_logs.as<double>();
  C.qual_lc = d.qlc.as<double>();
  C.qual_lw = d.qlw.as<double>();
  StutterDevBatch B;
  B.n_tasks = (uint32_t)tasks.size();
  B.tasks = d.tasks.as<StutterTask>();
  B.read_off = d.read_off.as<uint32_t>();
  B.read_bytes = d.read_bytes.as<uint8_t>();
  B.qual_bytes = d.qual_bytes.as<uint8_t>();
  B.read_seed = d.seed.as<int32_t>();
  B.lflank_off = d.lf_off.as<uint32_t>();
  B.lflank_bytes = d.lf.as<uint8_t>();
  B.rflank_off = d.rf_off.as<uint32_t>();
  B.rflank_bytes = d.rf.as<uint8_t>();
  B.allele_off = d.al_off.as<uint32_t>();
  B.allele_bytes = d.al.as<uint8_t>();
  B.allele_artifact_lp = d.art.as<double>();
  B.out_ll = d.out.as<double>();
  B.max_flank = max_flank;
  B.max_block = max_block;
  B.max_hap = max_hap;
  LTR_CUDA(ctx, cudaEventRecord(ctx->ev_start, ctx->main_stream));
  LTR_CUDA(ctx, launch_stutter(C, B, ctx->main_stream));
  LTR_CUDA(ctx, cudaEventRecord(ctx->ev_end, ctx->main_stream));
  LTR_CUDA(ctx, cudaMemcpyAsync(out_ll, d.out.p, n_ll * sizeof(double), cudaMemcpyDeviceToHost, ctx->main_stream));
  LTR_CUDA(ctx, cudaStreamSynchronize(ctx->main_stream));
  if (stats) {
    float ms = 0.f;
    LTR_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_end));
    stats->kernel_ms = ms;
    stats->viterbi_ms = ms;
    stats->n_launches = 1;
    stats->h2d_bytes = h2d;
    stats->d2h_bytes = n_ll * sizeof(double);
  }
  return LTR_OK;
}
