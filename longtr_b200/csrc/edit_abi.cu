// edit_abi.cu -- ltr_edit_distances / ltr_cluster_greedy: host side of the candidate-haplotype clustering kernels
// (HaplotypeGenerator::needleman_wunsch / greedy_clustering, reference src/SeqAlignment/HaplotypeGenerator.cpp:201-271).
// Validates, uploads, launches edit_kernel.cu in stream order, downloads.  No CPU fallback.
#include <cuda_runtime.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "ctx.h"
#include "edit_core.cuh"
#include "kernels.h"

using namespace ltr;

namespace {

struct Pool {  // device buffers of one call, freed in stream order on every exit path
  std::vector<DeviceBuffer*> all;
  ~Pool() {
    for (DeviceBuffer* b : all) {
      b->free();
      delete b;
    }
  }
  DeviceBuffer* get() {
    all.push_back(new DeviceBuffer());
    return all.back();
  }
};

template <typename T>
int upload(ltr_ctx* ctx, Pool& pool, const T* src, size_t count, T** dev, uint64_t* h2d) {
  DeviceBuffer* b = pool.get();
  const size_t bytes = count * sizeof(T);
  LTR_CUDA(ctx, b->alloc(bytes + 16));
  if (bytes) LTR_CUDA(ctx, cudaMemcpyAsync(b->p, src, bytes, cudaMemcpyHostToDevice, ctx->main_stream));
  if (h2d) *h2d += bytes;
  *dev = b->as<T>();
  return LTR_OK;
}

template <typename T>
int scratch(ltr_ctx* ctx, Pool& pool, size_t count, T** dev) {
  DeviceBuffer* b = pool.get();
  LTR_CUDA(ctx, b->alloc(count * sizeof(T) + 16));
  *dev = b->as<T>();
  return LTR_OK;
}

// seq_off monotone; returns the longest sequence through *max_len
int check_seqs(const uint32_t* seq_off, uint32_t n_seqs, uint32_t* max_len) {
  uint32_t mx = 0;
  for (uint32_t s = 0; s < n_seqs; ++s) {
    if (seq_off[s + 1] < seq_off[s]) return LTR_ERR_INVALID;
    mx = std::max(mx, seq_off[s + 1] - seq_off[s]);
  }
  *max_len = mx;
  return LTR_OK;
}

}  // namespace

extern "C" int ltr_edit_distances(ltr_ctx* ctx, const uint8_t* seq_bytes, const uint32_t* seq_off, uint32_t n_seqs,
                                  const uint32_t* pair_a, const uint32_t* pair_b, const int32_t* pair_T, uint32_t n_pairs,
                                  int32_t* out_score, ltr_job_stats* stats) {
  if (!ctx) return LTR_ERR_INVALID;
  if (stats) memset(stats, 0, sizeof(*stats));
  if (n_pairs == 0) return LTR_OK;
  if (!seq_off || !pair_a || !pair_b || !pair_T || !out_score || n_seqs == 0) return LTR_ERR_INVALID;
  uint32_t max_len = 0;
  if (check_seqs(seq_off, n_seqs, &max_len) != LTR_OK) return LTR_ERR_INVALID;
  if (seq_off[n_seqs] > 0 && !seq_bytes) return LTR_ERR_INVALID;
  uint64_t n_cells = 0;
  for (uint32_t p = 0; p < n_pairs; ++p) {
    if (pair_a[p] >= n_seqs || pair_b[p] >= n_seqs || pair_T[p] < 0 || pair_T[p] > kEditMaxThreshold) return LTR_ERR_INVALID;
    n_cells += (uint64_t)(seq_off[pair_a[p] + 1] - seq_off[pair_a[p]]) * (seq_off[pair_b[p] + 1] - seq_off[pair_b[p]]);
  }
  LTR_CUDA(ctx, cudaSetDevice(ctx->device));
  AllocScope alloc_scope(ctx->main_stream);
  cudaStream_t st = ctx->main_stream;
  Pool pool;
  uint64_t h2d = 0;
  EditArgs A;
  memset(&A, 0, sizeof(A));
  uint8_t* d_bytes = nullptr;
  uint32_t *d_off = nullptr, *d_a = nullptr, *d_b = nullptr, *d_words = nullptr;
  int32_t* d_T = nullptr;
  int rc = upload(ctx, pool, seq_bytes, (size_t)seq_off[n_seqs], &d_bytes, &h2d);
  if (rc == LTR_OK) rc = upload(ctx, pool, seq_off, (size_t)n_seqs + 1, &d_off, &h2d);
  if (rc == LTR_OK) rc = upload(ctx, pool, pair_a, (size_t)n_pairs, &d_a, &h2d);
  if (rc == LTR_OK) rc = upload(ctx, pool, pair_b, (size_t)n_pairs, &d_b, &h2d);
  if (rc == LTR_OK) rc = upload(ctx, pool, pair_T, (size_t)n_pairs, &d_T, &h2d);
  if (rc == LTR_OK) rc = scratch(ctx, pool, (size_t)n_pairs, &A.out);
  if (rc == LTR_OK) rc = scratch(ctx, pool, (size_t)n_pairs, &A.flagged);
  if (rc == LTR_OK) rc = scratch(ctx, pool, 4, &d_words);  // [0] Myers cursor, [1] exact cursor, [2] flagged pairs
  A.line_stride = (max_len + 4 + 15) & ~15u;
  if (rc == LTR_OK && max_len > (uint32_t)kEditStripRows)
    rc = scratch(ctx, pool, (size_t)edit_myers_warps(ctx->sm_count) * A.line_stride, &A.lines);
  if (rc == LTR_OK && max_len > 32u * kEditDpRows)
    rc = scratch(ctx, pool, (size_t)edit_exact_warps(ctx->sm_count) * A.line_stride, &A.dp_lines);
  if (rc != LTR_OK) return rc;
  A.seq_bytes = d_bytes;
  A.seq_off = d_off;
  A.pair_a = d_a;
  A.pair_b = d_b;
  A.pair_T = d_T;
  A.n_pairs = n_pairs;
  A.n_flagged = d_words + 2;
  LTR_CUDA(ctx, cudaMemsetAsync(d_words, 0, 4 * sizeof(uint32_t), st));
  LTR_CUDA(ctx, cudaEventRecord(ctx->ev_start, st));
  A.cursor = d_words;
  LTR_CUDA(ctx, launch_edit_myers(A, n_pairs, ctx->sm_count, st));
  A.cursor = d_words + 1;
  LTR_CUDA(ctx, launch_edit_exact(A, n_pairs, ctx->sm_count, st));
  LTR_CUDA(ctx, cudaEventRecord(ctx->ev_end, st));
  uint32_t n_flagged = 0;
  LTR_CUDA(ctx, cudaMemcpyAsync(out_score, A.out, (size_t)n_pairs * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  LTR_CUDA(ctx, cudaMemcpyAsync(&n_flagged, d_words + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  LTR_CUDA(ctx, cudaStreamSynchronize(st));
  if (stats) {
    float ms = 0.f;
    LTR_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_end));
    stats->n_pairs = n_pairs;
    stats->n_cells = n_cells;
    stats->n_fallback = n_flagged;
    stats->kernel_ms = ms;
    stats->n_launches = 2;
    stats->h2d_bytes = h2d;
    stats->d2h_bytes = (uint64_t)n_pairs * sizeof(int32_t);
  }
  return LTR_OK;
}

extern "C" int ltr_cluster_greedy(ltr_ctx* ctx, const uint8_t* seq_bytes, const uint32_t* seq_off, uint32_t n_seqs,
                                  const uint32_t* set_begin, const uint32_t* set_items, const int32_t* set_T,
                                  uint32_t n_sets, int32_t* out_centroid_of, int32_t* out_n_centroids, uint8_t* out_ok,
                                  ltr_job_stats* stats) {
  if (!ctx) return LTR_ERR_INVALID;
  if (stats) memset(stats, 0, sizeof(*stats));
  if (n_sets == 0) return LTR_OK;
  if (!seq_off || !set_begin || !set_T || !out_centroid_of || !out_n_centroids || !out_ok) return LTR_ERR_INVALID;
  uint32_t max_len = 0;
  if (n_seqs && check_seqs(seq_off, n_seqs, &max_len) != LTR_OK) return LTR_ERR_INVALID;
  if (n_seqs && seq_off[n_seqs] > 0 && !seq_bytes) return LTR_ERR_INVALID;
  for (uint32_t k = 0; k < n_sets; ++k)
    if (set_begin[k + 1] < set_begin[k] || set_T[k] < 0 || set_T[k] > kEditMaxThreshold) return LTR_ERR_INVALID;
  const uint32_t n_items = set_begin[n_sets];
  if (n_items && !set_items) return LTR_ERR_INVALID;
  std::vector<uint32_t> item_set(n_items);
  for (uint32_t k = 0; k < n_sets; ++k)
    for (uint32_t i = set_begin[k]; i < set_begin[k + 1]; ++i) {
      if (set_items[i] >= n_seqs) return LTR_ERR_INVALID;
      item_set[i] = k;
    }
  if (n_items == 0) {
    for (uint32_t k = 0; k < n_sets; ++k) {
      out_n_centroids[k] = 0;
      out_ok[k] = 1;
    }
    return LTR_OK;
  }
  LTR_CUDA(ctx, cudaSetDevice(ctx->device));
  AllocScope alloc_scope(ctx->main_stream);
  cudaStream_t st = ctx->main_stream;
  Pool pool;
  uint64_t h2d = 0;
  EditArgs A;
  ClusterDev C;
  memset(&A, 0, sizeof(A));
  memset(&C, 0, sizeof(C));
  uint8_t* d_bytes = nullptr;
  uint32_t *d_off = nullptr, *d_set_begin = nullptr, *d_item_set = nullptr, *d_item_seq = nullptr, *d_words = nullptr;
  int32_t* d_set_T = nullptr;
  int rc = upload(ctx, pool, seq_bytes, (size_t)seq_off[n_seqs], &d_bytes, &h2d);
  if (rc == LTR_OK) rc = upload(ctx, pool, seq_off, (size_t)n_seqs + 1, &d_off, &h2d);
  if (rc == LTR_OK) rc = upload(ctx, pool, set_begin, (size_t)n_sets + 1, &d_set_begin, &h2d);
  if (rc == LTR_OK) rc = upload(ctx, pool, set_T, (size_t)n_sets, &d_set_T, &h2d);
  if (rc == LTR_OK) rc = upload(ctx, pool, item_set.data(), (size_t)n_items, &d_item_set, &h2d);
  if (rc == LTR_OK) rc = upload(ctx, pool, set_items, (size_t)n_items, &d_item_seq, &h2d);
  if (rc == LTR_OK) rc = scratch(ctx, pool, (size_t)n_items, &C.best_score);
  if (rc == LTR_OK) rc = scratch(ctx, pool, (size_t)n_items, &C.centroid_of);
  if (rc == LTR_OK) rc = scratch(ctx, pool, (size_t)n_sets, &C.cur_centroid);
  if (rc == LTR_OK) rc = scratch(ctx, pool, (size_t)n_sets, &C.next_centroid);
  if (rc == LTR_OK) rc = scratch(ctx, pool, (size_t)n_sets, &C.n_centroids);
  if (rc == LTR_OK) rc = scratch(ctx, pool, (size_t)n_sets, &C.state);
  if (rc == LTR_OK) rc = scratch(ctx, pool, (size_t)n_items, &C.pair_item);
  if (rc == LTR_OK) rc = scratch(ctx, pool, (size_t)n_items, &C.pair_a);
  if (rc == LTR_OK) rc = scratch(ctx, pool, (size_t)n_items, &C.pair_b);
  if (rc == LTR_OK) rc = scratch(ctx, pool, (size_t)n_items, &C.pair_T);
  if (rc == LTR_OK) rc = scratch(ctx, pool, (size_t)n_items, &A.out);
  if (rc == LTR_OK) rc = scratch(ctx, pool, 4, &d_words);
  unsigned long long* d_stat = nullptr;
  if (rc == LTR_OK) rc = scratch(ctx, pool, 2, &d_stat);
  A.line_stride = (max_len + 4 + 15) & ~15u;
  if (rc == LTR_OK && max_len > (uint32_t)kEditStripRows)
    rc = scratch(ctx, pool, (size_t)edit_myers_warps(ctx->sm_count) * A.line_stride, &A.lines);
  if (rc != LTR_OK) return rc;
  LTR_CUDA(ctx, cudaMemsetAsync(d_stat, 0, 2 * sizeof(unsigned long long), st));
  C.seq_off = d_off;
  C.stat = d_stat;
  C.n_sets = n_sets;
  C.n_items = n_items;
  C.set_begin = d_set_begin;
  C.set_T = d_set_T;
  C.item_set = d_item_set;
  C.item_seq = d_item_seq;
  C.score = A.out;
  C.n_pairs = d_words;
  C.cursor = d_words + 1;
  A.seq_bytes = d_bytes;
  A.seq_off = d_off;
  A.pair_a = C.pair_a;
  A.pair_b = C.pair_b;
  A.pair_T = C.pair_T;
  A.n_pairs_ptr = C.n_pairs;
  A.cursor = C.cursor;
  LTR_CUDA(ctx, cudaEventRecord(ctx->ev_start, st));
  LTR_CUDA(ctx, launch_cluster(C, A, ctx->sm_count, st));
  LTR_CUDA(ctx, cudaEventRecord(ctx->ev_end, st));
  std::vector<uint8_t> state(n_sets);
  LTR_CUDA(ctx, cudaMemcpyAsync(out_centroid_of, C.centroid_of, (size_t)n_items * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  LTR_CUDA(ctx, cudaMemcpyAsync(out_n_centroids, C.n_centroids, (size_t)n_sets * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  LTR_CUDA(ctx, cudaMemcpyAsync(state.data(), C.state, (size_t)n_sets, cudaMemcpyDeviceToHost, st));
  unsigned long long h_stat[2] = {0, 0};
  LTR_CUDA(ctx, cudaMemcpyAsync(h_stat, d_stat, sizeof(h_stat), cudaMemcpyDeviceToHost, st));
  LTR_CUDA(ctx, cudaStreamSynchronize(st));
  for (uint32_t k = 0; k < n_sets; ++k) out_ok[k] = state[k] == 1 ? 1 : 0;
  if (stats) {
    float ms = 0.f;
    LTR_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_end));
    stats->kernel_ms = ms;
    stats->n_pairs = h_stat[0];
    stats->n_cells = h_stat[1];
    stats->n_launches = 1 + 15 * 4;
    stats->h2d_bytes = h2d;
    stats->d2h_bytes = (uint64_t)n_items * sizeof(int32_t) + (uint64_t)n_sets * (sizeof(int32_t) + 1);
  }
  return LTR_OK;
}
