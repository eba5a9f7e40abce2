// ctx.h -- private definitions shared by the C-ABI translation units (abi.cu, stutter_abi.cu).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "longtr_b200.h"

namespace ltr {

const int kNumStreams = 8;

// Device allocations are stream ordered (cudaMallocAsync on the context's main stream, pool never trimmed):
// creating and destroying a job costs no device synchronisation, which matters for the per-locus entry points.
// The entry points name the stream with an AllocScope; without one (never in this library) cudaMalloc is used.
inline cudaStream_t& alloc_stream_slot() {
  static thread_local cudaStream_t s = nullptr;
  return s;
}
struct AllocScope {
  cudaStream_t prev;
  explicit AllocScope(cudaStream_t s) : prev(alloc_stream_slot()) { alloc_stream_slot() = s; }
  ~AllocScope() { alloc_stream_slot() = prev; }
};

struct DeviceBuffer {
  void* p = nullptr;
  size_t bytes = 0;
  cudaStream_t stream = nullptr;
  bool async = false;
  cudaError_t alloc(size_t n) {
    free();
    if (n == 0) n = 8;
    stream = alloc_stream_slot();
    async = (stream != nullptr);
    cudaError_t e = async ? cudaMallocAsync(&p, n, stream) : cudaMalloc(&p, n);
    if (e == cudaSuccess) bytes = n;
    else p = nullptr;
    return e;
  }
  void free() {
    if (p) {
      if (async) cudaFreeAsync(p, stream);
      else cudaFree(p);
    }
    p = nullptr;
    bytes = 0;
  }
  template <typename T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace ltr

namespace ltr {
// Streams of one job in flight.  Jobs of a context take the lanes in turn, so the upload of job k+1 (h2d) and the
// download of job k-1 (d2h) overlap the kernels of job k (main + one stream per row / band class).
const int kLanes = 3;
struct JobLane {
  cudaStream_t main = nullptr, h2d = nullptr, d2h = nullptr;
  cudaStream_t cls[kNumStreams] = {nullptr};
  cudaEvent_t ev_cls[kNumStreams] = {nullptr};
  cudaEvent_t ev_init = nullptr, ev_collect = nullptr;  // band phase ordering (no timing)
};
}  // namespace ltr

struct ltr_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t main_stream = nullptr;                    // one-shot entry points (ltr_posteriors, ltr_stutter_ll)
  cudaEvent_t ev_start = nullptr, ev_end = nullptr;      // ... and their timing
  ltr::JobLane lanes[ltr::kLanes];
  unsigned next_lane = 0;
  int blocks_per_sm[4][32] = {{0}};
  int band_blocks_per_sm[16] = {0};  // by band class index
  int read_encoding = 0;             // ltr_ctx_set_read_encoding: 0 bytes, 1 one 4-bit stream (BAM nibble codes)
  int band_w = 0;                    // ltr_ctx_set_band: < 0 off, 0 automatic margin, > 0 margin in diagonals
  int plan_mode = 0;                 // ltr_ctx_set_plan: 0 automatic, 1 host plan (make_plan), 2 device plan (plan_kernels.cu)
  std::string last_error;
  void* stage[4] = {nullptr, nullptr, nullptr, nullptr};  // pinned host staging for the host plan's large arrays (grow-only)
  size_t stage_bytes[4] = {0, 0, 0, 0};
  std::vector<void*> result_blocks;  // pinned host blocks for per-job statistics, recycled
  std::vector<cudaEvent_t> timing_events, plain_events;  // events of finished jobs, recycled (creation costs microseconds
                                                         // that the per-locus entry points would pay on every call)
};

namespace ltr {

inline int fail_cuda(ltr_ctx* ctx, cudaError_t e, const char* what) {
  if (ctx) ctx->last_error = std::string(what) + ": " + cudaGetErrorString(e);
  return (e == cudaErrorMemoryAllocation) ? LTR_ERR_OOM : LTR_ERR_CUDA;
}

}  // namespace ltr

#define LTR_CUDA(ctx, call)                                          \
  do {                                                               \
    cudaError_t e__ = (call);                                        \
    if (e__ != cudaSuccess) return ltr::fail_cuda(ctx, e__, #call);  \
  } while (0)
