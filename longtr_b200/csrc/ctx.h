// ctx.h -- private definitions shared by the C-ABI translation units (abi.cu, stutter_abi.cu).
#pragma once
#include <cuda_runtime.h>

#include <string>

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

struct ltr_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t main_stream = nullptr;
  cudaStream_t streams[ltr::kNumStreams] = {nullptr};
  cudaEvent_t ev_start = nullptr, ev_vit = nullptr, ev_end = nullptr;
  cudaEvent_t ev_stream[ltr::kNumStreams] = {nullptr};
  cudaEvent_t ev_init = nullptr, ev_collect = nullptr;  // band phase ordering (no timing)
  int blocks_per_sm[2][32] = {{0}};
  int band_blocks_per_sm[16] = {0};  // by band class index
  int band_w = 0;                    // ltr_ctx_set_band: < 0 off, 0 automatic margin, > 0 margin in diagonals
  std::string last_error;
  void* stage[4] = {nullptr, nullptr, nullptr, nullptr};  // pinned host staging for the plan's large arrays (grow-only)
  size_t stage_bytes[4] = {0, 0, 0, 0};
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
