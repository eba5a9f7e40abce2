// microbench.cu -- FP64-pipe issue-rate probe used as the roofline denominator.
//
// the hot path is bound by FP64 instruction issue (DADD and
// DSETP share the FP64 pipe), a figure MEASURED_PEAKS.json does not carry, so it is measured
// on the box: kind 0 = independent DADD streams, kind 1 = DSETP streams, kind 2 = the
// DADD+DADD+DSETP+2xFSEL pattern of one max-plus term.  Reported in lane-operations/s.
#include <cuda_runtime.h>
#include <stdint.h>

#include "longtr_b200.h"

namespace {

template <int KIND>
__global__ void __launch_bounds__(256) fp64_probe(double* out, int iters, double c1, double c2) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = -1.0 - (double)(threadIdx.x + i);
  unsigned cnt = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (KIND == 0) {
        a[i] = a[i] + c1;
      } else if (KIND == 1) {
        cnt += (a[i] < c1 + (double)it) ? 1u : 0u;
      } else {
        const double x = a[i] + c1, y = a[(i + 1) & 7] + c2;
        a[i] = (x < y) ? y : x;
      }
    }
  }
  double s = (double)cnt;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 12345.678) out[0] = s;  // keep the work alive
}

}  // namespace

extern "C" int ltr_fp64_issue_rate(int device, int kind, double* lane_ops_per_s, double* ms_out) {
  if (!lane_ops_per_s || kind < 0 || kind > 2) return LTR_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return LTR_ERR_NO_DEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return LTR_ERR_NO_DEVICE;
  double* d = nullptr;
  if (cudaMalloc(&d, 64) != cudaSuccess) return LTR_ERR_OOM;
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    if (kind == 0) fp64_probe<0><<<blocks, threads>>>(d, iters, -1e-7, -2e-7);
    if (kind == 1) fp64_probe<1><<<blocks, threads>>>(d, iters, -1e-7, -2e-7);
    if (kind == 2) fp64_probe<2><<<blocks, threads>>>(d, iters, -1e-7, -2e-7);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return LTR_ERR_CUDA; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  // FP64-pipe operations per inner iteration: kind 0: 1 DADD, kind 1: 1 DSETP (+1 DADD for c1+it,
  // hoisted), kind 2: 2 DADD + 1 DSETP
  const double per = (kind == 2) ? 3.0 : 1.0;
  const double ops = (double)blocks * threads * (double)iters * 8.0 * per;
  *lane_ops_per_s = ops / ((double)best * 1e-3);
  if (ms_out) *ms_out = best;
  return LTR_OK;
}
