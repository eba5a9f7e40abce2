// plan_kernels.cu -- the job plan on the device (plan_device.cuh): de-duplication of the trimmed reads of every locus,
// offsets, compaction and the task lists of the banded / full-matrix Viterbi kernels, as five small kernels that run in
// stream order between the upload of a batch and its DP kernels.  HBM-bound integer / byte work: one warp per locus
// (its reads are contiguous, so the lanes read neighbouring lines), one thread per haplotype for the task passes.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"
#include "plan_device.cuh"

namespace ltr {

static constexpr int kPlanBlock = 128;

__global__ void __launch_bounds__(kPlanBlock) plan_dedupe_kernel(const PlanDev P) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t l = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; l < P.n_loci; l += warps)
    plan_locus_dedupe(P, l, lane, 32u);
}

// Exclusive scans over the loci: distinct reads, their bytes, rows of the distinct LL matrices.  One CTA; every thread
// sums a contiguous chunk of loci, the chunk totals are scanned in shared memory, the chunk is then written out.
__global__ void __launch_bounds__(1024) plan_scan_kernel(const PlanDev P) {
  __shared__ unsigned long long s_u[1024], s_b[1024], s_l[1024];
  const uint32_t t = threadIdx.x;
  const bool err = P.ctl[PLAN_CTL_ERR] != 0;
  const uint32_t chunk = (P.n_loci + 1023u) / 1024u;
  const uint32_t l0 = min(P.n_loci, t * chunk), l1 = min(P.n_loci, l0 + chunk);
  unsigned long long su = 0, sb = 0, sl = 0;
  if (!err)
    for (uint32_t l = l0; l < l1; ++l) {
      const uint32_t c = P.ucount[l];
      su += c;
      sb += P.ubytes[l];
      sl += (unsigned long long)c * (P.lhb[l + 1] - P.lhb[l]);
    }
  s_u[t] = su;
  s_b[t] = sb;
  s_l[t] = sl;
  __syncthreads();
  for (uint32_t d = 1; d < 1024u; d <<= 1) {  // inclusive Hillis-Steele scan of the chunk totals
    unsigned long long a = 0, b = 0, c = 0;
    if (t >= d) {
      a = s_u[t - d];
      b = s_b[t - d];
      c = s_l[t - d];
    }
    __syncthreads();
    s_u[t] += a;
    s_b[t] += b;
    s_l[t] += c;
    __syncthreads();
  }
  unsigned long long nu = s_u[t] - su, nb = s_b[t] - sb, nll = s_l[t] - sl;  // exclusive prefix of the chunk
  for (uint32_t l = l0; l < l1; ++l) {
    P.lub[l] = (uint32_t)nu;
    P.ubyte_off[l] = (uint32_t)nb;
    P.ull_off[l] = nll;
    if (!err) {
      const uint32_t c = P.ucount[l];
      nu += c;
      nb += P.ubytes[l];
      nll += (unsigned long long)c * (P.lhb[l + 1] - P.lhb[l]);
    }
  }
  if (t == 1023u) {
    const unsigned long long tu = s_u[1023], tb = s_b[1023], tl = s_l[1023];
    P.lub[P.n_loci] = (uint32_t)tu;
    P.ubyte_off[P.n_loci] = (uint32_t)tb;
    P.ull_off[P.n_loci] = tl;
    P.ctl[PLAN_CTL_N_UREADS] = (uint32_t)tu;
    P.stat[PLAN_STAT_PAIRS_COMPUTED] = tl;
    if (tb > 0xFFFFFFF0ull) atomicOr(P.ctl + PLAN_CTL_ERR, 2u);
    if (P.n_loci == 0 || tu == 0) P.uread_off[0] = 0u;
  }
}

__global__ void __launch_bounds__(kPlanBlock) plan_fill_kernel(const PlanDev P) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t l = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; l < P.n_loci; l += warps)
    plan_locus_fill(P, l, lane, 32u);
}

__global__ void __launch_bounds__(kPlanBlock) plan_tasks_kernel(const PlanDev P, int pass) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t l = blockIdx.x * blockDim.x + threadIdx.x; l < P.n_loci; l += stride) plan_locus_tasks(P, l, pass);
}

// Exclusive scan of the [class][locus] band counters (tasks and pairs), class after class, into list positions; one CTA,
// same scheme as plan_scan_kernel.  Thread 0 also places the cost buckets of the stream lists.
__global__ void __launch_bounds__(1024) plan_task_scan_kernel(const PlanDev P) {
  __shared__ unsigned long long s_t[1024], s_p[1024];
  __shared__ uint32_t s_cls_t[kBandClasses + 1], s_cls_p[kBandClasses + 1];
  const uint32_t t = threadIdx.x;
  if (t == 0) plan_task_scan(P);
  const uint32_t N = (uint32_t)kBandClasses * P.n_loci;
  const uint32_t chunk = (N + 1023u) / 1024u;
  const uint32_t i0 = min(N, t * chunk), i1 = min(N, i0 + chunk);
  unsigned long long st = 0, sp = 0;
  for (uint32_t i = i0; i < i1; ++i) {
    st += P.band_task_pos[i];
    sp += P.band_pair_pos[i];
  }
  s_t[t] = st;
  s_p[t] = sp;
  __syncthreads();
  for (uint32_t d = 1; d < 1024u; d <<= 1) {
    unsigned long long a = 0, b = 0;
    if (t >= d) {
      a = s_t[t - d];
      b = s_p[t - d];
    }
    __syncthreads();
    s_t[t] += a;
    s_p[t] += b;
    __syncthreads();
  }
  unsigned long long nt = s_t[t] - st, np = s_p[t] - sp;
  for (uint32_t i = i0; i < i1; ++i) {
    if (i % P.n_loci == 0) {  // first entry of a band class: where its lists start
      s_cls_t[i / P.n_loci] = (uint32_t)nt;
      s_cls_p[i / P.n_loci] = (uint32_t)np;
    }
    const uint32_t a = P.band_task_pos[i], b = P.band_pair_pos[i];
    P.band_task_pos[i] = (uint32_t)nt;
    P.band_pair_pos[i] = (uint32_t)np;
    nt += a;
    np += b;
  }
  if (t == 1023u) {
    s_cls_t[kBandClasses] = (uint32_t)s_t[1023];
    s_cls_p[kBandClasses] = (uint32_t)s_p[1023];
  }
  __syncthreads();
  if (t < (uint32_t)kBandClasses) {
    P.ctl[PLAN_CTL_BAND_TASK_BASE + t] = s_cls_t[t];
    P.ctl[PLAN_CTL_BAND_TASK_COUNT + t] = s_cls_t[t + 1] - s_cls_t[t];
    P.ctl[PLAN_CTL_BAND_INFO + 2 * t] = s_cls_p[t];
    P.ctl[PLAN_CTL_BAND_INFO + 2 * t + 1] = s_cls_p[t + 1] - s_cls_p[t];
  }
  if (t == 0) {
    const uint32_t total_t = s_cls_t[kBandClasses];
    P.ctl[PLAN_CTL_N_BAND_TASKS] = total_t < P.band_cap ? total_t : P.band_cap;
    P.ctl[PLAN_CTL_N_BAND_PAIRS] = s_cls_p[kBandClasses];
  }
}

// The steps in stream order.  ctl / stat / the band counters must be zero on entry (the caller memsets them on the same
// stream).
cudaError_t launch_device_plan(const PlanDev& P, int sm_count, cudaStream_t stream) {
  if (P.n_loci == 0) return cudaSuccess;
  const uint32_t warps_per_block = kPlanBlock / 32;
  const uint32_t max_blocks = (uint32_t)sm_count * 16u;
  uint32_t locus_blocks = (P.n_loci + warps_per_block - 1) / warps_per_block;
  locus_blocks = locus_blocks < max_blocks ? locus_blocks : max_blocks;
  uint32_t task_blocks = (P.n_loci + kPlanBlock - 1) / kPlanBlock;
  task_blocks = task_blocks < max_blocks ? task_blocks : max_blocks;
  plan_dedupe_kernel<<<locus_blocks, kPlanBlock, 0, stream>>>(P);
  plan_scan_kernel<<<1, 1024, 0, stream>>>(P);
  plan_fill_kernel<<<locus_blocks, kPlanBlock, 0, stream>>>(P);
  plan_tasks_kernel<<<task_blocks, kPlanBlock, 0, stream>>>(P, 0);
  plan_task_scan_kernel<<<1, 1024, 0, stream>>>(P);
  plan_tasks_kernel<<<task_blocks, kPlanBlock, 0, stream>>>(P, 1);
  return cudaGetLastError();
}

}  // namespace ltr
