// plan_kernels.cu -- the job plan on the device (plan_device.cuh): de-duplication of the trimmed reads of every locus,
// offsets, compaction and the task lists of the banded / full-matrix Viterbi kernels, as a handful of small kernels that run in
// stream order between the upload of a batch and its DP kernels.  HBM-bound integer / byte work: one warp per locus
// (its reads are contiguous, so the lanes read neighbouring lines), one thread per haplotype for the task passes.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"
#include "plan_device.cuh"

namespace ltr {

static constexpr int kPlanBlock = 128;

// Loci of at most 32 pooled reads (the common case): one read per lane, everything the lanes need from each other travels
// through shuffles instead of the global scratch arrays of plan_locus_dedupe -- same numbering, same outputs.
__device__ __forceinline__ void plan_locus_dedupe_warp(const PlanDev& P, uint32_t l, uint32_t lane, const uint8_t* bytes,
                                                       uint32_t origin, uint32_t limit) {
  const uint32_t r0 = P.lrb[l], R = P.lrb[l + 1] - r0;
  for (uint32_t h = P.lhb[l] + lane; h < P.lhb[l + 1]; h += 32u) P.hap_locus[h] = l;
  const bool have = lane < R;
  const uint32_t r = r0 + lane;
  uint32_t o0 = origin, len = 0;
  bool ok = true;
  if (have) {
    o0 = P.read_off[r];
    const uint32_t o1 = P.read_off[r + 1];
    ok = (o1 > o0) && (o1 <= limit) && (o0 >= origin);
    len = ok ? o1 - o0 : 0u;
    if (!ok) o0 = origin;
    P.read_locus[r] = l;
  }
  const uint8_t* s = bytes + (o0 - origin);
  const unsigned long long h = have ? plan_hash(s, len) : 0ull;
  if (__any_sync(0xFFFFFFFFu, !ok) && lane == 0) atomicOr(P.ctl + PLAN_CTL_ERR, 1u);
  const uint32_t max_m = __reduce_max_sync(0xFFFFFFFFu, len);
  if (lane == 0 && max_m) atomicMax(P.stat + PLAN_STAT_MAX_M, (unsigned long long)max_m);
  uint32_t rep = lane;  // first read of the locus with the same sequence
  for (uint32_t q = 0; q + 1 < R; ++q) {
    const unsigned long long hq = __shfl_sync(0xFFFFFFFFu, h, q);
    const uint32_t lq = __shfl_sync(0xFFFFFFFFu, len, q), oq = __shfl_sync(0xFFFFFFFFu, o0, q);
    if (have && q < lane && rep == lane && hq == h && lq == len)
      if (len == 0 || plan_equal(bytes + (oq - origin), s, len)) rep = q;
  }
  const bool is_rep = have && rep == lane;
  uint32_t rank = 0;  // by (length, first occurrence) among the representatives
  for (uint32_t q = 0; q < R; ++q) {
    const bool iq = __shfl_sync(0xFFFFFFFFu, is_rep ? 1u : 0u, q) != 0u;
    const uint32_t lq = __shfl_sync(0xFFFFFFFFu, len, q);
    rank += (iq && (lq < len || (lq == len && q < lane))) ? 1u : 0u;
  }
  const uint32_t k = __shfl_sync(0xFFFFFFFFu, rank, rep);
  if (have) {
    P.local_u[r] = k | (is_rep ? kPlanRep : 0u);
    if (is_rep) {
      P.tmp_len[r0 + k] = len;
      P.tmp_rep[r0 + k] = r;
    }
  }
  const uint32_t n_rep = __popc(__ballot_sync(0xFFFFFFFFu, is_rep));
  const uint32_t n_bytes = __reduce_add_sync(0xFFFFFFFFu, is_rep ? len : 0u);
  if (lane == 0) {
    P.ucount[l] = n_rep;
    P.ubytes[l] = n_bytes;
  }
}

// One warp per locus.  The raw reads of a locus are contiguous: the warp copies them into shared memory with coalesced
// 16-byte loads (kStageBytes per warp; loci that do not fit, e.g. 30 reads of a 1 kb VNTR, are read in place) and hashes /
// compares them there.
static constexpr uint32_t kStageBytes = 16384;
__global__ void __launch_bounds__(kPlanBlock) plan_dedupe_kernel(const PlanDev P) {
  extern __shared__ __align__(16) uint8_t plan_smem[];
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  uint8_t* stage = plan_smem + (size_t)(threadIdx.x >> 5) * kStageBytes;
  for (uint32_t l = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; l < P.n_loci; l += warps) {
    const uint32_t o0 = P.read_off[P.lrb[l]], o1 = P.read_off[P.lrb[l + 1]];
    const uint32_t origin = o0 & ~15u;
    const bool fits = (o1 >= o0) && (o1 <= P.raw_total) && (o1 - origin + 32u <= kStageBytes);
    if (fits) {
      __syncwarp();  // the previous locus' readers are done with the buffer
      const uint4* src = reinterpret_cast<const uint4*>(P.read_bytes + origin);  // buffer base + 512-byte pad: 16-aligned
      uint4* dst = reinterpret_cast<uint4*>(stage);
      const uint32_t n16 = (o1 - origin + 16u + 15u) >> 4;  // one word past the end for plan_load32
      for (uint32_t i = lane; i < n16; i += 32u) dst[i] = src[i];
      __syncwarp();
      if (P.lrb[l + 1] - P.lrb[l] <= 32u) plan_locus_dedupe_warp(P, l, lane, stage, origin, o1);
      else plan_locus_dedupe(P, l, lane, 32u, stage, origin, o1);
    } else {
      if (P.lrb[l + 1] - P.lrb[l] <= 32u) plan_locus_dedupe_warp(P, l, lane, P.read_bytes, 0u, P.raw_total);
      else plan_locus_dedupe(P, l, lane, 32u, P.read_bytes, 0u, P.raw_total);
    }
  }
}

// ---- multi-block exclusive scans -----------------------------------------------------------------------------------------
// Three small kernels per scan: (1) every CTA reduces a tile of kScanTile elements, (2) one CTA scans the tile sums,
// (3) every CTA scans its tile in shared memory on top of its tile offset and writes the results.  IO is a policy:
//   COLS columns scanned together; load(i, v[COLS]); store(i, prefix[COLS]); total(prefix[COLS]) once, at the end.
static constexpr int kScanThreads = 256, kScanPer = 4, kScanTile = kScanThreads * kScanPer;

struct LocusScanIO {  // distinct reads, their bytes, rows of the distinct LL matrices: per locus -> prefix sums
  static constexpr int COLS = 3;
  PlanDev P;
  __device__ uint32_t size() const { return P.n_loci; }
  __device__ void load(uint32_t l, unsigned long long* v) const {
    const bool err = P.ctl[PLAN_CTL_ERR] != 0;
    const uint32_t c = err ? 0u : P.ucount[l];
    v[0] = c;
    v[1] = err ? 0u : P.ubytes[l];
    v[2] = (unsigned long long)c * (P.lhb[l + 1] - P.lhb[l]);
  }
  __device__ void store(uint32_t l, const unsigned long long* v) const {
    P.lub[l] = (uint32_t)v[0];
    P.ubyte_off[l] = (uint32_t)v[1];
    P.ull_off[l] = v[2];
  }
  __device__ void total(const unsigned long long* v) const {
    store(P.n_loci, v);
    P.ctl[PLAN_CTL_N_UREADS] = (uint32_t)v[0];
    P.stat[PLAN_STAT_PAIRS_COMPUTED] = v[2];
    if (v[1] > 0xFFFFFFF0ull) atomicOr(P.ctl + PLAN_CTL_ERR, 2u);
    if (v[0] == 0) P.uread_off[0] = 0u;
  }
};

struct BandScanIO {  // [class][locus] band counters (tasks, pairs) -> list positions, class after class, in place
  static constexpr int COLS = 2;
  PlanDev P;
  __device__ uint32_t size() const { return (uint32_t)kBandClasses * P.n_loci; }
  __device__ void load(uint32_t i, unsigned long long* v) const {
    v[0] = P.band_task_pos[i];
    v[1] = P.band_pair_pos[i];
  }
  __device__ void store(uint32_t i, const unsigned long long* v) const {
    P.band_task_pos[i] = (uint32_t)v[0];
    P.band_pair_pos[i] = (uint32_t)v[1];
    if (i % P.n_loci == 0) {  // first entry of a band class: where its lists start
      const uint32_t c = i / P.n_loci;
      P.ctl[PLAN_CTL_BAND_TASK_BASE + c] = (uint32_t)v[0];
      P.ctl[PLAN_CTL_BAND_INFO + 2 * c] = (uint32_t)v[1];
    }
  }
  __device__ void total(const unsigned long long* v) const {
    P.ctl[PLAN_CTL_N_BAND_TASKS] = v[0] < P.band_cap ? (uint32_t)v[0] : P.band_cap;
    P.ctl[PLAN_CTL_N_BAND_PAIRS] = (uint32_t)v[1];
  }
};

template <typename IO>
__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const IO io, unsigned long long* __restrict__ partial) {
  __shared__ unsigned long long s_w[IO::COLS][kScanThreads / 32];
  const uint32_t n = io.size(), base = blockIdx.x * (uint32_t)kScanTile;
  unsigned long long acc[IO::COLS];
#pragma unroll
  for (int c = 0; c < IO::COLS; ++c) acc[c] = 0;
  for (int j = 0; j < kScanPer; ++j) {
    const uint32_t i = base + (uint32_t)j * kScanThreads + threadIdx.x;
    if (i < n) {
      unsigned long long v[IO::COLS];
      io.load(i, v);
#pragma unroll
      for (int c = 0; c < IO::COLS; ++c) acc[c] += v[c];
    }
  }
#pragma unroll
  for (int c = 0; c < IO::COLS; ++c) {
    for (int d = 16; d > 0; d >>= 1) acc[c] += __shfl_down_sync(0xFFFFFFFFu, acc[c], d);
    if ((threadIdx.x & 31) == 0) s_w[c][threadIdx.x >> 5] = acc[c];
  }
  __syncthreads();
  if (threadIdx.x < IO::COLS) {
    unsigned long long t = 0;
    for (int w = 0; w < kScanThreads / 32; ++w) t += s_w[threadIdx.x][w];
    partial[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = t;
  }
}

// One CTA: exclusive scan of the n_tiles tile sums of every column in place; the grand totals go to io.total().
template <typename IO>
__global__ void __launch_bounds__(1024) scan_partials_kernel(const IO io, unsigned long long* __restrict__ partial, uint32_t n_tiles) {
  __shared__ unsigned long long s_v[IO::COLS][1024];
  const uint32_t t = threadIdx.x;
  const uint32_t chunk = (n_tiles + 1023u) / 1024u;
  const uint32_t i0 = min(n_tiles, t * chunk), i1 = min(n_tiles, i0 + chunk);
  unsigned long long own[IO::COLS];
#pragma unroll
  for (int c = 0; c < IO::COLS; ++c) {
    unsigned long long a = 0;
    for (uint32_t i = i0; i < i1; ++i) a += partial[(size_t)c * n_tiles + i];
    own[c] = a;
    s_v[c][t] = a;
  }
  __syncthreads();
  for (uint32_t d = 1; d < 1024u; d <<= 1) {  // inclusive Hillis-Steele scan of the chunk totals
    unsigned long long a[IO::COLS];
#pragma unroll
    for (int c = 0; c < IO::COLS; ++c) a[c] = (t >= d) ? s_v[c][t - d] : 0ull;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < IO::COLS; ++c) s_v[c][t] += a[c];
    __syncthreads();
  }
#pragma unroll
  for (int c = 0; c < IO::COLS; ++c) {
    unsigned long long run = s_v[c][t] - own[c];
    for (uint32_t i = i0; i < i1; ++i) {
      const unsigned long long v = partial[(size_t)c * n_tiles + i];
      partial[(size_t)c * n_tiles + i] = run;
      run += v;
    }
  }
  if (t == 1023u) {
    unsigned long long tot[IO::COLS];
#pragma unroll
    for (int c = 0; c < IO::COLS; ++c) tot[c] = s_v[c][1023];
    io.total(tot);
  }
}

template <typename IO>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const IO io, const unsigned long long* __restrict__ partial) {
  __shared__ unsigned long long s_e[IO::COLS][kScanTile + kScanTile / 32];  // padded: thread t scans elements kScanPer * t ...
  __shared__ unsigned long long s_w[IO::COLS][kScanThreads / 32];
  const uint32_t n = io.size(), base = blockIdx.x * (uint32_t)kScanTile;
  auto slot = [](uint32_t e) { return e + (e >> 5); };
  for (int j = 0; j < kScanPer; ++j) {
    const uint32_t e = (uint32_t)j * kScanThreads + threadIdx.x, i = base + e;
    unsigned long long v[IO::COLS];
#pragma unroll
    for (int c = 0; c < IO::COLS; ++c) v[c] = 0;
    if (i < n) io.load(i, v);
#pragma unroll
    for (int c = 0; c < IO::COLS; ++c) s_e[c][slot(e)] = v[c];
  }
  __syncthreads();
  unsigned long long sum[IO::COLS];
#pragma unroll
  for (int c = 0; c < IO::COLS; ++c) {
    unsigned long long a = 0;
    for (int j = 0; j < kScanPer; ++j) a += s_e[c][slot(threadIdx.x * kScanPer + j)];
    sum[c] = a;
    unsigned long long inc = a;  // inclusive scan of the thread sums inside the warp
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long up = __shfl_up_sync(0xFFFFFFFFu, inc, d);
      if ((threadIdx.x & 31) >= d) inc += up;
    }
    if ((threadIdx.x & 31) == 31) s_w[c][threadIdx.x >> 5] = inc;
    sum[c] = inc - a;  // exclusive inside the warp
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < IO::COLS; ++c) {
    unsigned long long off = partial[(size_t)c * gridDim.x + blockIdx.x];
    for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) off += s_w[c][w];
    unsigned long long run = off + sum[c];
    for (int j = 0; j < kScanPer; ++j) {
      const uint32_t sl = slot(threadIdx.x * kScanPer + j);
      const unsigned long long v = s_e[c][sl];
      s_e[c][sl] = run;
      run += v;
    }
  }
  __syncthreads();
  for (int j = 0; j < kScanPer; ++j) {
    const uint32_t e = (uint32_t)j * kScanThreads + threadIdx.x, i = base + e;
    if (i < n) {
      unsigned long long v[IO::COLS];
#pragma unroll
      for (int c = 0; c < IO::COLS; ++c) v[c] = s_e[c][slot(e)];
      io.store(i, v);
    }
  }
}

template <typename IO>
static void launch_scan(const IO& io, uint32_t n, unsigned long long* partial, cudaStream_t stream) {
  const uint32_t n_tiles = (n + kScanTile - 1) / kScanTile;
  scan_reduce_kernel<IO><<<n_tiles, kScanThreads, 0, stream>>>(io, partial);
  scan_partials_kernel<IO><<<1, 1024, 0, stream>>>(io, partial, n_tiles);
  scan_apply_kernel<IO><<<n_tiles, kScanThreads, 0, stream>>>(io, partial);
}

// Positions of the cost buckets of the stream lists + band class sizes from the scanned class bases (one small CTA).
__global__ void plan_task_finish_kernel(const PlanDev P) {
  if (threadIdx.x == 0) plan_task_scan(P);
  if (threadIdx.x < (uint32_t)kBandClasses) {
    const uint32_t c = threadIdx.x;
    const uint32_t t0 = P.ctl[PLAN_CTL_BAND_TASK_BASE + c], p0 = P.ctl[PLAN_CTL_BAND_INFO + 2 * c];
    const uint32_t t1 = (c + 1 < (uint32_t)kBandClasses) ? P.ctl[PLAN_CTL_BAND_TASK_BASE + c + 1] : P.ctl[PLAN_CTL_N_BAND_TASKS];
    const uint32_t p1 = (c + 1 < (uint32_t)kBandClasses) ? P.ctl[PLAN_CTL_BAND_INFO + 2 * c + 2] : P.ctl[PLAN_CTL_N_BAND_PAIRS];
    P.ctl[PLAN_CTL_BAND_TASK_COUNT + c] = t1 - t0;
    P.ctl[PLAN_CTL_BAND_INFO + 2 * c + 1] = p1 - p0;
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

// Reads handed over as a 4-bit stream (BAM's sequence encoding, ltr_ctx_set_read_encoding): every thread expands 8 packed
// bytes into 16 bases.  HBM-bound: n / 2 bytes in, n bytes out.
__global__ void __launch_bounds__(256) plan_unpack_kernel(const uint8_t* __restrict__ packed, uint8_t* __restrict__ out,
                                                          uint32_t n_bases) {
  const uint32_t n_groups = (n_bases + 15u) >> 4;
  for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += gridDim.x * blockDim.x) {
    const uint32_t b0 = g << 4;
    const uint32_t n_in = ((n_bases + 1u) >> 1) - (b0 >> 1);  // packed bytes left from this group on
    uint32_t w[2] = {0u, 0u};
    if (n_in >= 8u) {
      const uint2 v = *reinterpret_cast<const uint2*>(packed + (b0 >> 1));
      w[0] = v.x;
      w[1] = v.y;
    } else {
      for (uint32_t k = 0; k < n_in; ++k) w[k >> 2] |= (uint32_t)packed[(b0 >> 1) + k] << (8u * (k & 3u));
    }
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t word = 0u;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t b = (uint32_t)(4 * q + k);               // base within the group
        const uint32_t byte = (w[b >> 3] >> (8u * ((b >> 1) & 3u))) & 0xFFu;
        const uint32_t nib = (b & 1u) ? (byte & 15u) : (byte >> 4);
        const uint32_t ch = (b0 + b < n_bases) ? (uint32_t)(uint8_t)"=ACMGRSVTWYHKDBN"[nib] : 0u;
        word |= ch << (8 * k);
      }
      o[q] = word;
    }
    if (b0 + 16u <= n_bases) {
      *reinterpret_cast<uint4*>(out + b0) = make_uint4(o[0], o[1], o[2], o[3]);
    } else {
      for (uint32_t b = 0; b0 + b < n_bases; ++b) out[b0 + b] = (uint8_t)(o[b >> 2] >> (8u * (b & 3u)));
    }
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
  if (P.packed && P.raw_total) {
    uint32_t blocks = ((P.raw_total + 15u) / 16u + 255u) / 256u;
    blocks = blocks < max_blocks ? blocks : max_blocks;
    plan_unpack_kernel<<<blocks, 256, 0, stream>>>(P.packed, const_cast<uint8_t*>(P.read_bytes), P.raw_total);
  }
  const size_t dedupe_smem = (size_t)warps_per_block * kStageBytes;
  cudaFuncSetAttribute(plan_dedupe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dedupe_smem);  // per device
  plan_dedupe_kernel<<<locus_blocks, kPlanBlock, dedupe_smem, stream>>>(P);
  LocusScanIO lio;
  lio.P = P;
  launch_scan(lio, P.n_loci, P.scan_partial, stream);
  plan_fill_kernel<<<locus_blocks, kPlanBlock, 0, stream>>>(P);
  plan_tasks_kernel<<<task_blocks, kPlanBlock, 0, stream>>>(P, 0);
  BandScanIO bio;
  bio.P = P;
  launch_scan(bio, (uint32_t)kBandClasses * P.n_loci, P.scan_partial, stream);
  plan_task_finish_kernel<<<1, 32, 0, stream>>>(P);
  plan_tasks_kernel<<<task_blocks, kPlanBlock, 0, stream>>>(P, 1);
  return cudaGetLastError();
}

}  // namespace ltr
