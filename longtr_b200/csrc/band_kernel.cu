// band_kernel.cu -- kernel 1b: banded anti-diagonal Viterbi with an exactness certificate (band_core.cuh).
//
// viterbi_band_kernel<K, G, SYM>: persistent warps; a warp takes a round of 32 / G consecutive pairs of its band class
// (one per group of G lanes; consecutive pairs share the haplotype and their reads are sorted by length, so the
// groups run almost the same number of anti-diagonals) and walks them in lock step: the branch-free double step from
// the first anti-diagonal on -- followed, during the prologue, by a fix-up that gives the boundary cells of the step
// their closed forms -- and a general step around the end cells.  FP64 max-plus on the FP64 pipe like
// viterbi_stream_kernel: 9 DADD + 4 DSETP per cell (7 + 4 for symmetric parameters).
// band_expand_kernel turns the plan's (haplotype, read range) tasks into the pair list; band_collect_kernel turns
// the pairs the band could not certify into tasks of viterbi_stream_kernel (runs of consecutive reads stay one task).
#include <cuda_runtime.h>
#include <stdint.h>

#include "band_core.cuh"
#include "kernels.h"

namespace ltr {

static constexpr unsigned kFull = 0xFFFFFFFFu;
static constexpr int kBandBlockThreads = 128;

struct SmemTable {  // the lane's closed-form boundary cells, [2K][3][32] doubles per warp
  const double* base;
  __device__ __forceinline__ double x(int q) const { return base[(3 * q) * 32]; }
  __device__ __forceinline__ double y(int q) const { return base[(3 * q + 1) * 32]; }
  __device__ __forceinline__ double z(int q) const { return base[(3 * q + 2) * 32]; }
};

#ifndef LTR_BAND_MINBLOCKS
#define LTR_BAND_MINBLOCKS 4
#endif
// Unroll factor of the steady-state loop (one iteration = one double step).  Unrolling removes the register moves of
// the sliding character windows on the loop's back edge (2.75 of 27 issue slots per cell) at the price of code size.
#ifndef LTR_BAND_UNROLL
#define LTR_BAND_UNROLL 2
#endif
static constexpr int kBandUnroll = LTR_BAND_UNROLL;
#ifndef LTR_BAND_MINBLOCKS_G4
#define LTR_BAND_MINBLOCKS_G4 4
#endif
template <int K, int G, bool SYM>
__global__ void __launch_bounds__(kBandBlockThreads, (G == 4 ? LTR_BAND_MINBLOCKS_G4 : (K <= 4 ? LTR_BAND_MINBLOCKS : (K <= 6 ? 3 : 2))))
viterbi_band_kernel(const VitConsts C, const DevBatch B, const BandArgs A) {
  constexpr int PPR = 32 / G;  // pairs per round
  extern __shared__ __align__(16) double band_smem[];
  const int lane = threadIdx.x & 31;
  const int lg = lane & (G - 1);
  const int grp = lane / G;
  double* tb = band_smem + (size_t)(threadIdx.x >> 5) * (6 * K * 32) + lane;
  SmemTable T;
  T.base = tb;
  constexpr int W = 2 * K * G;
  const uint2* __restrict__ pairs = A.pairs + A.info[0];
  const uint32_t n_pairs = A.info[1];
  const uint32_t n_rounds = (n_pairs + (uint32_t)PPR - 1u) / (uint32_t)PPR;
  while (true) {
    uint32_t round = 0, att = 0, fl = 0;
    if (lane == 0) {
      round = atomicAdd(A.cursor, 1u);
      att = *(volatile uint32_t*)(A.counters + 0);
      fl = *(volatile uint32_t*)(A.counters + 1);
    }
    round = __shfl_sync(kFull, round, 0);
    if (round >= n_rounds) break;
    att = __shfl_sync(kFull, att, 0);
    fl = __shfl_sync(kFull, fl, 0);
    const uint32_t base = round * (uint32_t)PPR;
    const bool active = (base + (uint32_t)grp) < n_pairs;
    const uint2 pr = pairs[active ? base + (uint32_t)grp : base];  // idle groups shadow the round's first pair
    const uint32_t g = pr.x, u = pr.y;
    const uint32_t hoff = B.hap_off[g];
    const int32_t hlen = (int32_t)(B.hap_off[g + 1] - hoff);
    const uint32_t qb = B.read_off[u];
    BandPair R;
    R.n = hlen - 2 * C.cut;
    R.m = (int32_t)(B.read_off[u + 1] - qb);
    R.hap = B.hap_bytes + hoff + C.cut;
    R.read = B.read_bytes + qb;
    const BandGeom geo = band_geometry(R.n, R.m, W);
    R.d0 = geo.dlo + 2 * K * lg;
    const uint32_t l = B.hap_locus[g];
    const uint32_t hb0 = B.locus_hap_begin[l];
    const uint32_t H = B.locus_hap_begin[l + 1] - hb0;
    double* out = B.out_ll + B.ll_off[l] + (unsigned long long)(u - B.locus_read_begin[l]) * H + (g - hb0);
    // Most pairs so far could not be certified (noisy reads, band too narrow): stop trying, hand the rest to the
    // full-matrix kernel.  Performance heuristic only -- both routes give the reference's bits.
    if (A.abandon_after && att >= A.abandon_after && 2u * fl > att) {
      if (active && lg == 0) *out = band_mark(kBandAbandoned);
      continue;
    }
    // ---- per-lane closed forms of the boundary cells of its diagonals, initial state -----------------------------
#pragma unroll
    for (int q = 0; q < 2 * K; ++q) {
      double bx, by, bz;
      band_boundary(C, R, R.d0 + q, bx, by, bz);
      tb[(3 * q) * 32] = bx;
      tb[(3 * q + 1) * 32] = by;
      tb[(3 * q + 2) * 32] = bz;
    }
    BandLane<K> L;
    band_lane_reset<K>(L, C);
    const int32_t s_end = R.n + R.m - 2;
    const int32_t s_pro = (int32_t)__reduce_max_sync(kFull, (uint32_t)band_prologue_steps(geo.dlo, W));
    const int32_t s_end_min = (int32_t)__reduce_min_sync(kFull, (uint32_t)s_end);
    const int32_t s_end_max = (int32_t)__reduce_max_sync(kFull, (uint32_t)s_end);
    double F = C.imp;
    bool got = false;
    int32_t s = 0;
#define LTR_BAND_GENERAL_STEP()                                    \
  do {                                                             \
    if (s & 1) {                                                   \
      double yr = __shfl_down_sync(kFull, L.A[0], 1);              \
      if (lg == G - 1) yr = C.imp;                                 \
      band_general_step<K, 1, SYM>(L, C, R, T, s, yr, F, got);     \
    } else {                                                       \
      double zl = __shfl_up_sync(kFull, L.B[K - 1], 1);            \
      if (lg == 0) zl = C.imp;                                     \
      band_general_step<K, 0, SYM>(L, C, R, T, s, zl, F, got);     \
    }                                                              \
    ++s;                                                           \
  } while (0)
    // Prologue and steady state share one loop body: plain double steps on the register windows; during the prologue
    // (s < s_pro) the cells that are boundary cells at the step are then replaced by their closed forms.  The windows of
    // lanes whose cells still lie before the matrix read up to W/2 bytes in front of the strings (padded buffers).
    if (s + 1 < s_end_min) {
      band_windows_init<K>(L, R, s);
      const uint8_t* hp = R.hap + ((s - R.d0) >> 1) + 1;
      const uint8_t* rp = R.read + ((s + R.d0) >> 1) + K + 1;
#pragma unroll 1
      for (; s < s_pro && s + 1 < s_end_min; s += 2) {
        const int32_t nh = (int32_t)*hp, nr = (int32_t)*rp;
        ++hp;
        ++rp;
        double zl = __shfl_up_sync(kFull, L.B[K - 1], 1);
        if (lg == 0) zl = C.imp;
        band_fast_even<K, SYM>(L, C, zl);
        band_fixup<K, 0>(L, R, T, s);
        double yr = __shfl_down_sync(kFull, L.A[0], 1);
        if (lg == G - 1) yr = C.imp;
        band_fast_odd<K, SYM>(L, C, yr, nh, nr);
        band_fixup<K, 1>(L, R, T, s + 1);
      }
#pragma unroll kBandUnroll
      for (; s + 1 < s_end_min; s += 2) {
        const int32_t nh = (int32_t)*hp, nr = (int32_t)*rp;  // consumed after both steps
        ++hp;
        ++rp;
        double zl = __shfl_up_sync(kFull, L.B[K - 1], 1);
        if (lg == 0) zl = C.imp;
        band_fast_even<K, SYM>(L, C, zl);
        double yr = __shfl_down_sync(kFull, L.A[0], 1);
        if (lg == G - 1) yr = C.imp;
        band_fast_odd<K, SYM>(L, C, yr, nh, nr);
      }
    }
    while (s <= s_end_max) LTR_BAND_GENERAL_STEP();
#undef LTR_BAND_GENERAL_STEP
    // ---- result: the lane that owns diagonal de holds the score ----------------------------------------------------
    const double thr = band_threshold(C, A.gap, R.n, R.m, geo.w);
    const bool mine = got && active;
    const bool ok = mine && (F > thr);
    if (mine) *out = ok ? F : band_mark(F);
    const unsigned done = __ballot_sync(kFull, mine), good = __ballot_sync(kFull, ok);
    const uint32_t cells = __reduce_add_sync(kFull, mine ? (uint32_t)band_cells(R.n, R.m, W, geo.dlo) : 0u);
    if (lane == 0) {
      atomicAdd(A.cells_evaluated, (unsigned long long)cells);
      atomicAdd(A.counters + 0, (uint32_t)__popc(done));
      if (done != good) atomicAdd(A.counters + 1, (uint32_t)__popc(done & ~good));
    }
  }
}

// pairs[cum[t] + i] = (task t's haplotype, read_begin + i)
__global__ void band_expand_kernel(const BandTask* __restrict__ tasks, const uint32_t* __restrict__ cum,
                                   uint32_t n_tasks, uint2* __restrict__ pairs) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tasks) return;
  const BandTask bt = tasks[t];
  uint2* dst = pairs + cum[t];
  for (uint32_t r = bt.read_begin; r < bt.read_end; ++r) dst[r - bt.read_begin] = make_uint2(bt.hap, r);
}

// One thread per band task.  An uncertified pair either gets a second chance in a wider band class that is certain to
// certify it (band_retry_class: S.retry_pairs, one list per class) or joins a run of consecutive reads that becomes ONE
// task of the stream kernel of the haplotype's row class.  Two passes so that the appended tasks end up heaviest first
// (the persistent warps of the stream kernel then finish together, like on the plan's own sorted list): pass 0 counts the
// runs per (row class, cost bucket = floor(log2 cost)) and the retries per band class, band_bucket_scan_kernel turns the
// counts into list positions, pass 1 writes the tasks and the retry pairs.
__global__ void band_collect_kernel(const VitConsts C, const DevBatch B, const BandTask* __restrict__ tasks,
                                    const uint32_t* __restrict__ n_tasks_ptr, uint32_t task_cap, const BandCollect S,
                                    int pass) {
  uint32_t n_tasks = *n_tasks_ptr;
  n_tasks = n_tasks < task_cap ? n_tasks : task_cap;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_tasks; t += gridDim.x * blockDim.x) {
  const BandTask bt = tasks[t];
  const uint32_t g = bt.hap;
  const uint32_t l = B.hap_locus[g];
  const uint32_t hb0 = B.locus_hap_begin[l];
  const uint32_t H = B.locus_hap_begin[l + 1] - hb0;
  const uint32_t rb0 = B.locus_read_begin[l];
  const int32_t n = (int32_t)(B.hap_off[g + 1] - B.hap_off[g]) - 2 * C.cut;
  const int kr = rows_per_lane_hd(n, S.kmax);
  int strips = (n - 1 + 32 * kr - 1) / (32 * kr);
  if (strips < 1) strips = 1;
  const double* col = B.out_ll + B.ll_off[l] + (g - hb0);
  uint32_t run_begin = 0, n_bad = 0, n_retry = 0;
  unsigned long long cells = 0;
  bool in_run = false;
  for (uint32_t r = bt.read_begin; r <= bt.read_end; ++r) {
    bool bad = false;
    if (r < bt.read_end) {
      const double v = col[(unsigned long long)(r - rb0) * H];
      bad = band_marked(v);
      if (bad && S.retry_pairs) {
        const int rc = band_retry_class(C, S.gap, n, (int32_t)(B.read_off[r + 1] - B.read_off[r]), band_unmark(v), S.retry_rho_pct);
        if (rc >= 0) {
          bad = false;
          ++n_retry;
          if (pass == 0) {
            atomicAdd(S.retry_count + rc, 1u);
          } else {
            const uint32_t k = S.retry_info[2 * rc] + atomicAdd(S.retry_fill + rc, 1u);
            if (k < S.retry_cap) S.retry_pairs[k] = make_uint2(g, r);
          }
        }
      }
    }
    if (bad) {
      ++n_bad;
      cells += (unsigned long long)n * (unsigned long long)(B.read_off[r + 1] - B.read_off[r]);
      if (!in_run) {
        in_run = true;
        run_begin = r;
      }
    } else if (in_run) {
      in_run = false;
      // same cost model as the plan (viterbi_host.h): rows per lane x strips x stream length
      const unsigned long long cost =
          (unsigned long long)kr * (unsigned long long)strips * ((unsigned long long)(B.read_off[r] - B.read_off[run_begin]) + 32ull);
      int bucket = 63 - __clzll((long long)cost);
      bucket = bucket > 31 ? 31 : bucket;
      const int slot = kr * 32 + bucket;
      if (pass == 0) {
        atomicAdd(S.bucket_count + slot, 1u);
      } else {
        const uint32_t k = S.bucket_base[slot] + atomicAdd(S.bucket_fill + slot, 1u);
        if (k < S.cap[kr]) {
          Task T;
          T.hap = g;
          T.read_begin = run_begin;
          T.read_end = r;
          S.tasks[kr][k] = T;
        }
      }
    }
  }
  if (pass == 0 && (n_bad || n_retry)) {
    atomicAdd(S.n_uncertified, (unsigned long long)(n_bad + n_retry));
    atomicAdd(S.cells_uncertified, cells);
    if (n_retry) atomicAdd(S.n_retried, (unsigned long long)n_retry);
  }
  }
}

// One thread per row class: positions of the cost buckets behind the plan's tasks, heaviest bucket first; new task count.
// Thread 31: segments of the retry lists, {first pair, number of pairs} per band class for the second band round.
__global__ void band_bucket_scan_kernel(const BandCollect S) {
  const int kr = threadIdx.x;
  if (kr == 31 && S.retry_pairs) {
    uint32_t running = 0;
    for (int c = 0; c < kBandClasses; ++c) {
      uint32_t cnt = S.retry_count[c];
      if (running + cnt > S.retry_cap) cnt = S.retry_cap - running;
      S.retry_info[2 * c] = running;
      S.retry_info[2 * c + 1] = cnt;
      running += cnt;
    }
  }
  if (kr > 16) return;
  uint32_t running = *S.count[kr];
  for (int bkt = 31; bkt >= 0; --bkt) {
    S.bucket_base[kr * 32 + bkt] = running;
    running += S.bucket_count[kr * 32 + bkt];
  }
  if (S.cap[kr]) *S.count[kr] = running < S.cap[kr] ? running : S.cap[kr];
}

// After the second band round: a retried pair that is still marked (the containment argument says there is none) becomes
// a one-read task of the stream kernel.
__global__ void band_retry_check_kernel(const VitConsts C, const DevBatch B, const BandCollect S) {
  const uint32_t n = S.retry_info[2 * (kBandClasses - 1)] + S.retry_info[2 * (kBandClasses - 1) + 1];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint2 pr = S.retry_pairs[i];
    const uint32_t g = pr.x, r = pr.y;
    const uint32_t l = B.hap_locus[g];
    const uint32_t hb0 = B.locus_hap_begin[l];
    const uint32_t H = B.locus_hap_begin[l + 1] - hb0;
    const double v = B.out_ll[B.ll_off[l] + (unsigned long long)(r - B.locus_read_begin[l]) * H + (g - hb0)];
    if (!band_marked(v)) continue;
    const int32_t nrow = (int32_t)(B.hap_off[g + 1] - B.hap_off[g]) - 2 * C.cut;
    const int kr = rows_per_lane_hd(nrow, S.kmax);
    const uint32_t k = atomicAdd(S.count[kr], 1u);
    if (k < S.cap[kr]) {
      Task T;
      T.hap = g;
      T.read_begin = r;
      T.read_end = r + 1;
      S.tasks[kr][k] = T;
    }
    atomicAdd(S.n_retry_failed, 1ull);
  }
}

// ------------------------------------------------------------------------------------------------------------------
typedef void (*BandKernel)(const VitConsts, const DevBatch, const BandArgs);

static BandKernel band_kernel_for(int cls, bool sym) {  // cls: band class index (band_class_k / band_class_g)
  switch (cls) {
    case 0: return sym ? viterbi_band_kernel<4, 4, true> : viterbi_band_kernel<4, 4, false>;
    case 1: return sym ? viterbi_band_kernel<3, 8, true> : viterbi_band_kernel<3, 8, false>;
    case 2: return sym ? viterbi_band_kernel<4, 8, true> : viterbi_band_kernel<4, 8, false>;
    case 3: return sym ? viterbi_band_kernel<6, 8, true> : viterbi_band_kernel<6, 8, false>;
    case 4: return sym ? viterbi_band_kernel<8, 8, true> : viterbi_band_kernel<8, 8, false>;
    case 5: return sym ? viterbi_band_kernel<6, 16, true> : viterbi_band_kernel<6, 16, false>;
    case 6: return sym ? viterbi_band_kernel<4, 32, true> : viterbi_band_kernel<4, 32, false>;
    case 7: return sym ? viterbi_band_kernel<5, 32, true> : viterbi_band_kernel<5, 32, false>;
    case 8: return sym ? viterbi_band_kernel<6, 32, true> : viterbi_band_kernel<6, 32, false>;
    case 9: return sym ? viterbi_band_kernel<7, 32, true> : viterbi_band_kernel<7, 32, false>;
    case 10: return sym ? viterbi_band_kernel<8, 32, true> : viterbi_band_kernel<8, 32, false>;
    default: return nullptr;
  }
}

static size_t band_block_smem(int cls) { return (size_t)(kBandBlockThreads / 32) * 6 * band_class_k(cls) * 32 * sizeof(double); }

int band_block_threads() { return kBandBlockThreads; }

// Resident CTAs per SM of band class cls: the smaller of the symmetric-parameter and the general instance.
int band_blocks_per_sm(int cls) {
  BandKernel f_sym = band_kernel_for(cls, true), f_gen = band_kernel_for(cls, false);
  if (!f_sym || !f_gen) return 0;
  int nb_sym = 0, nb_gen = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb_gen, f_gen, kBandBlockThreads, band_block_smem(cls)) != cudaSuccess) return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb_sym, f_sym, kBandBlockThreads, band_block_smem(cls)) != cudaSuccess) return 0;
  return nb_sym < nb_gen ? nb_sym : nb_gen;
}

cudaError_t launch_band(int cls, int grid_blocks, cudaStream_t stream, const VitConsts& C, const DevBatch& B,
                        const BandArgs& A) {
  // symmetric parameters (D2M == I2M, M2I == M2D): two additions per cell fewer, same bits (finish_cell_sym)
  const bool sym = (C.d2m == C.i2m) && (C.m2i == C.m2d);
  BandKernel f = band_kernel_for(cls, sym);
  if (!f) return cudaErrorInvalidValue;
  f<<<grid_blocks, kBandBlockThreads, band_block_smem(cls), stream>>>(C, B, A);
  return cudaGetLastError();
}

cudaError_t launch_band_expand(const BandTask* tasks, const uint32_t* cum, uint32_t n_tasks, uint2* pairs,
                               cudaStream_t stream) {
  if (n_tasks == 0) return cudaSuccess;
  band_expand_kernel<<<(n_tasks + 127) / 128, 128, 0, stream>>>(tasks, cum, n_tasks, pairs);
  return cudaGetLastError();
}

cudaError_t launch_band_collect(const VitConsts& C, const DevBatch& B, const BandTask* tasks, const uint32_t* n_tasks_ptr,
                                uint32_t task_cap, const BandCollect& S, int sm_count, cudaStream_t stream) {
  if (task_cap == 0) return cudaSuccess;
  uint32_t grid = (task_cap + 127u) / 128u;
  const uint32_t grid_max = (uint32_t)sm_count * 16u;
  grid = grid < grid_max ? grid : grid_max;
  band_collect_kernel<<<grid, 128, 0, stream>>>(C, B, tasks, n_tasks_ptr, task_cap, S, 0);
  band_bucket_scan_kernel<<<1, 32, 0, stream>>>(S);
  band_collect_kernel<<<grid, 128, 0, stream>>>(C, B, tasks, n_tasks_ptr, task_cap, S, 1);
  return cudaGetLastError();
}

cudaError_t launch_band_retry_check(const VitConsts& C, const DevBatch& B, const BandCollect& S, int sm_count,
                                    cudaStream_t stream) {
  if (!S.retry_pairs) return cudaSuccess;
  band_retry_check_kernel<<<sm_count * 2, 128, 0, stream>>>(C, B, S);
  return cudaGetLastError();
}

}  // namespace ltr
