// viterbi_kernels.cu -- sm_100a kernels of the LongTR read x haplotype Viterbi
// (HapAligner::align_seq_to_hap, reference src/SeqAlignment/HapAligner.cpp:236-343).
//
// One persistent warp per (haplotype, read stream) task; see viterbi_core.cuh for the
// per-lane algorithm.  FP64 max-plus on the FP64 pipe, no tensor cores: this is not a
// contraction.  Per DP cell the fast kernel issues 9 DADD + 4 DSETP (+10 selects); the
// reference recipe counts 17 FP64 ops per cell -- the 4 ops of the
// per-row bail-out test are not evaluated: a pair is certified from its final score
// (viterbi_core.cuh) and the exact kernel (MODE_FULL) re-runs the rest.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"
#include "viterbi_core.cuh"

namespace ltr {

static constexpr unsigned kFullMask = 0xFFFFFFFFu;
static constexpr int kBlockThreads = 128;
// Unroll factor of the fast loop.  Unrolling twice removes the ~4 register moves per cell ptxas puts on the loop's
// back edge, but the loop has to stay inside the ~6 KB L0 instruction cache: K = 9 unrolled twice is 9 KB and
// measured 9 % slower; unrolling only the small row classes (K <= 5) measured -0.3 % on config 3.  Default: none.
#ifndef LTR_FAST_UNROLL_MAX_K
#define LTR_FAST_UNROLL_MAX_K 0
#endif
template <int K>
struct FastUnroll {
  static constexpr int value = (K <= LTR_FAST_UNROLL_MAX_K) ? 2 : 1;
};

template <int K, int MODE>
__device__ __forceinline__ void run_task(const VitConsts& C, const DevBatch& B, const Task& T,
                                         const FailSink& fail, XY* sxy, uint32_t* sb,
                                         unsigned char* wsmem, int lane) {
  const uint32_t g = T.hap;
  const uint32_t hoff = B.hap_off[g];
  const int32_t hlen = (int32_t)(B.hap_off[g + 1] - hoff);
  const uint32_t l = B.hap_locus[g];
  const uint32_t hb0 = B.locus_hap_begin[l];
  const uint32_t H = B.locus_hap_begin[l + 1] - hb0;
  const uint32_t rb0 = B.locus_read_begin[l];
  const unsigned long long out_base = B.ll_off[l] + (g - hb0);
  const int32_t n = hlen - 2 * C.cut;

  if (hlen <= 60 || n < 1) {  // HapAligner.cpp:241-244
    for (uint32_t p = T.read_begin + lane; p < T.read_end; p += 32)
      B.out_ll[out_base + (unsigned long long)(p - rb0) * H] = C.imp;
    return;
  }
  const uint8_t* hap = B.hap_bytes + hoff + C.cut;
  if (n == 1) {  // no DP rows: the result is row 0 of the reference matrices
    for (uint32_t p = T.read_begin + lane; p < T.read_end; p += 32) {
      const uint32_t qb = B.read_off[p];
      const int32_t m = (int32_t)(B.read_off[p + 1] - qb);
      const int32_t dn = n - m;
      double v;
      if ((dn < 0 ? -dn : dn) > 600)
        v = -700.0;
      else
        v = single_row_result(C, m, (m - 1 < n) ? (int32_t)hap[m - 1] : 0, (int32_t)B.read_bytes[qb],
                              (int32_t)hap[0]);
      B.out_ll[out_base + (unsigned long long)(p - rb0) * H] = v;
    }
    return;
  }

  StripCtx S;
  S.hap = hap;
  S.read_bytes = B.read_bytes;
  S.read_off = B.read_off;
  S.out_ll = B.out_ll;
  S.out_base = out_base;
  S.H = H;
  S.rb0 = rb0;
  S.hap_index = g;
  S.qs = B.read_off[T.read_begin];
  S.Q = B.read_off[T.read_end] - S.qs;
  S.n = n;
  S.h0 = (int32_t)hap[0];
  S.fail = fail;
  S.sxy = sxy;
  S.sb = sb;
  S.bnd = reinterpret_cast<XY*>(wsmem);
  S.tx = reinterpret_cast<double*>(wsmem + 2 * 32 * sizeof(XY));
  S.tz = S.tx + 2 * K * 32;
  S.txo = S.tz + 2 * K * 32;

  const StripPlan P = plan_strips(n - 1, K);
  int32_t row_start = 1;
  LaneStream<K> LS;
  for (int s = 0; s < P.strips; ++s) {
    const int32_t rows = P.base + (s < P.rem ? 1 : 0);
    S.first_strip = (s == 0);
    S.last_strip = (s == P.strips - 1);
    int32_t i0, nrows, t_last;
    lane_geometry(K, lane, rows, row_start, i0, nrows, t_last);
    S.t_last = t_last;
    lane_stream_reset<K>(LS, C, S, lane, i0, nrows, T.read_begin);
    // boundary window: positions [0,32) -> bnd[0], positions [32,64) prefetched into registers.  Strip 0 evaluates
    // the row-0 closed form on the fly; later strips read what the previous strip's last lane left in the scratch line.
    BoundaryCursor& bc = reinterpret_cast<BoundaryCursor*>(S.txo + 2 * 32)[lane];  // shared memory, see warp_smem_bytes
    boundary_cursor_reset(bc, S, T.read_begin);
    XY nxt;
    if (s == 0) {
      S.bnd[lane] = boundary_at(C, S, bc, (uint32_t)lane);
      nxt = boundary_at(C, S, bc, 32u + (uint32_t)lane);
    } else {
      S.bnd[lane] = sxy[lane];
      nxt = sxy[32 + lane];
    }
    __syncwarp();
    const uint32_t nsteps = S.Q + (uint32_t)t_last;
    uint32_t step = 0;
    while (step < nsteps) {
      const uint32_t chunk_end = (step + 32u < nsteps) ? step + 32u : nsteps;
      while (step < chunk_end) {
        // event-driven: plain columns for as many steps as every lane of the warp allows, else one general step
        const uint32_t d = lane_plain_distance<K>(LS, S, step - (uint32_t)lane);
        uint32_t nfast = __reduce_min_sync(kFullMask, d);
        nfast = (nfast < chunk_end - step) ? nfast : (chunk_end - step);
        if (nfast > 0u) {
          const uint32_t fast_end = step + nfast;
#pragma unroll FastUnroll<K>::value
          for (; step < fast_end; ++step) {
            const double rx = __shfl_up_sync(kFullMask, LS.L.Xout, 1);
            const double ry = __shfl_up_sync(kFullMask, LS.L.Yout, 1);
            lane_fast_step<K, MODE>(LS, C, S, lane, step - (uint32_t)lane, rx, ry);
          }
        } else {
          const double rx = __shfl_up_sync(kFullMask, LS.L.Xout, 1);
          const double ry = __shfl_up_sync(kFullMask, LS.L.Yout, 1);
          const uint32_t rb = __shfl_up_sync(kFullMask, LS.L.Bout, 1);
          const uint32_t pos = step - (uint32_t)lane;
          if (pos < S.Q) lane_stream_step<K, MODE>(LS, C, S, lane, pos, rx, ry, rb);
          ++step;
        }
      }
      if (step < nsteps) {  // next window of the scratch line: positions [step, step+32)
        S.bnd[((step >> 5) & 1u) * 32u + lane] = nxt;
        nxt = (s == 0) ? boundary_at(C, S, bc, step + 32u + (uint32_t)lane) : sxy[step + 32u + lane];
        __syncwarp();
      }
    }
    row_start += rows;
    __syncwarp();  // strip hand-off through the scratch line: order the last lane's stores before lane 0's loads
  }
}

// ctrl[0] = task cursor, ctrl[1] = number of tasks (for MODE_FULL: the fail count written by
// the MODE_FAST launch that ran before on the same stream).
template <int K, int MODE>
__global__ void __launch_bounds__(kBlockThreads, (K <= 10 ? 4 : (K <= 12 ? 3 : 2)))
viterbi_stream_kernel(const VitConsts C, const DevBatch B, const Task* __restrict__ tasks,
                      const uint32_t* __restrict__ ntasks_ptr, uint32_t task_cap, uint32_t* cursor,
                      const FailSink fail, XY* scratch_xy, uint32_t* scratch_b, uint32_t scratch_stride) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31;
  const uint32_t warp_global = (blockIdx.x * (uint32_t)blockDim.x + threadIdx.x) >> 5;
  uint32_t ntasks = *ntasks_ptr;
  if (ntasks > task_cap) ntasks = task_cap;
  XY* sxy = scratch_xy + (size_t)warp_global * scratch_stride;
  uint32_t* sb = scratch_b + (size_t)warp_global * scratch_stride;
  unsigned char* wsmem = smem + (size_t)(threadIdx.x >> 5) * warp_smem_bytes(K);
  while (true) {
    uint32_t ti = 0;
    if (lane == 0) ti = atomicAdd(cursor, 1u);
    ti = __shfl_sync(kFullMask, ti, 0);
    if (ti >= ntasks) break;
    Task T;
    T.hap = tasks[ti].hap;
    T.read_begin = tasks[ti].read_begin;
    T.read_end = tasks[ti].read_end;
    run_task<K, MODE>(C, B, T, fail, sxy, sb, wsmem, lane);
  }
}

// ------------------------------------------------------------------------------------------
typedef void (*VitKernel)(const VitConsts, const DevBatch, const Task*, const uint32_t*, uint32_t,
                          uint32_t*, const FailSink, XY*, uint32_t*, uint32_t);

template <int MODE>
static VitKernel kernel_for(int k) {
  switch (k) {
#define LTR_CASE(KK) case KK: return viterbi_stream_kernel<KK, MODE>;
    LTR_CASE(1) LTR_CASE(2) LTR_CASE(3) LTR_CASE(4) LTR_CASE(5) LTR_CASE(6) LTR_CASE(7) LTR_CASE(8)
    LTR_CASE(9) LTR_CASE(10) LTR_CASE(11) LTR_CASE(12) LTR_CASE(13) LTR_CASE(14) LTR_CASE(15) LTR_CASE(16)
#undef LTR_CASE
    default: return nullptr;
  }
}

static VitKernel kernel_for_mode(int k, int mode) {
  switch (mode) {
    case MODE_FAST: return kernel_for<MODE_FAST>(k);
    case MODE_FULL: return kernel_for<MODE_FULL>(k);
    case MODE_FAST | MODE_SYM: return kernel_for<MODE_FAST | MODE_SYM>(k);
    case MODE_FULL | MODE_SYM: return kernel_for<MODE_FULL | MODE_SYM>(k);
    default: return nullptr;
  }
}

int viterbi_max_rows_per_lane() { return 16; }

int viterbi_block_threads() { return kBlockThreads; }

static size_t block_smem_bytes(int k) { return (size_t)(kBlockThreads / 32) * warp_smem_bytes(k); }

int viterbi_blocks_per_sm(int k, int mode) {
  // the smaller of the general and the symmetric-parameter instance (the launcher picks one from the parameters)
  int best = 1 << 30;
  for (int sym = 0; sym < 2; ++sym) {
    VitKernel f = kernel_for_mode(k, mode | (sym ? MODE_SYM : 0));
    if (!f) return 0;
    const size_t smem = block_smem_bytes(k);
    if (cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, f, kBlockThreads, smem) != cudaSuccess) return 0;
    best = nb < best ? nb : best;
  }
  return best;
}

// Scratch entries a warp needs for a stream of q bytes (window prefetch reads ahead of the stream).
uint32_t viterbi_scratch_entries(uint32_t q) { return q + 128u; }

cudaError_t launch_viterbi(int k, int mode, int grid_blocks, cudaStream_t stream, const VitConsts& C,
                           const DevBatch& B, const Task* tasks, const uint32_t* ntasks_ptr,
                           uint32_t task_cap, uint32_t* cursor, const FailSink& fail, XY* sxy,
                           uint32_t* sb, uint32_t scratch_stride) {
  // symmetric parameters (D2M == I2M, M2I == M2D): two additions per cell fewer, same bits (finish_cell_sym)
  const bool sym = (C.d2m == C.i2m) && (C.m2i == C.m2d);
  VitKernel f = kernel_for_mode(k, mode | (sym ? MODE_SYM : 0));
  if (!f) return cudaErrorInvalidValue;
  f<<<grid_blocks, kBlockThreads, block_smem_bytes(k), stream>>>(C, B, tasks, ntasks_ptr, task_cap, cursor,
                                                                 fail, sxy, sb, scratch_stride);
  return cudaGetLastError();
}

}  // namespace ltr
