// edit_kernel.cu -- thresholded unit-cost edit distances and greedy clustering on the device.
//
// Replaces HaplotypeGenerator::needleman_wunsch / greedy_clustering (reference
// src/SeqAlignment/HaplotypeGenerator.cpp:201-235, 238-271); arithmetic and case analysis in edit_core.cuh.
//
//   edit_myers_kernel   one warp per pair, Myers' bit-vector recurrence.  The longer string lies along the rows (the
//                       distance is symmetric): lane t owns rows 32t .. 32t+31 of a 1024-row strip as two 32-bit words and
//                       works on column step - t (skew of one column per lane); the horizontal delta of its last row
//                       and the text byte travel to lane t+1 in one SHFL.UP.  Strips hand the deltas of their last row
//                       to the next strip through a per-warp byte line.  ~35 warp instructions per step of up to 1024
//                       cells; HBM traffic is the two strings.  Answers ED < T ? ED : T + 1 and lists the pairs with
//                       ED == T for the exact pass.
//   edit_exact_kernel   the cell recurrence with the reference's row test (8 rows per lane, 256-row strips), only for
//                       the listed pairs: T or T + 1 exactly as the reference decides.
//   cluster_*_kernel    greedy_clustering as <= 15 rounds over ALL sets of a batch at once: round r compares every
//                       sequence behind centroid r of its set with that centroid (the comparisons the reference makes,
//                       regrouped by centroid instead of by sequence), keeps the first minimum below T, and promotes the
//                       first sequence left without a centroid.  No host synchronisation between rounds.
#include <cuda_runtime.h>
#include <stdint.h>

#include "edit_core.cuh"
#include "kernels.h"

namespace ltr {

static constexpr int kEditBlock = 128;
static constexpr unsigned kFull = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t edit_next_pair(uint32_t* cursor, int lane) {
  uint32_t p = 0;
  if (lane == 0) p = atomicAdd(cursor, 1u);
  return __shfl_sync(kFull, p, 0);
}

__global__ void __launch_bounds__(kEditBlock) edit_myers_kernel(const EditArgs A) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t n_pairs = A.n_pairs_ptr ? *A.n_pairs_ptr : A.n_pairs;
  int8_t* line = A.lines ? A.lines + (size_t)warp * A.line_stride : nullptr;
  for (;;) {
    const uint32_t p = edit_next_pair(A.cursor, lane);
    if (p >= n_pairs) break;
    uint32_t ia = A.pair_a[p], ib = A.pair_b[p];
    int32_t n = (int32_t)(A.seq_off[ia + 1] - A.seq_off[ia]), m = (int32_t)(A.seq_off[ib + 1] - A.seq_off[ib]);
    const int32_t T = A.pair_T[p];
    int32_t d = n - m;
    d = d < 0 ? -d : d;
    if (d > T || n == 0 || m == 0) {  // :203-206; empty strings: see edit_core.cuh
      if (lane == 0) A.out[p] = (d <= T && n == 0) ? m : T + 1;
      continue;
    }
    if (n < m) {  // rows = the longer string
      const uint32_t ti = ia; ia = ib; ib = ti;
      const int32_t tn = n; n = m; m = tn;
    }
    const uint8_t* a = A.seq_bytes + A.seq_off[ia];
    const uint8_t* b = A.seq_bytes + A.seq_off[ib];
    int32_t acc = 0;  // sum of the last row's horizontal deltas (lane that owns row n)
    for (int32_t r0 = 0; r0 < n; r0 += kEditStripRows) {
      const bool first = (r0 == 0), last_strip = (r0 + kEditStripRows >= n);
      const int32_t rows_here = last_strip ? (n - r0) : (int32_t)kEditStripRows;
      const int32_t t_last = (rows_here - 1) >> 5;
      const int32_t row0 = r0 + lane * 32;
      MyersLane L;
      myers_lane_load(L, a, row0, n);
      const int out_bit = (last_strip && lane == t_last) ? ((n - 1) & 31) : 31;
      int32_t pack = 0, chunk = 0;
      const int32_t n_steps = m + t_last;
      for (int32_t step = 0; step < n_steps; ++step) {
        if ((step & 31) == 0) {  // text bytes and incoming deltas of the next 32 columns
          const int32_t jj = step + lane;
          chunk = 0;
          if (jj < m) chunk = (int32_t)b[jj] | ((first ? 2 : ((int32_t)line[jj] + 1)) << 8);
          __syncwarp();
        }
        const int32_t fresh = __shfl_sync(kFull, chunk, step & 31);
        const int32_t passed = __shfl_up_sync(kFull, pack, 1);
        const int32_t in = lane == 0 ? fresh : passed;
        const int32_t j = step - lane;
        if (lane <= t_last && j >= 0 && j < m) {
          const int c = in & 0xff;
          const uint32_t eq = myers_eq(L, a, row0, n, c);
          const int hout = myers_block_step(L, eq, (in >> 8) - 1, out_bit);
          if (lane == t_last) {
            if (last_strip) acc += hout;
            else line[j] = (int8_t)hout;
          }
          pack = c | ((hout + 1) << 8);
        }
      }
      __syncwarp();
    }
    const int32_t owner = ((n - 1) & (kEditStripRows - 1)) >> 5;
    const int32_t ed = n + __shfl_sync(kFull, acc, owner);
    if (lane == 0) {
      A.out[p] = ed > T ? T + 1 : ed;
      if (ed == T && A.flagged) A.flagged[atomicAdd(A.n_flagged, 1u)] = p;
    }
  }
}

__global__ void __launch_bounds__(kEditBlock) edit_exact_kernel(const EditArgs A) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t n_list = *A.n_flagged;
  int32_t* line = A.dp_lines + (size_t)warp * A.line_stride;  // dp[last row of the previous strip][1..m]
  for (;;) {
    const uint32_t q = edit_next_pair(A.cursor, lane);
    if (q >= n_list) break;
    const uint32_t p = A.flagged[q];
    const uint32_t ia = A.pair_a[p], ib = A.pair_b[p];
    const uint8_t* a = A.seq_bytes + A.seq_off[ia];
    const uint8_t* b = A.seq_bytes + A.seq_off[ib];
    const int32_t n = (int32_t)(A.seq_off[ia + 1] - A.seq_off[ia]), m = (int32_t)(A.seq_off[ib + 1] - A.seq_off[ib]);
    const int32_t T = A.pair_T[p];
    const int32_t strip = 32 * kEditDpRows;
    int32_t result = 0;
    bool fired = false;
    for (int32_t r0 = 0; r0 < n; r0 += strip) {
      const bool first = (r0 == 0), last_strip = (r0 + strip >= n);
      const int32_t rows_here = last_strip ? (n - r0) : strip;
      const int32_t t_last = (rows_here - 1) / kEditDpRows, k_last = (rows_here - 1) % kEditDpRows;
      const int32_t i0 = r0 + lane * kEditDpRows + 1;
      EditDpLane L;
      edit_dp_lane_load(L, a, i0, n);
      const int32_t n_steps = m + t_last;
      for (int32_t step = 0; step < n_steps; ++step) {
        int32_t top = __shfl_up_sync(kFull, L.bottom, 1);
        const int32_t j = step - lane + 1;  // dp column
        if (lane <= t_last && j >= 1 && j <= m) {
          if (lane == 0) top = first ? j : line[j];  // row 0: dp[0][j] = j (:216-218)
          edit_dp_column(L, top, (int32_t)b[j - 1], j, i0, n - m);
          if (lane == t_last) {
            int32_t vlast = L.left[kEditDpRows - 1];
#pragma unroll
            for (int k = 0; k < kEditDpRows; ++k)
              if (k == k_last) vlast = L.left[k];
            if (!last_strip) line[j] = vlast;
            else if (j == m) result = vlast;
          }
        }
        __syncwarp();  // lane 0's read of line[j] comes before the last lane's later overwrite of the same entry
      }
#pragma unroll
      for (int k = 0; k < kEditDpRows; ++k)
        if (i0 + k <= n && L.rowmin[k] > T) fired = true;  // :228-231
    }
    const int32_t owner = ((n - 1) % strip) / kEditDpRows;
    result = __shfl_sync(kFull, result, owner);
    const bool any_fired = __any_sync(kFull, fired);
    if (lane == 0) A.out[p] = any_fired ? T + 1 : result;
  }
}

static uint32_t edit_grid(uint32_t n_pairs, int sm_count, int blocks_per_sm) {
  const uint32_t warps_per_block = kEditBlock / 32;
  uint32_t blocks = (n_pairs + warps_per_block - 1) / warps_per_block;
  const uint32_t cap = (uint32_t)sm_count * (uint32_t)blocks_per_sm;
  blocks = blocks < cap ? blocks : cap;
  return blocks ? blocks : 1;
}

uint32_t edit_myers_warps(int sm_count) { return (uint32_t)sm_count * 8u * (kEditBlock / 32); }
uint32_t edit_exact_warps(int sm_count) { return (uint32_t)sm_count * 2u * (kEditBlock / 32); }

// pair_cap: upper bound of the number of pairs (the launch does not know a device-side count)
cudaError_t launch_edit_myers(const EditArgs& A, uint32_t pair_cap, int sm_count, cudaStream_t stream) {
  if (pair_cap == 0) return cudaSuccess;
  edit_myers_kernel<<<edit_grid(pair_cap, sm_count, 8), kEditBlock, 0, stream>>>(A);
  return cudaGetLastError();
}

cudaError_t launch_edit_exact(const EditArgs& A, uint32_t pair_cap, int sm_count, cudaStream_t stream) {
  if (pair_cap == 0) return cudaSuccess;
  edit_exact_kernel<<<edit_grid(pair_cap, sm_count, 2), kEditBlock, 0, stream>>>(A);
  return cudaGetLastError();
}

// ---- greedy clustering ---------------------------------------------------------------------------------------------------
// A set is a list of items, an item names a sequence (the same strings may sit in several sets, e.g. one per threshold
// of HaplotypeGenerator.cpp:403).  All indices below are item indices unless they say "sequence".
__global__ void cluster_begin_kernel(const ClusterDev C) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C.n_items) {
    C.best_score[i] = INT32_MAX;
    C.centroid_of[i] = (i == C.set_begin[C.item_set[i]]) ? 0 : -1;  // :240-241: the first sequence is a centroid
  }
  if (i < C.n_sets) {
    const uint32_t b = C.set_begin[i], e = C.set_begin[i + 1];
    C.cur_centroid[i] = b;
    C.next_centroid[i] = UINT32_MAX;
    C.n_centroids[i] = e > b ? 1 : 0;
    C.state[i] = e > b + 1 ? 0 : 1;  // nothing to compare: done
  }
  if (i == 0) {
    *C.n_pairs = 0;
    *C.cursor = 0;
  }
}

__global__ void cluster_pairs_kernel(const ClusterDev C) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C.n_items) return;
  const uint32_t s = C.item_set[i];
  const uint32_t cur = C.cur_centroid[s];
  if (C.state[s] != 0 || i <= cur) return;
  const uint32_t slot = atomicAdd(C.n_pairs, 1u);
  if (C.stat) {  // statistics: comparisons made and their matrix cells, as the reference would fill them
    const uint32_t sa = C.item_seq[i], sb = C.item_seq[cur];
    atomicAdd(C.stat + 0, 1ull);
    atomicAdd(C.stat + 1, (unsigned long long)(C.seq_off[sa + 1] - C.seq_off[sa]) * (unsigned long long)(C.seq_off[sb + 1] - C.seq_off[sb]));
  }
  C.pair_item[slot] = i;
  C.pair_a[slot] = C.item_seq[i];  // needleman_wunsch(seqs[i], centroids[j], ...) (:249)
  C.pair_b[slot] = C.item_seq[cur];
  C.pair_T[slot] = C.set_T[s];
}

__global__ void cluster_update_kernel(const ClusterDev C) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= *C.n_pairs) return;
  const uint32_t i = C.pair_item[q], s = C.item_set[i];
  const int32_t score = C.score[q], T = C.set_T[s];
  if (score < T && score < C.best_score[i]) {  // :252-255, strict: the first minimum wins
    C.best_score[i] = score;
    C.centroid_of[i] = (int32_t)(C.cur_centroid[s] - C.set_begin[s]);
  }
  if (C.best_score[i] == INT32_MAX) atomicMin(&C.next_centroid[s], i);
}

__global__ void cluster_advance_kernel(const ClusterDev C) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s == 0) {
    *C.n_pairs = 0;
    *C.cursor = 0;
  }
  if (s >= C.n_sets || C.state[s] != 0) return;
  const uint32_t nx = C.next_centroid[s];
  C.next_centroid[s] = UINT32_MAX;
  if (nx == UINT32_MAX) {
    C.state[s] = 1;
    return;
  }
  C.cur_centroid[s] = nx;  // :261
  C.centroid_of[nx] = (int32_t)(nx - C.set_begin[s]);
  C.n_centroids[s] += 1;
  if (C.n_centroids[s] > 15) C.state[s] = 2;               // :262-264
  else if (nx + 1 >= C.set_begin[s + 1]) C.state[s] = 1;   // the new centroid is the last item
}

cudaError_t launch_cluster(const ClusterDev& C, const EditArgs& A, int sm_count, cudaStream_t stream) {
  if (C.n_items == 0 && C.n_sets == 0) return cudaSuccess;
  const uint32_t n = C.n_items > C.n_sets ? C.n_items : C.n_sets;
  const uint32_t tb = 256, gb_seq = (n + tb - 1) / tb, gb_set = (C.n_sets + tb - 1) / tb ? (C.n_sets + tb - 1) / tb : 1;
  cluster_begin_kernel<<<gb_seq, tb, 0, stream>>>(C);
  for (int round = 0; round < 15; ++round) {
    cluster_pairs_kernel<<<gb_seq, tb, 0, stream>>>(C);
    edit_myers_kernel<<<edit_grid(C.n_items, sm_count, 8), kEditBlock, 0, stream>>>(A);
    cluster_update_kernel<<<gb_seq, tb, 0, stream>>>(C);
    cluster_advance_kernel<<<gb_set, tb, 0, stream>>>(C);
  }
  return cudaGetLastError();
}

}  // namespace ltr
