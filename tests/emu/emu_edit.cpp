// TEST INFRASTRUCTURE ONLY: host-side lane emulator of the edit-distance / clustering kernels (edit_kernel.cu).
// Drives the SAME per-lane functions the device kernels use (longtr_b200/csrc/edit_core.cuh, compiled with
// LTR_HOST_EMU) through a sequential simulation of one warp: 32 lanes, a skew of one column per lane, the SHFL.UP
// hand-off and the per-warp strip line.  The round structure of the clustering kernels is replayed on top.
#define LTR_HOST_EMU 1
#include <limits.h>
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "edit_core.cuh"

using namespace ltr;

namespace {

// edit_myers_kernel for one pair: returns ED < T ? ED : T + 1; *flag = (ED == T)
int32_t myers_pair(const uint8_t* a0, int32_t n, const uint8_t* b0, int32_t m, int32_t T, int* flag) {
  *flag = 0;
  int32_t d = n - m;
  d = d < 0 ? -d : d;
  if (d > T || n == 0 || m == 0) return (d <= T && n == 0) ? m : T + 1;
  const uint8_t *a = a0, *b = b0;
  if (n < m) {
    std::swap(a, b);
    std::swap(n, m);
  }
  std::vector<int8_t> line((size_t)m + 64, 0);
  int32_t acc[32] = {0};
  for (int32_t r0 = 0; r0 < n; r0 += kEditStripRows) {
    const bool first = (r0 == 0), last_strip = (r0 + kEditStripRows >= n);
    const int32_t rows_here = last_strip ? (n - r0) : (int32_t)kEditStripRows;
    const int32_t t_last = (rows_here - 1) >> 5;
    MyersLane L[32];
    int out_bit[32];
    int32_t pack[32] = {0}, chunk[32] = {0};
    for (int lane = 0; lane < 32; ++lane) {
      myers_lane_load(L[lane], a, r0 + lane * 32, n);
      out_bit[lane] = (last_strip && lane == t_last) ? ((n - 1) & 31) : 31;
    }
    const int32_t n_steps = m + t_last;
    for (int32_t step = 0; step < n_steps; ++step) {
      if ((step & 31) == 0)
        for (int lane = 0; lane < 32; ++lane) {
          const int32_t jj = step + lane;
          chunk[lane] = 0;
          if (jj < m) chunk[lane] = (int32_t)b[jj] | ((first ? 2 : ((int32_t)line[jj] + 1)) << 8);
        }
      const int32_t fresh = chunk[step & 31];
      int32_t prev[32];
      for (int lane = 0; lane < 32; ++lane) prev[lane] = pack[lane];
      for (int lane = 0; lane < 32; ++lane) {
        const int32_t in = lane == 0 ? fresh : prev[lane - 1];
        const int32_t j = step - lane;
        if (lane <= t_last && j >= 0 && j < m) {
          const int c = in & 0xff;
          const uint32_t eq = myers_eq(L[lane], a, r0 + lane * 32, n, c);
          const int hout = myers_block_step(L[lane], eq, (in >> 8) - 1, out_bit[lane]);
          if (lane == t_last) {
            if (last_strip) acc[lane] += hout;
            else line[j] = (int8_t)hout;
          }
          pack[lane] = c | ((hout + 1) << 8);
        }
      }
    }
  }
  const int32_t owner = ((n - 1) & (kEditStripRows - 1)) >> 5;
  const int32_t ed = n + acc[owner];
  *flag = (ed == T);
  return ed > T ? T + 1 : ed;
}

// edit_exact_kernel for one pair
int32_t exact_pair(const uint8_t* a, int32_t n, const uint8_t* b, int32_t m, int32_t T) {
  const int32_t strip = 32 * kEditDpRows;
  std::vector<int32_t> line((size_t)m + 2, 0);
  int32_t result[32] = {0};
  bool fired = false;
  for (int32_t r0 = 0; r0 < n; r0 += strip) {
    const bool first = (r0 == 0), last_strip = (r0 + strip >= n);
    const int32_t rows_here = last_strip ? (n - r0) : strip;
    const int32_t t_last = (rows_here - 1) / kEditDpRows, k_last = (rows_here - 1) % kEditDpRows;
    EditDpLane L[32];
    for (int lane = 0; lane < 32; ++lane) edit_dp_lane_load(L[lane], a, r0 + lane * kEditDpRows + 1, n);
    const int32_t n_steps = m + t_last;
    for (int32_t step = 0; step < n_steps; ++step) {
      int32_t prev[32];
      for (int lane = 0; lane < 32; ++lane) prev[lane] = L[lane].bottom;
      // lane 0 reads the line before the last lane writes it in the same step (device: different entries anyway)
      for (int lane = 0; lane < 32; ++lane) {
        const int32_t i0 = r0 + lane * kEditDpRows + 1;
        int32_t top = lane ? prev[lane - 1] : 0;
        const int32_t j = step - lane + 1;
        if (lane <= t_last && j >= 1 && j <= m) {
          if (lane == 0) top = first ? j : line[j];
          edit_dp_column(L[lane], top, (int32_t)b[j - 1], j, i0, n - m);
          if (lane == t_last) {
            const int32_t vlast = L[lane].left[k_last];
            if (!last_strip) line[j] = vlast;
            else if (j == m) result[lane] = vlast;
          }
        }
      }
    }
    for (int lane = 0; lane < 32; ++lane)
      for (int k = 0; k < kEditDpRows; ++k)
        if (r0 + lane * kEditDpRows + 1 + k <= n && L[lane].rowmin[k] > T) fired = true;
  }
  const int32_t owner = ((n - 1) % strip) / kEditDpRows;
  return fired ? T + 1 : result[owner];
}

}  // namespace

extern "C" int32_t ltr_emu_edit_score(const uint8_t* a, int32_t n, const uint8_t* b, int32_t m, int32_t T, int32_t* flagged) {
  int flag = 0;
  int32_t s = myers_pair(a, n, b, m, T, &flag);
  if (flagged) *flagged = flag;
  if (flag) s = exact_pair(a, n, b, m, T);
  return s;
}

// cluster_*_kernel, one set: the rounds of launch_cluster with the Myers score only (ED == T needs no exact pass: both of
// its possible answers fail `score < T`).
extern "C" int32_t ltr_emu_greedy_cluster(const uint8_t* seq_bytes, const uint32_t* seq_off, const uint32_t* items,
                                          int32_t n_items, int32_t T, int32_t* centroid_of, int32_t* n_centroids_out) {
  std::vector<int32_t> best((size_t)std::max(n_items, 1), INT32_MAX);
  for (int32_t i = 0; i < n_items; ++i) centroid_of[i] = -1;
  *n_centroids_out = n_items > 0 ? 1 : 0;
  if (n_items <= 1) {
    if (n_items == 1) centroid_of[0] = 0;
    return 1;
  }
  centroid_of[0] = 0;
  int32_t cur = 0, n_centroids = 1, state = 0;
  for (int round = 0; round < 15 && state == 0; ++round) {
    uint32_t next = UINT32_MAX;
    for (int32_t i = cur + 1; i < n_items; ++i) {
      const uint32_t sa = items[i], sb = items[cur];
      int flag;
      const int32_t score = myers_pair(seq_bytes + seq_off[sa], (int32_t)(seq_off[sa + 1] - seq_off[sa]),
                                       seq_bytes + seq_off[sb], (int32_t)(seq_off[sb + 1] - seq_off[sb]), T, &flag);
      if (score < T && score < best[i]) {
        best[i] = score;
        centroid_of[i] = cur;
      }
      if (best[i] == INT32_MAX) next = std::min(next, (uint32_t)i);
    }
    if (next == UINT32_MAX) state = 1;
    else {
      cur = (int32_t)next;
      centroid_of[cur] = cur;
      n_centroids += 1;
      if (n_centroids > 15) state = 2;
      else if (cur + 1 >= n_items) state = 1;
    }
  }
  *n_centroids_out = n_centroids;
  return state == 1 ? 1 : 0;
}
