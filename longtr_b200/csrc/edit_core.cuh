// edit_core.cuh -- per-lane logic of the thresholded unit-cost edit distance.
//
// HaplotypeGenerator::needleman_wunsch (reference src/SeqAlignment/HaplotypeGenerator.cpp:201-235) fills the
// (n+1) x (m+1) int32 matrix of the unit-cost edit distance (gap 1, mismatch 1, match 0; rows = cent_seq, columns =
// read_seq) and answers T + 1 early when |n - m| > T (:203-206) or when some row i has
//     min over j >= 1 of  dp[i][j] + |(n - m) - (i - j)|  >  T          (:220-231, the minimum starts at 1000).
// With ED the true distance, T < 1000 and n, m >= 1 that function is
//     ED <  T : ED       (a row never fires: the optimal path crosses row i at some column j, where the term is <= ED
//                         for j >= 1, and <= ED + 1 at (i, 1) when the path only touches column 0 of that row)
//     ED >  T : T + 1    (the last row's minimum is dp[n][m] itself)
//     ED == T : T or T + 1, depending on the row test
// and for empty strings: n == 0 -> m (no row is visited), m == 0 < n -> T + 1 (every row's minimum stays 1000).
//
// Two per-lane building blocks, both exact integer arithmetic:
//   * Myers' bit-vector recurrence (block form, Hyyro's horizontal-delta hand-off): one lane owns 32 rows as bit
//     vectors Pv / Mv of vertical deltas; a column costs ~17 word operations for 32 cells.  Gives ED for every pair.
//   * the plain cell recurrence with the reference's row test, 8 rows per lane -- only for the pairs with ED == T.
// The same functions compile for the device and, with LTR_HOST_EMU, for the CPU lane emulator of tests/emu.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__) && !defined(LTR_HOST_EMU)
#define LTE_HD __device__ __forceinline__
#define LTE_COLD static __device__ __noinline__
#else
#define LTE_HD inline
#define LTE_COLD static inline
#endif

namespace ltr {

enum { kEditStripRows = 32 * 32 };  // Myers kernel: rows of one strip (32 lanes x 32-bit words)
enum { kEditDpRows = 8 };           // exact kernel: rows per lane (256 rows per strip)
enum { kEditRowMinStart = 1000 };   // HaplotypeGenerator.cpp:221
enum { kEditMaxThreshold = 999 };   // beyond it the reference's starting minimum changes the function

// 0..3 for the exact bytes 'A','C','T','G' (in that order), 4 for any other byte.
LTE_HD int edit_base_code(int c) {
  const int k = (c >> 1) & 3;
  return ((int)((0x47544341u >> (8 * k)) & 0xffu) == c) ? k : 4;
}

struct MyersLane {
  uint32_t pv, mv;   // vertical deltas +1 / -1 of the lane's 32 rows in the current column
  uint32_t peq[4];   // bit r set: pattern row r of this lane is base code k
};

// Rows row0 .. row0+31 of pattern a[0..n): column 0 of the matrix is dp[i][0] = i (:212-214), all deltas +1.
LTE_HD void myers_lane_load(MyersLane& L, const uint8_t* a, int32_t row0, int32_t n) {
  L.pv = 0xFFFFFFFFu;
  L.mv = 0u;
  L.peq[0] = L.peq[1] = L.peq[2] = L.peq[3] = 0u;
  for (int r = 0; r < 32; ++r) {
    const int32_t i = row0 + r;
    if (i < n) {
      const int code = edit_base_code((int)a[i]);
      if (code == 0) L.peq[0] |= 1u << r;
      if (code == 1) L.peq[1] |= 1u << r;
      if (code == 2) L.peq[2] |= 1u << r;
      if (code == 3) L.peq[3] |= 1u << r;
    }
  }
}

// Rare text bytes (ambiguity codes, lower case): compare the pattern bytes themselves.  Kept out of line so that its
// registers and predicates do not weigh on the hot loop.
LTE_COLD uint32_t myers_eq_any_byte(const uint8_t* a, int32_t row0, int32_t n, int c) {
  uint32_t eq = 0u;
  for (int r = 0; r < 32; ++r)
    if (row0 + r < n && (int)a[row0 + r] == c) eq |= 1u << r;
  return eq;
}

// Match word of text byte c against the lane's rows (the reference compares bytes: 'N' == 'N', 'a' != 'A').
LTE_HD uint32_t myers_eq(const MyersLane& L, const uint8_t* a, int32_t row0, int32_t n, int c) {
  const int code = edit_base_code(c);
  if (code < 4) {
    const uint32_t lo = (code & 1) ? L.peq[1] : L.peq[0];
    const uint32_t hi = (code & 1) ? L.peq[3] : L.peq[2];
    return (code & 2) ? hi : lo;
  }
  return myers_eq_any_byte(a, row0, n, c);
}

// One column of one 32-row block.  hin: horizontal delta (-1, 0, +1) of the row above the block; returns the
// horizontal delta of the block's row `out_bit` (31 = the row handed to the next block).
LTE_HD int myers_block_step(MyersLane& L, uint32_t eq, int hin, int out_bit) {
  const uint32_t pv = L.pv, mv = L.mv;
  const uint32_t hin_neg = (hin < 0) ? 1u : 0u, hin_pos = (hin > 0) ? 1u : 0u;
  const uint32_t xv = eq | mv;
  eq |= hin_neg;
  const uint32_t xh = (((eq & pv) + pv) ^ pv) | eq;
  uint32_t ph = mv | ~(xh | pv);
  uint32_t mh = pv & xh;
  const int hout = (int)((ph >> out_bit) & 1u) - (int)((mh >> out_bit) & 1u);
  ph = (ph << 1) | hin_pos;
  mh = (mh << 1) | hin_neg;
  L.pv = mh | ~(xv | ph);
  L.mv = ph & xv;
  return hout;
}

// ---- exact recurrence with the row test ------------------------------------------------------------------------------
struct EditDpLane {
  int32_t ac[kEditDpRows];      // pattern bytes of the lane's rows (0x100 past the end: matches nothing)
  int32_t left[kEditDpRows];    // dp[i][j-1]
  int32_t rowmin[kEditDpRows];  // running minimum of the row test
  int32_t diag_in;              // dp[i0-1][j-1]
  int32_t bottom;               // dp of the lane's last row in the column just computed (next lane's `top`)
};

// i0: dp row of the lane's first row.
LTE_HD void edit_dp_lane_load(EditDpLane& L, const uint8_t* a, int32_t i0, int32_t n) {
  for (int k = 0; k < kEditDpRows; ++k) {
    const int32_t i = i0 + k;
    L.ac[k] = (i <= n) ? (int32_t)a[i - 1] : 0x100;
    L.left[k] = i;  // dp[i][0] = i (:212-214)
    L.rowmin[k] = kEditRowMinStart;
  }
  L.diag_in = i0 - 1;
  L.bottom = 0;
}

// Column j (>= 1) of the lane's rows; top = dp[i0-1][j]; d = n - m.
LTE_HD void edit_dp_column(EditDpLane& L, int32_t top, int32_t bc, int32_t j, int32_t i0, int32_t d) {
  int32_t up = top, dg = L.diag_in;
  for (int k = 0; k < kEditDpRows; ++k) {
    const int32_t s = (L.ac[k] == bc) ? 0 : 1;
    int32_t v = (up < L.left[k] ? up : L.left[k]) + 1;
    const int32_t w = dg + s;
    v = w < v ? w : v;
    dg = L.left[k];
    L.left[k] = v;
    up = v;
    int32_t t = d - (i0 + k - j);  // :226
    t = t < 0 ? -t : t;
    t += v;
    L.rowmin[k] = t < L.rowmin[k] ? t : L.rowmin[k];
  }
  L.diag_in = top;
  L.bottom = L.left[kEditDpRows - 1];
}

}  // namespace ltr
