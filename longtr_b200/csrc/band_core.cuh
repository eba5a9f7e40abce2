// band_core.cuh -- per-lane logic of the banded anti-diagonal Viterbi (kernel 1b).
//
// Same function as viterbi_core.cuh -- the score of HapAligner::align_seq_to_hap for one (haplotype, read) pair,
// reference src/SeqAlignment/HapAligner.cpp:236-343 -- evaluated only on the diagonals d = j - i of a band
// [dlo, dlo + W) that contains the origin's diagonal 0 and the end cell's diagonal de = m - n with a margin of w
// diagonals on either side, and accepted only when the result PROVES that the band was wide enough:
//
//  * Every value of the reference's matrices is the maximum over chains of cells (a chain starts at the origin,
//    may run along row 0 or column 0 through the closed-form boundary cells, then through the interior) of the
//    chain's sequentially rounded sum; max and "+ constant" commute under round-to-nearest, so restricting the DP to
//    a band yields exactly max{chains inside the band}.
//  * A chain that leaves the band and still ends in (n-1, m-1) makes at least |de| + 2w + 2 horizontal/vertical
//    moves in at least two runs (one out, one back; runs are separated by M cells because D is only reached from M
//    or D and I from M or I).  Each run pays open = min(|M2D|, |M2I|) for its first move and at least
//    ext = min(|D2D|, |I2I|) for every further one (all transition parameters <= 0, emissions < 0), except for at most
//    one move, the M cell of the boundary row/column (HapAligner.cpp:266-279).  Its value is therefore
//    <= U = -(2 open + (|de| + 2w - 1) ext) (+1e-3 for the rounding of at most ~1e4 additions of magnitude < 1e4).
//  * Hence F_band > U  =>  F = F_band bit for bit; and F > fast_thr additionally certifies that the reference's
//    per-row bail-out (HapAligner.cpp:297-306) cannot fire (viterbi_core.cuh).  Pairs that are not certified are
//    marked and re-run over the full matrix by viterbi_stream_kernel.  Results are exact either way.
//
// Mapping: a group of G lanes (8, or 4 for the narrowest class) owns one pair; lane l keeps 2K consecutive diagonals
// (W = 2 K G).  All lanes advance
// along the anti-diagonals s = i + j in lock step: at step s the lane evaluates its K cells with d = s (mod 2), which
// are independent of each other (X comes from the same diagonal two steps back, Y from diagonal d+1 and Z from
// diagonal d-1 one step back), so one double crosses a lane boundary per step (SHFL.UP on even steps, SHFL.DOWN on
// odd steps).  Steps that may touch a boundary cell, cells outside the matrix or the end cell run band_general_step;
// all others run the branch-free band_fast_even / band_fast_odd pair on characters kept in register windows.
//
// Compiled for the device and, with LTR_HOST_EMU, for the CPU lane emulator of the unit tests (tests/emu).
#pragma once
#include "viterbi_core.cuh"

namespace ltr {

// A pair the band could not certify is marked in the LL matrix with kBandUncertified - F_band (>= 2: log-likelihoods are
// <= 0).  F_band is a lower bound of the pair's true score -- exactly the best chain inside the band -- so it tells which
// wider band is certain to certify the pair (band_retry_class).
static constexpr double kBandUncertified = 2.0;
static constexpr double kBandAbandoned = -4.0e9;     // "F_band" of a pair that was not evaluated at all
LTR_HD double band_mark(double f_band) { return kBandUncertified - f_band; }
LTR_HD bool band_marked(double v) { return v >= kBandUncertified; }
LTR_HD double band_unmark(double v) { return kBandUncertified - v; }

// Band classes: G lanes per pair, K cells per lane per step, W = 2 K G diagonals.  The narrowest class spreads its 32
// diagonals over 4 lanes (8 pairs per warp) so that the loop overhead of a double step is shared by 8 cells.
// The widest classes give half a warp (W = 192) or a whole warp (W = 256 .. 512 in steps of 64) to one pair (noisy
// reads, long repeats): a pair pays for the band of its class, so the steps between classes are what it wastes.
static constexpr int kBandClasses = 11;
LTR_HHD int band_class_k(int c) {
  return c == 0 ? 4 : c == 1 ? 3 : c == 2 ? 4 : c == 3 ? 6 : c == 4 ? 8 : c == 5 ? 6 : c == 6 ? 4 : c == 7 ? 5 : c == 8 ? 6 : c == 9 ? 7 : 8;
}
LTR_HHD int band_class_g(int c) { return c == 0 ? 4 : (c <= 4 ? 8 : (c == 5 ? 16 : 32)); }
LTR_HHD int band_class_w(int c) { return 2 * band_class_k(c) * band_class_g(c); }

struct BandGeom {
  int32_t dlo;  // lowest diagonal of the band (even)
  int32_t w;    // guaranteed margin: [min(0,de) - w, max(0,de) + w] lies inside the band (negative: band too narrow)
};

LTR_HHD BandGeom band_geometry(int32_t n, int32_t m, int32_t W) {
  const int32_t de = m - n;
  const int32_t lo = de < 0 ? de : 0, hi = de < 0 ? 0 : de;
  const int32_t slack = W - 1 - (hi - lo);
  BandGeom g;
  int32_t half = slack / 2;
  if (slack < 0) half = 0;
  g.dlo = lo - half;
  if (g.dlo & 1) g.dlo -= 1;  // even: the step parity of a lane's cells is the same for every pair of a warp
  const int32_t dhi = g.dlo + W - 1;
  const int32_t wl = lo - g.dlo, wh = dhi - hi;
  g.w = (slack < 0) ? -1 : (wl < wh ? wl : wh);
  return g;
}

// Number of interior cells (i >= 1, j >= 1) of the n x m matrix that lie inside the band: what the band kernel has to
// evaluate for the pair (statistics / roofline accounting).
LTR_HHD unsigned long long band_cells(int32_t n, int32_t m, int32_t W, int32_t dlo) {
  const long long dhi = (long long)dlo + W - 1;
  long long total = (long long)(n - 1) * W;
  // rows whose band starts left of column 1: i + dlo < 1  ->  i in [1, min(n-1, -dlo)], each loses 1 - (i + dlo)
  long long a = -(long long)dlo;
  if (a > n - 1) a = n - 1;
  if (a >= 1) total -= a * (1 - (long long)dlo) - a * (a + 1) / 2;
  // rows whose band ends right of column m-1: i + dhi > m-1 -> i in [max(1, m - dhi), n-1], each loses i + dhi - (m-1)
  long long b0 = (long long)m - dhi;
  if (b0 < 1) b0 = 1;
  if (b0 <= n - 1) {
    const long long cnt = (long long)(n - 1) - b0 + 1;
    total -= cnt * (dhi - (m - 1)) + (b0 + (n - 1)) * cnt / 2;
  }
  return total > 0 ? (unsigned long long)total : 0ull;
}

// Cost of the gap moves a chain cannot avoid (see the header comment): every maximal run of horizontal (or of vertical)
// moves is opened from an M cell (D comes from M or D, I from M or I: HapAligner.cpp:289-292), so its first move costs at
// least `open` = min(|M2D|, |M2I|) and every further one at least `ext` = min(|D2D|, |I2I|); open >= ext is enforced.
struct BandGap {
  double open, ext;
};

// Banded evaluation: which pairs go to the band kernel, and with which band class (set up by band_policy(),
// viterbi_host.h; used by the host plan and by the device plan, plan_device.cuh).
struct BandPolicy {
  bool on = false;
  BandGap gap = {0.0, 0.0};  // open = min(|M2D|, |M2I|), ext = min(|D2D|, |I2I|) (open >= ext enforced)
  int w_fixed = 0;           // > 0: margin requested with ltr_ctx_set_band
  double budget0 = 0.0, budget_per_row = 0.0;  // automatic margin: error budget B(n) = budget0 + n * budget_per_row
  int max_share_pct = 80;    // a pair is banded when its band holds at most this share of the full matrix (measured: config 4
                             // 12.3 k loci/s at 55, 12.9 k at 85; config 3 indifferent)
};
LTR_HHD int band_margin_needed(const BandPolicy& bp, int n) {
  if (bp.w_fixed > 0) return bp.w_fixed;
  const double B = bp.budget0 + bp.budget_per_row * (double)n;
  int w = (int)ceil(((B - bp.gap.open) / bp.gap.ext + 1.0) / 2.0);
  w = w < 255 ? w : 255;
  return w > 2 ? w : 2;
}
// Band class of a pair (index into band_class_k) or -1: not banded (too short, band not narrower than ~half the matrix,
// length difference beyond the widest class).
LTR_HHD int band_class_of(int hlen, int n, int m, const BandPolicy& bp) {
  if (!bp.on || hlen <= 60 || n < 2 || m < 2) return -1;
  const int w_need = band_margin_needed(bp, n);
  for (int c = 0; c < kBandClasses; ++c) {
    const int W = band_class_w(c);
    if (band_geometry(n, m, W).w < w_need) continue;
    return ((unsigned long long)(n + m) * (unsigned long long)(W / 2) * 100ull <= (unsigned long long)bp.max_share_pct * (unsigned long long)n * (unsigned long long)m) ? c : -1;
  }
  return -1;
}

// Pair certified from its banded score?  A chain that leaves the band [lo - w, hi + w] and ends on diagonal de makes
// T >= |de| + 2w + 2 gap moves in at least two runs (out and back, an M cell between them); at most one of them is the
// cheap M cell of the boundary row / column (D2M / I2M instead of a gap cost, HapAligner.cpp:266-279), and that one sits in
// a run that still pays its opening.  Its value is therefore <= U = -(2 open + (|de| + 2w - 1) ext) + 1e-3.
// thr = max(fast_thr, U).
LTR_HD double band_threshold(const VitConsts& C, const BandGap& gap, int32_t n, int32_t m, int32_t w) {
  int32_t de = m - n;
  if (de < 0) de = -de;
  // rounding allowance: 1e-3 plus the worst case of ~4(n+m) additions on magnitudes <= ~12(n+m) (only matters for
  // strings of 10^5 bases and more)
  const double len = (double)n + (double)m;
  const double u = -(2.0 * gap.open + gap.ext * (double)(de + 2 * w - 1)) + 1e-3 + 1e-14 * len * len;
  return u > C.fast_thr ? u : C.fast_thr;
}

// Second chance of a pair whose banded score f_band failed the certificate.  The band of a wider class contains the band
// just evaluated (band_geometry centres both on the same diagonals), so its score F' satisfies f_band <= F' <= F; a class
// whose threshold lies below f_band therefore certifies the pair for sure.  Returns the narrowest such class when its band
// is cheaper than rho_pct % of the full matrix, else -1 (full matrix).  f_band comes out of band_unmark: 1e-6 covers the
// rounding of the marker.
LTR_HD int band_retry_class(const VitConsts& C, const BandGap& gap, int32_t n, int32_t m, double f_band, int rho_pct) {
  for (int c = 0; c < kBandClasses; ++c) {
    const int W = band_class_w(c);
    const BandGeom geo = band_geometry(n, m, W);
    if (geo.w < 0) continue;
    if (!(f_band - 1e-6 > band_threshold(C, gap, n, m, geo.w))) continue;
    return ((unsigned long long)(n + m) * (unsigned long long)(W / 2) * 100ull <=
            (unsigned long long)rho_pct * (unsigned long long)n * (unsigned long long)m) ? c : -1;
  }
  return -1;
}

struct BandPair {       // one (haplotype, read) pair as seen by a lane of its group
  const uint8_t* hap;   // trimmed haplotype: hap[i] is the character of DP row i
  const uint8_t* read;  // read[j] is the character of DP column j
  int32_t n, m;
  int32_t d0;           // first of the lane's 2K diagonals
};

template <int K>
struct BandLane {
  double X[2 * K];  // X of the latest cell on local diagonal q
  double A[K];      // Y slots: after an even step A[k] = Y(diag 2k), after an odd step A[k] = Y(diag 2k+1)
  double B[K];      // Z slots: after an even step B[k] = Z(diag 2k), after an odd step B[k] = Z(diag 2k+1)
  int32_t hw[K];      // fast loop: hw[k] = hap[ib - k]
  int32_t rw[K + 1];  // fast loop: rw[k] = read[jb + k]
};

// Closed-form boundary cell of diagonal d (row 0 for d >= 0, column 0 for d < 0): the X, Y, Z interior cells consume
// (HapAligner.cpp:263-280).  Z of a row-0 cell and Y of a column-0 cell only feed other boundary cells: imp.
LTR_HD void band_boundary(const VitConsts& C, const BandPair& R, int32_t d, double& bx, double& by, double& bz) {
  if (d >= 0) {
    const int32_t j = d;
    const int32_t hj = (j < R.n) ? (int32_t)R.hap[j] : 0;
    const XY b = row0_boundary(C, j, hj, (int32_t)R.read[0]);
    bx = b.x;
    by = b.y;
    bz = C.imp;
  } else {
    const int32_t c1 = (R.m > 1) ? (int32_t)R.read[1] : 0;
    const double e1 = ((int32_t)R.hap[0] == c1) ? C.match : C.mismatch;  // emit(h[0], r[1]), HapAligner.cpp:276
    double Mi, Ii, Di;
    col0_cell(C, -d, e1, Mi, Ii, Di);
    const XYZ o = finish_cell(C, Mi, Ii, Di);
    bx = o.x;
    by = C.imp;
    bz = o.z;
  }
}

template <int K>
LTR_HD void band_lane_reset(BandLane<K>& L, const VitConsts& C) {
#pragma unroll
  for (int q = 0; q < 2 * K; ++q) L.X[q] = C.imp;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    L.A[k] = C.imp;
    L.B[k] = C.imp;
    L.hw[k] = 0;
    L.rw[k] = 0;
  }
  L.rw[K] = 0;
}

// One anti-diagonal step with every special case: boundary cells (tb: the lane's closed forms, indexed by local
// diagonal), interior cells by the recurrence with characters fetched from memory, and the end cell (n-1, m-1), whose max(D, max(I, M)) is the pair's score (HapAligner.cpp:308-309).
// P = parity of the step (the lane's cells are the local diagonals q = 2k + P); nb = Z of the lower neighbour's last
// diagonal (P == 0) or Y of the upper neighbour's first diagonal (P == 1), imp at the edges of the band.
template <int K, int P, bool SYM, typename TB>
LTR_HD void band_general_step(BandLane<K>& L, const VitConsts& C, const BandPair& R, const TB& tb, int32_t s,
                              double nb, double& F, bool& got) {
#pragma unroll
  for (int kk = 0; kk < K; ++kk) {
    const int k = (P == 0) ? (K - 1 - kk) : kk;  // order matters: the Y/Z slots are updated in place
    const int q = 2 * k + P;
    const int32_t d = R.d0 + q;
    const int32_t ad = d < 0 ? -d : d;
    const double yin = (P == 0) ? L.A[k] : ((k == K - 1) ? nb : L.A[k + 1 < K ? k + 1 : k]);
    const double zin = (P == 0) ? ((k == 0) ? nb : L.B[k >= 1 ? k - 1 : 0]) : L.B[k];
    const double xin = L.X[q];
    // branch-free: the recurrence is evaluated for every cell (clamped characters) and replaced by imp / the closed form
    // where the cell lies before the matrix / on its boundary -- lanes of a warp are in all three cases at once
    const int32_t i = (s - d) >> 1, j = (s + d) >> 1;
    const int32_t ii = i < 0 ? 0 : (i < R.n ? i : R.n - 1), jj = j < 0 ? 0 : (j < R.m ? j : R.m - 1);
    const double e = ((int32_t)R.hap[ii] == (int32_t)R.read[jj]) ? C.match : C.mismatch;
    const double M = e + xin;
    const double I = C.match + yin;
    const double D = zin;
    const XYZ o = finish_cell_t<SYM>(C, M, I, D);
    // cells before the matrix (s < |d|) keep whatever the recurrence makes of the imp-initialised state: they only feed
    // boundary cells (replaced here) and other cells before the matrix, never an interior cell
    const bool bnd = s == ad;
    const double x = bnd ? tb.x(q) : o.x;
    const double y = bnd ? tb.y(q) : o.y;
    const double z = bnd ? tb.z(q) : o.z;
    if (s > ad && i == R.n - 1 && j == R.m - 1) {
      F = vmax(D, vmax(I, M));
      got = true;
    }
    L.X[q] = x;
    L.A[k] = y;
    L.B[k] = z;
  }
}

// M = X + emit(h, r).  (Two predicated additions instead of select + add were tried: ptxas turns them back into
// selects, 227 instead of 178 instructions per double step for K = 3.)
LTR_HD double band_emit_add(const VitConsts& C, int32_t h, int32_t r, double x) {
  return ((h == r) ? C.match : C.mismatch) + x;
}

// Fill the character windows for an even step s: ib = (s - d0)/2, jb = (s + d0)/2.
template <int K>
LTR_HD void band_windows_init(BandLane<K>& L, const BandPair& R, int32_t s) {
  const int32_t ib = (s - R.d0) >> 1, jb = (s + R.d0) >> 1;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    int32_t i = ib - k;
    i = i < 0 ? 0 : (i < R.n ? i : R.n - 1);
    L.hw[k] = (int32_t)R.hap[i];
  }
#pragma unroll
  for (int k = 0; k <= K; ++k) {
    int32_t j = jb + k;
    j = j < 0 ? 0 : (j < R.m ? j : R.m - 1);
    L.rw[k] = (int32_t)R.read[j];
  }
}

// Plain even step: cells q = 2k (i = ib - k, j = jb + k).  zl = Z of diagonal -1 (lower neighbour lane).
template <int K, bool SYM>
LTR_HD void band_fast_even(BandLane<K>& L, const VitConsts& C, double zl) {
#pragma unroll
  for (int k = K - 1; k >= 0; --k) {
    const double M = band_emit_add(C, L.hw[k], L.rw[k], L.X[2 * k]);
    const double I = C.match + L.A[k];
    const double D = (k == 0) ? zl : L.B[k >= 1 ? k - 1 : 0];
    const XYZ o = finish_cell_t<SYM>(C, M, I, D);
    L.X[2 * k] = o.x;
    L.A[k] = o.y;
    L.B[k] = o.z;
  }
}

// Plain odd step: cells q = 2k+1 (i = ib - k, j = jb + k + 1).  yr = Y of diagonal 2K (upper neighbour lane).
// Then the windows move on by one row and one column; nh, nr are the characters hap[ib + 1], read[jb + K + 1].
template <int K, bool SYM>
LTR_HD void band_fast_odd(BandLane<K>& L, const VitConsts& C, double yr, int32_t nh, int32_t nr) {
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double M = band_emit_add(C, L.hw[k], L.rw[k + 1], L.X[2 * k + 1]);
    const double I = C.match + ((k == K - 1) ? yr : L.A[k + 1 < K ? k + 1 : k]);
    const double D = L.B[k];
    const XYZ o = finish_cell_t<SYM>(C, M, I, D);
    L.X[2 * k + 1] = o.x;
    L.A[k] = o.y;
    L.B[k] = o.z;
  }
#pragma unroll
  for (int k = K - 1; k >= 1; --k) L.hw[k] = L.hw[k - 1];
  L.hw[0] = nh;
#pragma unroll
  for (int k = 0; k < K; ++k) L.rw[k] = L.rw[k + 1];
  L.rw[K] = nr;
}

// Prologue: after a plain step of parity P at step s, the cells that are boundary cells at this step (s == |d|) take
// their closed forms.  Cells before the matrix were computed from the imp-initialised state and garbage characters;
// they are never consumed by an interior cell.
template <int K, int P, typename TB>
LTR_HD void band_fixup(BandLane<K>& L, const BandPair& R, const TB& tb, int32_t s) {
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int q = 2 * k + P;
    const int32_t d = R.d0 + q;
    const int32_t ad = d < 0 ? -d : d;
    if (s == ad) {
      L.X[q] = tb.x(q);
      L.A[k] = tb.y(q);
      L.B[k] = tb.z(q);
    }
  }
}

// First even step from which every cell of the band is interior (s >= |d| + 2 for every diagonal of the band).
LTR_HHD int32_t band_prologue_steps(int32_t dlo, int32_t W) {
  const int32_t dhi = dlo + W - 1;
  const int32_t a = (-dlo > dhi) ? -dlo : dhi;
  const int32_t s = a + 2;
  return s + (s & 1);
}

struct BandTask {  // one haplotype against the unique reads [read_begin, read_end) of its locus, all of one band class
  uint32_t hap;
  uint32_t read_begin;
  uint32_t read_end;
};

}  // namespace ltr
