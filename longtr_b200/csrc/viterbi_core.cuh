// viterbi_core.cuh -- per-lane logic of the read-stream Viterbi wavefront.
//
// Computes what HapAligner::align_seq_to_hap computes for one (haplotype, read) pair
// (reference: src/SeqAlignment/HapAligner.cpp:236-343; arithmetic recipe in SURVEY.md
// Appendix B), reorganised for a GPU warp:
//
//  * A warp owns one candidate haplotype (one strip of <= 32*K of its rows); lane t keeps
//    K (or K-1) consecutive DP rows in registers.  ALL pooled reads of the locus are
//    streamed through the warp back to back, so the pipeline fill/drain of the wavefront
//    is paid once per haplotype, not once per pair.  Lane t works on stream position
//    s - t at step s (skew of one column per lane).
//  * The three-state recurrence is carried in "pre-added" form.  For a finished cell
//    (i,j) with values M,I,D the lane forms
//        X(i,j) = max(M+M2M, max(D+D2M, I+I2M))    -> diagonal input of cell (i+1,j+1)
//        Y(i,j) = max(M+M2I, I+I2I)                -> upper    input of cell (i+1,j)
//        Z(i,j) = max(M+M2D, D+D2D)                == D(i,j+1)
//    so that  M(i,j) = emit + X(i-1,j-1),  I(i,j) = MATCH + Y(i-1,j),  D(i,j) = Z(i,j-1).
//    These are exactly the additions/maxima of HapAligner.cpp:287-295, evaluated in the
//    same order on the same doubles, hence bit-identical results; only two doubles
//    (X,Y) cross a lane boundary per step and only X,Z persist per row.
//  * Column 0 and row 0 of the reference matrices are closed forms (HapAligner.cpp:263-280)
//    and enter as boundary values (tables tabI/tabD hold the reference's repeatedly-added
//    prefix sums so that even non-integer transition parameters round identically).
//  * Row bail-out (HapAligner.cpp:297-306).  MODE_FULL evaluates it literally.  MODE_FAST
//    does not evaluate it at all: when the final score F of a pair exceeds
//    fast_thr = -600 + slack, no row can fail the test (every row holds a cell of an optimal
//    path whose value minus the band penalty is >= F - slack; proof and the parameter
//    condition it needs in DESIGN.md section 4), so the pair is certified from F alone;
//    the remaining pairs are re-run in MODE_FULL.  Results are exact either way.
//  * The stream loop is event driven: as long as no lane of the warp is at a read start /
//    read end (lane_plain_distance), the warp runs lane_fast_step -- shuffle, one character,
//    one DP column, nothing else -- and falls back to the general lane_stream_step for the
//    ~33 of every |read| steps in which some lane crosses a read boundary.
//  * Nothing in the per-step path is computed by a single lane: the row-0 boundary of
//    the read stream is produced 32 positions at a time by all lanes (boundary_at) into a
//    double-buffered shared-memory window that lane 0 consumes, and the column-0
//    state a lane needs when it starts a read comes from per-lane tables in shared
//    memory (two variants, for the two values of the reference's column-0 emission).
//
// The same code is compiled for the device and, with LTR_HOST_EMU, for a host-side
// lane emulator used ONLY by the CPU unit tests (tests/emu) to check this logic without
// a GPU.  The product never runs the emulator.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__) && !defined(LTR_HOST_EMU)
#define LTR_HD __device__ __forceinline__
#define LTR_HHD __host__ __device__ inline
#define LTR_DEVICE_CODE 1
#else
#define LTR_HD inline
#define LTR_HHD inline
#include <cmath>
#include <cstring>
#endif

namespace ltr {

// Kernel modes.  Bit 0: the per-row bail-out is evaluated literally (MODE_FULL) or certified from the final score
// (MODE_FAST); bit 1 (MODE_SYM): D2M == I2M and M2I == M2D, the cell is evaluated by finish_cell_sym (same bits, two
// additions fewer).  The launcher picks the SYM variants from the parameters.
enum { MODE_FAST = 0, MODE_FULL = 1, MODE_SYM = 2 };
#define LTR_MODE_IS_FULL(MODE) (((MODE) & 1) != 0)

// Parameters shared by every task of a launch (passed by value as a kernel argument).
struct VitConsts {
  double m2m, d2m, i2m, m2i, i2i, m2d, d2d;  // (double)(float) transition log-probs
  double match, mismatch;                    // (double)(float) emissions, HapAligner.cpp:260-261
  double imp;                                // IMPOSSIBLE = -1e9, HapAligner.cpp:20
  double fast_thr;                           // MODE_FAST: a pair with final score > fast_thr cannot bail out
  float d2d_f;                               // float LOG_DEL_TO_DEL for the int*float band term (:298)
  int32_t cut;                               // 35 - INDEL_FLANK_LEN (:245-246)
  const double* tabI;                        // tabI[0] = IMP, tabI[i] = I(i,0)           (:277-279)
  const double* tabD;                        // tabD[0] = IMP, tabD[j] = D(0,j)           (:270-271)
  int32_t tab_len;
};

struct alignas(16) XY {
  double x, y;
};

LTR_HD double vmax(double a, double b) { return (a < b) ? b : a; }  // std::max

LTR_HD uint32_t hi32(double v) {
#ifdef LTR_DEVICE_CODE
  return (uint32_t)__double2hiint(v);
#else
  uint64_t u;
  std::memcpy(&u, &v, 8);
  return (uint32_t)(u >> 32);
#endif
}

LTR_HD float fmul_nofma(float a, float b) {
#ifdef LTR_DEVICE_CODE
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}

LTR_HD double ldg_d(const double* p) {
#ifdef LTR_DEVICE_CODE
  return __ldg(p);
#else
  return *p;
#endif
}

template <int K>
struct Lane {
  // DP state
  double Xin;    // X(row i0-1, previous column): arrived from the lane above one step ago
  double X[K];   // X[r]  = X(row r, previous column), updated in place
  double Z[K];   // Z[r]  = D(row r, current column)
  int32_t hc[K]; // haplotype characters of the lane's rows (0xFFFF = no row)
  double rowmax[K];   // MODE_FULL: max_j (best + band penalty)
  double Xout, Yout;  // X,Y of the lane's last real row at the column just finished
  uint32_t Bout;      // 1 if some row of this pair, in this or an upper lane, is bad
  // geometry (fixed per task/strip)
  int32_t i0;     // DP row index of r = 0
  int32_t nrows;  // number of real rows (K or K-1; 0 for lanes below the haplotype)
  // cursor
  int32_t j;   // column within the current read
  int32_t m;   // length of the current read
  int32_t dn;  // n - m
};

// X,Y,Z of a finished cell.
struct XYZ {
  double x, y, z;
};

LTR_HD XYZ finish_cell(const VitConsts& C, double M, double I, double D) {
  XYZ o;
  o.x = vmax(M + C.m2m, vmax(D + C.d2m, I + C.i2m));
  o.y = vmax(M + C.m2i, I + C.i2i);
  o.z = vmax(M + C.m2d, D + C.d2d);
  return o;
}

// Same values when D2M == I2M and M2I == M2D (Dindel defaults, ONT-like set): max(D + c, I + c) == max(D, I) + c bit
// for bit (rounding is monotone) and M + M2I is M + M2D, so the cell needs 5 additions instead of 7.
LTR_HD XYZ finish_cell_sym(const VitConsts& C, double M, double I, double D) {
  XYZ o;
  const double t = M + C.m2i;
  o.x = vmax(M + C.m2m, vmax(D, I) + C.d2m);
  o.y = vmax(t, I + C.i2i);
  o.z = vmax(t, D + C.d2d);
  return o;
}
template <bool SYM>
LTR_HD XYZ finish_cell_t(const VitConsts& C, double M, double I, double D) {
  return SYM ? finish_cell_sym(C, M, I, D) : finish_cell(C, M, I, D);
}

// Closed-form column 0 of row i (HapAligner.cpp:274-280) for column-0 emission e1.
LTR_HD void col0_cell(const VitConsts& C, int32_t i, double e1, double& Mi, double& Ii, double& Di) {
  const int32_t ia = (i < C.tab_len) ? i : (C.tab_len - 1);  // rows past the haplotype: garbage nobody consumes
  Ii = ldg_d(C.tabI + ia);
  Mi = (ldg_d(C.tabI + ia - 1) + C.i2m) + e1;
  Di = C.imp;
}

// M,I,D of the lane's last real row at the column just computed (consumed at read ends only).
struct LastCell {
  double M, I, D;
};

// ----------------------------------------------------------------------------------
// One DP column (j >= 1) for the lane's rows.  c = read[j]; rx,ry = X(i0-1,j), Y(i0-1,j).
// Pass 1 forms every M of the column from the previous column's X (no dependence between
// rows); pass 2 walks the rows top-down for the I chain and rewrites X,Z in place.
// ----------------------------------------------------------------------------------
template <int K, int MODE>
LTR_HD LastCell lane_column(Lane<K>& L, const VitConsts& C, int32_t c, double rx, double ry) {
  L.j += 1;
  double M[K];
#pragma unroll
  for (int r = 0; r < K; ++r) {
    const double e = (L.hc[r] == c) ? C.match : C.mismatch;
    M[r] = e + (r == 0 ? L.Xin : L.X[r - 1]);
  }
  L.Xin = rx;
  const int32_t d0 = L.dn - L.i0 + L.j;  // diagonal offset (n-m)-(i-j) of row r = 0
  double yup = ry;
  double ya = 0, yb = 0;
  LastCell a, b;
  a.M = a.I = a.D = b.M = b.I = b.D = C.imp;
#pragma unroll
  for (int r = 0; r < K; ++r) {
    const double I = C.match + yup;
    const double D = L.Z[r];
    const XYZ o = finish_cell_t<(MODE & MODE_SYM) != 0>(C, M[r], I, D);
    if (LTR_MODE_IS_FULL(MODE)) {
      // literal HapAligner.cpp:297-298: best + (float)(|(n-m)-(i-j)|) * D2D
      const double best = vmax(D, vmax(I, M[r]));
      int32_t ad = d0 - r;
      ad = ad < 0 ? -ad : ad;
      const float pen = fmul_nofma((float)ad, C.d2d_f);
      const double v = best + (double)pen;
      if (v > L.rowmax[r]) L.rowmax[r] = v;
    }
    L.X[r] = o.x;
    yup = o.y;
    L.Z[r] = o.z;
    if (r == K - 1) { a.M = M[r]; a.I = I; a.D = D; ya = o.y; }
    if (r == K - 2) { b.M = M[r]; b.I = I; b.D = D; yb = o.y; }
  }
  const bool full_lane = (L.nrows == K);
  L.Xout = (K == 1 || full_lane) ? L.X[K - 1] : L.X[K >= 2 ? K - 2 : 0];
  L.Yout = (K == 1 || full_lane) ? ya : yb;
  if (K == 1 || full_lane) return a;
  return b;
}

// True iff one of the lane's real rows fails the per-row test of the finished read.
template <int K, int MODE>
LTR_HD bool lane_rows_bad(const Lane<K>& L) {
  bool bad = false;
#pragma unroll
  for (int r = 0; r < K; ++r) {
    if (r < L.nrows) {
      if (LTR_MODE_IS_FULL(MODE)) bad |= (L.rowmax[r] < -600.0);
    }
  }
  return bad;
}

// Row 0 of the reference matrices as seen by DP row 1 (closed forms, HapAligner.cpp:263-272).
// hj = h[j] (0 when j >= n: SURVEY 8a-1 policy for the reference's out-of-range read),
// c0 = read[0].  Produces X(0,j), Y(0,j).
LTR_HD XY row0_boundary(const VitConsts& C, int32_t j, int32_t hj, int32_t c0) {
  const int32_t jj = (j < C.tab_len) ? j : (C.tab_len - 1);
  const double e = (hj == c0) ? C.match : C.mismatch;
  const double M0 = (j == 0) ? e : ((ldg_d(C.tabD + jj - 1) + C.d2m) + e);
  const double I0 = C.imp;
  const double D0 = ldg_d(C.tabD + jj);
  const XYZ o = finish_cell(C, M0, I0, D0);
  XY b;
  b.x = o.x;
  b.y = o.y;
  return b;
}

// Result of a pair whose haplotype has a single row (n == 1): row 0 at column m-1.
LTR_HD double single_row_result(const VitConsts& C, int32_t m, int32_t h_at_last, int32_t c0, int32_t h0) {
  if (m == 1) return vmax(C.imp, vmax(C.imp, (h0 == c0) ? C.match : C.mismatch));
  const int32_t j = m - 1;
  const int32_t jj = (j < C.tab_len) ? j : (C.tab_len - 1);
  const double M0 = (ldg_d(C.tabD + jj - 1) + C.d2m) + ((h_at_last == c0) ? C.match : C.mismatch);
  return vmax(ldg_d(C.tabD + jj), vmax(C.imp, M0));
}

}  // namespace ltr

// ======================================================================================
// Stream driver: what one lane does at one step of the read stream.
// ======================================================================================
namespace ltr {

struct Task {           // one haplotype against reads [read_begin, read_end) of its locus
  uint32_t hap;
  uint32_t read_begin;
  uint32_t read_end;
};

struct DevBatch {       // flattened batch, device pointers (layout: include/longtr_b200.h)
  const uint8_t* hap_bytes;
  const uint32_t* hap_off;
  const uint32_t* hap_locus;          // [n_haps] locus of each haplotype
  const uint8_t* read_bytes;          // padded with >= 8 readable bytes
  const uint32_t* read_off;
  const uint32_t* locus_hap_begin;
  const uint32_t* locus_read_begin;
  const unsigned long long* ll_off;   // [n_loci+1]
  double* out_ll;
};

struct FailSink {       // pairs that MODE_FAST could not certify; consumed by MODE_FULL
  Task* items;
  uint32_t* count;
  uint32_t capacity;
};

// Per-warp shared memory: boundary window + column-0 tables.
//   bnd [2][32] XY      double-buffered window of the scratch line read by lane 0
//   tx  [2][K][32]      X(row r, column 0) for column-0 emission variant v (0 mismatch, 1 match)
//   tz  [2][K][32]      Z(row r, column 0)
//   txo [2][32]         X of the lane's last real row at column 0
//   cur [32]            per-lane BoundaryCursor (kept out of the register file: it is touched once per 32 steps)
#if defined(__CUDACC__)
__host__ __device__
#endif
inline constexpr size_t warp_smem_bytes(int K) {
  return 2 * 32 * sizeof(XY) + (size_t)(4 * K + 2) * 32 * sizeof(double) + 32 * 4 * sizeof(uint32_t);
}

struct StripCtx {       // warp-uniform description of the strip being streamed
  const uint8_t* hap;   // trimmed haplotype: hap[i] is DP row/column-0 character i
  const uint8_t* read_bytes;
  const uint32_t* read_off;
  double* out_ll;
  unsigned long long out_base;  // ll_off(locus) + hap index within locus
  uint32_t H;                   // haplotypes of the locus (row stride of the LL matrix)
  uint32_t rb0;                 // first pooled read of the locus
  uint32_t hap_index;
  uint32_t qs, Q;               // stream = read_bytes[qs, qs+Q)
  int32_t n, h0;
  int32_t t_last;               // lane owning the strip's last row
  bool first_strip, last_strip;
  XY* sxy;                      // scratch line: boundary (X,Y) entering the strip's first row, per position;
                                // the strip's last lane overwrites it in place for the next strip
  uint32_t* sb;                 // per position: bad flag handed to the next strip (valid at read ends)
  XY* bnd;                      // shared memory (see warp_smem_bytes)
  double* tx;
  double* tz;
  double* txo;
  FailSink fail;
};

template <int K>
struct LaneStream {     // per-lane cursor over the stream
  Lane<K> L;
  int32_t p;            // current read index
  uint32_t qe;          // end offset of the current read
  int32_t cnext;        // prefetched character for the next step
  const uint8_t* rptr;  // address of the character after cnext
};

LTR_HD uint32_t fail_append(const FailSink& F, uint32_t hap, uint32_t read) {
#ifdef LTR_DEVICE_CODE
  const uint32_t k = atomicAdd(F.count, 1u);
#else
  const uint32_t k = (*F.count)++;
#endif
  if (k < F.capacity) {
    F.items[k].hap = hap;
    F.items[k].read_begin = read;
    F.items[k].read_end = read + 1;
  }
  return k;
}

// Row-0 boundary of the stream for strip 0, produced on the fly: every 32 steps each lane evaluates the closed form
// for ONE stream position (window position = lane) straight into the shared-memory window that lane 0 consumes --
// nothing goes through global memory.  The cursor remembers which read the lane's next position falls into
// (positions of a lane grow by 32 per refill, reads are >= 1 base long).
struct alignas(16) BoundaryCursor {
  uint32_t p;       // read index
  uint32_t qb, qe;  // its byte range
  uint32_t pad;
};
LTR_HD void boundary_cursor_reset(BoundaryCursor& bc, const StripCtx& T, uint32_t read_begin) {
  bc.p = read_begin;
  bc.qb = T.read_off[read_begin];
  bc.qe = T.read_off[read_begin + 1];
}
LTR_HD XY boundary_at(const VitConsts& C, const StripCtx& T, BoundaryCursor& slot, uint32_t pos) {
  XY b;
  b.x = C.imp;
  b.y = C.imp;
  if (pos >= T.Q) return b;  // window slots past the end of the stream are never consumed
  const uint32_t q = T.qs + pos;
  BoundaryCursor bc = slot;
  while (q >= bc.qe) {
    bc.p += 1;
    bc.qb = bc.qe;
    bc.qe = T.read_off[bc.p + 1];
  }
  slot = bc;
  const int32_t j = (int32_t)(q - bc.qb);
  const int32_t c0 = (int32_t)T.read_bytes[bc.qb];
  const int32_t hj = (j < T.n) ? (int32_t)T.hap[j] : 0;
  return row0_boundary(C, j, hj, c0);
}

// Per strip: geometry, haplotype characters and the lane's column-0 tables.
template <int K>
LTR_HD void lane_stream_reset(LaneStream<K>& S, const VitConsts& C, const StripCtx& T, int lane,
                              int32_t i0, int32_t nrows, uint32_t read_begin) {
  S.L.i0 = i0;
  S.L.nrows = nrows;
#pragma unroll
  for (int r = 0; r < K; ++r) {
    S.L.hc[r] = (r < nrows) ? (int32_t)T.hap[i0 + r] : 0xFFFF;
    S.L.X[r] = C.imp;
    S.L.Z[r] = C.imp;
    S.L.rowmax[r] = C.imp;
  }
  for (int v = 0; v < 2; ++v) {
    const double e1 = v ? C.match : C.mismatch;
    double xo = C.imp;
#pragma unroll
    for (int r = 0; r < K; ++r) {
      double Mi, Ii, Di;
      col0_cell(C, i0 + r, e1, Mi, Ii, Di);
      const XYZ o = finish_cell(C, Mi, Ii, Di);
      T.tx[(v * K + r) * 32 + lane] = o.x;
      T.tz[(v * K + r) * 32 + lane] = o.z;
      if (r == nrows - 1) xo = o.x;
    }
    T.txo[v * 32 + lane] = xo;
  }
  S.L.Xin = C.imp;
  S.L.Xout = C.imp;
  S.L.Yout = C.imp;
  S.L.Bout = 0;
  S.L.j = 0;
  S.L.m = 0x7FFFFFFF;
  S.L.dn = 0;
  S.p = (int32_t)read_begin - 1;
  S.qe = T.qs;
  S.cnext = (int32_t)T.read_bytes[T.qs];
  S.rptr = T.read_bytes + T.qs + 1;
}

// One step of lane `lane` at stream position pos (0 <= pos < Q).  rx, ry, rbad are the values
// the lane above exported at the previous step; lane 0 takes them from the boundary window.
template <int K, int MODE>
LTR_HD void lane_stream_step(LaneStream<K>& S, const VitConsts& C, const StripCtx& T, int lane,
                             uint32_t pos, double rx, double ry, uint32_t rbad) {
  Lane<K>& L = S.L;
  const uint32_t q = T.qs + pos;
  const int32_t c = S.cnext;
  S.cnext = (int32_t)*S.rptr;
  S.rptr += 1;
  if (lane == 0) {
    const XY b = T.bnd[((pos >> 5) & 1u) * 32u + (pos & 31u)];
    rx = b.x;
    ry = b.y;
  }
  LastCell last;
  last.M = last.I = last.D = C.imp;
  if (q == S.qe) {
    // ---- first character of the next read: column 0 from the lane's tables ------------------
    S.p += 1;
    S.qe = T.read_off[S.p + 1];
    const int32_t m = (int32_t)(S.qe - q);
    const int32_t c1 = (m > 1) ? S.cnext : 0;
    const int v = (T.h0 == c1) ? 1 : 0;  // column-0 emission emit(h[0], r[1]), HapAligner.cpp:276
    L.j = 0;
    L.m = m;
    L.dn = T.n - m;
    L.Xin = rx;
#pragma unroll
    for (int r = 0; r < K; ++r) {
      L.X[r] = T.tx[(v * K + r) * 32 + lane];
      L.Z[r] = T.tz[(v * K + r) * 32 + lane];
      if (LTR_MODE_IS_FULL(MODE)) L.rowmax[r] = C.imp;
    }
    L.Xout = T.txo[v * 32 + lane];
    L.Yout = C.imp;  // Y(.,0) is never consumed: column 0 is not produced by the recurrence
    if (m == 1 && L.nrows > 0)
      col0_cell(C, L.i0 + L.nrows - 1, v ? C.match : C.mismatch, last.M, last.I, last.D);
  } else {
    // ---- DP column j >= 1 ------------------------------------------------------------------
    last = lane_column<K, MODE>(L, C, c, rx, ry);
  }
  if (L.j == L.m - 1) {
    // ---- the lane has finished the current read ---------------------------------------------
    uint32_t bad_in = rbad;
    if (lane == 0) bad_in = T.first_strip ? 0u : T.sb[pos];
    const uint32_t bad = (lane_rows_bad<K, MODE>(L) ? 1u : 0u) | bad_in;
    L.Bout = bad;
    if (lane == T.t_last) {
      if (T.last_strip) {
        double* dst = T.out_ll + T.out_base + (unsigned long long)((uint32_t)S.p - T.rb0) * T.H;
        const int32_t adn = L.dn < 0 ? -L.dn : L.dn;
        if (adn > 600) {
          *dst = -700.0;                       // HapAligner.cpp:249-252
        } else if (LTR_MODE_IS_FULL(MODE)) {
          *dst = bad ? -700.0 : vmax(last.D, vmax(last.I, last.M));      // HapAligner.cpp:300-309
        } else {
          const double F = vmax(last.D, vmax(last.I, last.M));
          *dst = F;
          // certified by the final score alone (needs a column j >= 1, i.e. m >= 2); else exact re-run
          if (!(L.m >= 2 && F > C.fast_thr)) fail_append(T.fail, T.hap_index, (uint32_t)S.p);
        }
      } else {
        T.sb[pos] = bad;
      }
    }
  }
  if (!T.last_strip) {  // warp-uniform
    if (lane == T.t_last) {
      XY o;
      o.x = L.Xout;
      o.y = L.Yout;
      T.sxy[pos] = o;
    }
  }
}

// Number of consecutive steps, starting with the one that processes stream position pos, in which this lane
// only computes plain DP columns (no read start, no read end).  0 for lanes outside the stream.
template <int K>
LTR_HD uint32_t lane_plain_distance(const LaneStream<K>& S, const StripCtx& T, uint32_t pos) {
  if (pos >= T.Q) return 0u;  // not started yet (pos wrapped below 0) or finished
  const uint32_t q = T.qs + pos;
  if (q == S.qe) return 0u;   // first character of a read
  return S.qe - 1u - q;       // the column at qe-1 finishes the read
}

// One plain DP column (the caller guarantees lane_plain_distance >= 1 for every lane of the warp).
template <int K, int MODE>
LTR_HD void lane_fast_step(LaneStream<K>& S, const VitConsts& C, const StripCtx& T, int lane, uint32_t pos,
                           double rx, double ry) {
  const int32_t c = S.cnext;
  S.cnext = (int32_t)*S.rptr;
  S.rptr += 1;
  if (lane == 0) {
    const XY b = T.bnd[((pos >> 5) & 1u) * 32u + (pos & 31u)];
    rx = b.x;
    ry = b.y;
  }
  lane_column<K, MODE>(S.L, C, c, rx, ry);
  if (!T.last_strip) {  // warp-uniform
    if (lane == T.t_last) {
      XY o;
      o.x = S.L.Xout;
      o.y = S.L.Yout;
      T.sxy[pos] = o;
    }
  }
}

// Row class of a haplotype with n DP rows/columns (n = trimmed length): K rows per lane, K <= kmax.
LTR_HHD int rows_per_lane_hd(int n, int kmax) {
  const int R = n - 1;
  if (R <= 32) return 1;
  const int strips = (R + 32 * kmax - 1) / (32 * kmax);
  const int per = (R + strips - 1) / strips;
  const int k = (per + 31) / 32;
  return k < 1 ? 1 : k;
}

// Row split of a task: R = n-1 DP rows over S strips of <= 32*K rows, lanes get K or K-1 rows.
struct StripPlan {
  int32_t strips, base, rem;
};
LTR_HD StripPlan plan_strips(int32_t R, int K) {
  StripPlan P;
  P.strips = (R + 32 * K - 1) / (32 * K);
  if (P.strips < 1) P.strips = 1;
  P.base = R / P.strips;
  P.rem = R % P.strips;
  return P;
}
// Geometry of lane `lane` in a strip of `rows` rows starting at DP row row_start.
LTR_HD void lane_geometry(int K, int lane, int32_t rows, int32_t row_start, int32_t& i0, int32_t& nrows,
                          int32_t& t_last) {
  if (rows >= 32) {
    const int32_t a = rows - 32 * (K - 1);  // lanes [0,a) carry K rows, the others K-1
    nrows = (lane < a) ? K : (K - 1);
    i0 = row_start + lane * (K - 1) + (lane < a ? lane : a);
    t_last = 31;
  } else {
    nrows = (lane < rows) ? 1 : 0;
    i0 = row_start + lane;
    t_last = rows - 1;
  }
}

}  // namespace ltr
