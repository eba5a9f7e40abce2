// stutter_core.cuh -- per-lane logic of the homopolymer / --stutter-align-len path (kernel 2).
//
// Computes what HapAligner::process_read does with short_ == 1 for one (read, haplotype) pair
// (reference: src/SeqAlignment/HapAligner.cpp:855-975): two quality-aware flank alignments
// (align_seq_to_hap_short, :27-163) joined at a seed base (compute_aln_logprob, :165-233), where the repeat
// block collapses into one matrix row that marginalises over PCR stutter artifacts of -6..+6 repeat units
// (StutterAlignerClass::align_stutter_region_reverse, src/SeqAlignment/StutterAlignerClass.cpp:55-166) with the
// reference's approximate log-sum-exp (src/mathops.cpp:98-107; fasterexp/fasterlog bit tricks,
// src/fastonebigheader.h:206-218, 348-357).  Recipe: SURVEY.md Appendix C.
//
// The reference only takes this path for period-1 repeats (HapAligner.cpp:552), so the code is specialised for
// period 1: artifact sizes D = -6..+6 bases, the per-read prefix tables of load_read (:12-53) shrink to
// match[] (kept in shared memory) while the <= 6-term insertion / deletion prefixes are re-summed on the fly
// in the reference's order.  All sums are evaluated in the reference's operation order on the same doubles;
// the only reordering is inside fast_log_sum_exp's accumulation of fasterexp() values, which is exact in
// double (<= 2^12 single-precision terms within 2^11 of each other) and therefore order-independent.
//
// Work split of one warp per (read, haplotype) pair, both flanks one after the other:
//   phase A  flank rows above the repeat block : anti-diagonal wavefront, lane t owns 2 rows, columns streamed
//   phase B  the stutter row                    : one read column per lane, 13 artifact sizes each
//   phase C  flank rows below the repeat block  : wavefront again, seeded by the stutter row
// The same functions compile for the device and, with LTR_HOST_EMU, for the CPU lane emulator of tests/emu.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__) && !defined(LTR_HOST_EMU)
#define LTS_HD __device__ __forceinline__
#define LTS_DEVICE_CODE 1
#else
#define LTS_HD inline
#include <cstring>
#endif

namespace ltr {

static const double kStutImpossible = -1000000000.0;  // HapAligner.cpp:20
enum { kStutRows = 2 };                               // flank rows per lane in the wavefront phases
enum { ROW_OFF = 0, ROW_FIRST = 1, ROW_NORMAL = 2, ROW_AFTER_STUTTER = 3 };

struct StutConsts {
  double i2i, i2m, d2d, d2m, m2m, m2i, m2d;  // (double)(float) AlignmentModel parameters (HapAligner.h:16-22)
  double log_thresh;                         // LOG_THRESH = log(0.001) from the host's libm (mathops.h:36)
  const double* int_logs;                    // int_logs[k] = log(k) from the host's libm, int_logs[0] = -1000
  const double* qual_lc;                     // [256] BaseQuality::log_prob_correct by raw quality byte
  const double* qual_lw;                     // [256] BaseQuality::log_prob_error
};

LTS_HD double smax(double a, double b) { return (a < b) ? b : a; }  // std::max

// ---- fasterexp / fasterlog, single precision, no FMA contraction (fastonebigheader.h:206-218, 348-357) ----
LTS_HD float faster_exp(float p) {
#ifdef LTS_DEVICE_CODE
  const float x = __fmul_rn(1.442695040f, p);
  const float clipp = (x < -126.0f) ? -126.0f : x;
  const float y = __fmul_rn(8388608.0f, __fadd_rn(clipp, 126.94269504f));
  return __uint_as_float(__float2uint_rz(y));
#else
  const float x = 1.442695040f * p;
  const float clipp = (x < -126.0f) ? -126.0f : x;
  const float y = (float)(1 << 23) * (clipp + 126.94269504f);
  const uint32_t u = (uint32_t)y;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
#endif
}
LTS_HD float faster_log(float x) {
#ifdef LTS_DEVICE_CODE
  const float y = __uint2float_rn(__float_as_uint(x));
  return __fadd_rn(__fmul_rn(y, 8.2629582881927490e-8f), -87.989971088f);
#else
  uint32_t u;
  std::memcpy(&u, &x, 4);
  float y = (float)u;
  y *= 8.2629582881927490e-8f;
  return y - 87.989971088f;
#endif
}
LTS_HD float d2f(double v) {
#ifdef LTS_DEVICE_CODE
  return __double2float_rn(v);
#else
  return (float)v;
#endif
}
// One term of fast_log_sum_exp's second pass.
LTS_HD double lse_term(const StutConsts& C, double v, double mx) {
  const double diff = v - mx;
  return (diff > C.log_thresh) ? (double)faster_exp(d2f(diff)) : 0.0;
}
LTS_HD double lse_finish(double mx, double total) { return mx + (double)faster_log(d2f(total)); }

// ---- one flank of one pair, as the warp sees it -------------------------------------------------------
struct FlankView {
  const uint8_t* seq;   // read flank, left to right in alignment order (right flank: already reversed)
  const uint8_t* qual;  // per-column Phred+33 byte, same order
  const double* tlc;    // [256] log P(correct) by quality byte (BaseQuality, src/base_quality.h:29-75): a byte per column
  const double* tlw;    // [256] log P(error)     plus two small tables instead of two double arrays keeps more warps resident
  int32_t L;            // columns
  const uint8_t* blk;   // repeat-block allele in alignment order (reversed for the right flank)
  int32_t B;            // its length (>= 1)
  const int32_t* um;    // um[k*B + pos]: run of matches at lag k+1 ending at block position pos, k < n_del
  int32_t n_del;        // min(6, B)  (StutterAlignerClass.h:64-66)
  const double* match;  // match[p]: load_read's match_probs_ for read position p (StutterAlignerClass.cpp:33-35)
  const double* art_lp; // [13] log_prob_pcr_artifact(allele, D), D = -6..6 (RepeatStutterInfo.h:53-61)
};

LTS_HD double emit_at(const FlankView& F, int32_t p, int32_t c) {
  const double* t = ((int32_t)F.seq[p] == c) ? F.tlc : F.tlw;
  return t[F.qual[p]];
}

// match[p] (load_read): the read walked backwards from p against the block walked backwards from its end.
LTS_HD double stutter_match_prob(const FlankView& F, int32_t p) {
  const int32_t n = (p + 1 < F.B) ? p + 1 : F.B;
  double lp = 0.0;
  for (int32_t j = 0; j < n; ++j) lp += emit_at(F, p - j, (int32_t)F.blk[F.B - 1 - j]);
  return lp;
}

// Visits, in the reference's order, the terms that align_pcr_insertion_reverse / align_pcr_deletion_reverse push
// into log_probs_ for the read segment of base_len bases ending at column j and artifact size D != 0.
template <typename Visit>
LTS_HD void stutter_region_terms(const StutConsts& C, const FlankView& F, int32_t base_len, int32_t j, int32_t D,
                                 Visit& visit) {
  const int32_t B = F.B;
  if (D > 0) {  // StutterAlignerClass.cpp:59-104
    // ins_probs_[offset][D-1]: D read bases ending at j against the block's last base
    double ins = 0.0;
    const int32_t last = (int32_t)F.blk[B - 1];
    const int32_t avail = j + 1;
    for (int32_t q = 0; q < D; ++q)
      if (q < avail) ins += emit_at(F, j - q, last);
    double lp = (-C.int_logs[B + 1] + ins) + ((base_len > D) ? F.match[j - D] : 0.0);
    visit(lp);
    int32_t lim = base_len - D;
    lim = lim < 0 ? 0 : lim;
    lim = lim > B ? B : lim;
    int32_t i = 0;
    for (; i > -lim; --i) {
      if (-i + 1 < B) {
        const int32_t run = F.um[B - 1 + i];  // lag-1 table
        if (run == 0) {
          const int32_t c_old = (int32_t)F.blk[B - 1 + i], c_new = (int32_t)F.blk[B - 2 + i];
          for (int32_t idx = i - 1; idx >= i - D; --idx) {
            lp -= emit_at(F, j + idx, c_old);
            lp += emit_at(F, j + idx, c_new);
          }
          visit(lp);
        } else {
          visit(C.int_logs[run] + lp);
          i -= (run - 1);
        }
      } else {
        visit(lp);
      }
    }
    if (i > -B) visit(C.int_logs[B + i] + lp);
  } else {  // D < 0, StutterAlignerClass.cpp:106-154
    const int32_t* um = F.um + (size_t)(-D - 1) * B;
    double lp = -C.int_logs[B + D + 1];
    if (j - D <= F.L - 1) {
      // match_probs_[offset+D] - del_probs_[offset+D][-D-1]
      const int32_t p = j - D;
      double del = 0.0;
      for (int32_t q = 0; q < -D; ++q) del += emit_at(F, p - q, (int32_t)F.blk[B - 1 - q]);
      lp += F.match[p] - del;
    } else {
      for (int32_t q = 0; q > -base_len; --q) lp += emit_at(F, j + q, (int32_t)F.blk[B - 1 + q + D]);
    }
    visit(lp);
    int32_t i = 0;
    for (; i > -base_len; --i) {
      const int32_t run = um[B - 1 + i];
      if (run == 0) {
        lp -= emit_at(F, j + i, (int32_t)F.blk[B - 1 + i + D]);
        lp += emit_at(F, j + i, (int32_t)F.blk[B - 1 + i]);
        visit(lp);
      } else {
        visit(C.int_logs[run] + lp);
        i -= (run - 1);
      }
    }
    if (-i < B + D) visit(C.int_logs[B + D + i] + lp);
  }
}

struct MaxVisit {
  double mx;
  bool any;
  LTS_HD void operator()(double v) {
    mx = any ? smax(mx, v) : v;
    any = true;
  }
};
struct SumVisit {
  const StutConsts* C;
  double mx, total;
  LTS_HD void operator()(double v) { total += lse_term(*C, v, mx); }
};

// align_stutter_region_reverse for (column j, artifact D).  The reference collects the terms in a vector and
// reduces them with fast_log_sum_exp (max, then sum of fasterexp): two passes over the same term sequence, so that
// no per-thread buffer (local memory) is needed.
LTS_HD double stutter_region_ll(const StutConsts& C, const FlankView& F, int32_t base_len, int32_t j, int32_t D) {
  if (D == 0) return F.match[j];
  MaxVisit mv;
  mv.mx = 0.0;
  mv.any = false;
  stutter_region_terms(C, F, base_len, j, D, mv);
  SumVisit sv;
  sv.C = &C;
  sv.mx = mv.mx;
  sv.total = 0.0;
  stutter_region_terms(C, F, base_len, j, D, sv);
  return lse_finish(mv.mx, sv.total);
}

// ---- all six insertion sizes of a column in one walk -------------------------------------------------------------------
// align_pcr_insertion_reverse (StutterAlignerClass.cpp:59-104) for D = 1..6 walks the block with the SAME lag-1 run table
// and the same skips; only how far it walks (lim_D, non-increasing in D) and how many read bases each position update
// touches (D of them: idx = i-1 .. i-D) differ, and the updates of size D are a prefix of those of size D+1.  One walk
// therefore serves all six: every emission pair is loaded once instead of once per size, and the six running sums are
// independent chains of additions (each the reference's own sequence of operations, so every term is bit-identical).
// visit(d, v): term v of artifact size D = d + 1.
template <typename Visit>
LTS_HD void stutter_insertion_terms(const StutConsts& C, const FlankView& F, int32_t j, Visit& visit) {
  const int32_t B = F.B;
  const int32_t last = (int32_t)F.blk[B - 1];
  const int32_t avail = j + 1;
  double lp[6];
  int32_t lim[6];
  {
    double ins = 0.0;  // ins_probs_[offset][D-1]: D read bases ending at j against the block's last base
#pragma unroll
    for (int32_t d = 0; d < 6; ++d) {
      const int32_t D = d + 1;
      if (d < avail) ins += emit_at(F, j - d, last);
      int32_t base_len = B + D;
      base_len = (base_len < avail) ? base_len : avail;
      lp[d] = (-C.int_logs[B + 1] + ins) + ((base_len > D) ? F.match[j - D] : 0.0);
      visit(d, lp[d]);
      int32_t l = base_len - D;
      l = l < 0 ? 0 : l;
      lim[d] = l > B ? B : l;
    }
  }
  // sizes 1 .. m are still walking (lim is non-increasing in D); a size that stops at position i adds its closing term
  // (every array index below is a compile-time constant after unrolling: the six sums stay in registers)
  int32_t m = 6;
  int32_t i = 0;
#define LTS_CLOSE_FINISHED_SIZES()                                         \
  _Pragma("unroll") for (int32_t d = 5; d >= 0; --d) {                     \
    if (m == d + 1 && !(i > -lim[d])) {                                    \
      if (i > -B) visit(d, C.int_logs[B + i] + lp[d]);                     \
      m = d;                                                               \
    }                                                                      \
  }
  LTS_CLOSE_FINISHED_SIZES()
  while (m > 0) {
    int32_t step = 1;
    if (-i + 1 < B) {
      const int32_t run = F.um[B - 1 + i];  // lag-1 table
      if (run == 0) {
        const int32_t c_old = (int32_t)F.blk[B - 1 + i], c_new = (int32_t)F.blk[B - 2 + i];
#pragma unroll
        for (int32_t k = 1; k <= 6; ++k) {
          if (k <= m) {
            const double eo = emit_at(F, j + i - k, c_old), en = emit_at(F, j + i - k, c_new);
#pragma unroll
            for (int32_t d = 0; d < 6; ++d)
              if (d + 1 >= k && d < m) {
                lp[d] -= eo;
                lp[d] += en;
              }
          }
        }
#pragma unroll
        for (int32_t d = 0; d < 6; ++d)
          if (d < m) visit(d, lp[d]);
      } else {
        const double lr = C.int_logs[run];
#pragma unroll
        for (int32_t d = 0; d < 6; ++d)
          if (d < m) visit(d, lr + lp[d]);
        step = run;
      }
    } else {
#pragma unroll
      for (int32_t d = 0; d < 6; ++d)
        if (d < m) visit(d, lp[d]);
    }
    i -= step;
    LTS_CLOSE_FINISHED_SIZES()
  }
#undef LTS_CLOSE_FINISHED_SIZES
}

struct MaxVisit6 {
  double mx[6];
  uint32_t any;
  LTS_HD void operator()(int32_t d, double v) {
    mx[d] = ((any >> d) & 1u) ? smax(mx[d], v) : v;
    any |= 1u << d;
  }
};
struct SumVisit6 {
  const StutConsts* C;
  double mx[6], total[6];
  LTS_HD void operator()(int32_t d, double v) { total[d] += lse_term(*C, v, mx[d]); }
};
// out[d] = align_stutter_region_reverse(base_len, j, D = d + 1), all six at once (two walks: maxima, then sums).
LTS_HD void stutter_insertion_lls(const StutConsts& C, const FlankView& F, int32_t j, double* out) {
  MaxVisit6 mv;
  mv.any = 0u;
#pragma unroll
  for (int32_t d = 0; d < 6; ++d) mv.mx[d] = 0.0;
  stutter_insertion_terms(C, F, j, mv);
  SumVisit6 sv;
  sv.C = &C;
#pragma unroll
  for (int32_t d = 0; d < 6; ++d) {
    sv.mx[d] = mv.mx[d];
    sv.total[d] = 0.0;
  }
  stutter_insertion_terms(C, F, j, sv);
#pragma unroll
  for (int32_t d = 0; d < 6; ++d) out[d] = lse_finish(mv.mx[d], sv.total[d]);
}

// The stutter row at column j (HapAligner.cpp:79-107): 13 artifact sizes combined by fast_log_sum_exp.
// prevM = match row of the haplotype base preceding the block.
LTS_HD double stutter_row_cell(const StutConsts& C, const FlankView& F, const double* prevM, int32_t j) {
  double probs[13];
  double ins_ll[6];
  stutter_insertion_lls(C, F, j, ins_ll);
#pragma unroll 1
  for (int32_t a = 0; a < 13; ++a) {
    const int32_t D = a - 6;
    int32_t base_len = F.B + D;
    base_len = (base_len < j + 1) ? base_len : (j + 1);
    if (base_len >= 0) {
      const double prob = (D > 0) ? ins_ll[D - 1] : stutter_region_ll(C, F, base_len, j, D);
      const double pre = (j - base_len < 0) ? 0.0 : prevM[j - base_len];
      probs[a] = (F.art_lp[a] + prob) + pre;
    } else {
      probs[a] = kStutImpossible;
    }
  }
  double mx = probs[0];
#pragma unroll
  for (int32_t a = 1; a < 13; ++a) mx = smax(mx, probs[a]);
  double total = 0.0;
#pragma unroll
  for (int32_t a = 0; a < 13; ++a) total += lse_term(C, probs[a], mx);
  return lse_finish(mx, total);
}

// ---- wavefront over flank rows (HapAligner.cpp:36-44 row 0, :112-158 other rows) ----------------------------
struct RowState {
  double M, I, D;  // this row at the previous column
};
struct FlankLane {
  RowState r[kStutRows];
  double upM, upD;    // the row above the lane's first row at the previous column
  double left;        // running sum of lc[] (row 0 only): becomes left_prob
  int32_t type[kStutRows];
  int32_t hc[kStutRows];
  double outM, outD;  // lane's last active row at the column just computed (exported to the lane below)
};

LTS_HD void flank_lane_reset(FlankLane& Ln) {
#pragma unroll
  for (int k = 0; k < kStutRows; ++k) {
    Ln.r[k].M = Ln.r[k].I = Ln.r[k].D = kStutImpossible;
    Ln.type[k] = ROW_OFF;
    Ln.hc[k] = 0xFFFF;
  }
  Ln.upM = Ln.upD = kStutImpossible;
  Ln.left = 0.0;
  Ln.outM = Ln.outD = kStutImpossible;
}

// One column j for the lane's rows.  aboveM/aboveD = the row above the lane's first row at column j.
// Returns M of the lane's rows through Mout[] (the caller records the last column).
LTS_HD void flank_lane_column(FlankLane& Ln, const StutConsts& C, const FlankView& F, int32_t j, double aboveM,
                              double aboveD, double* Mout) {
  const int32_t c = (int32_t)F.seq[j];
  const int32_t qj = (int32_t)F.qual[j];
  const double lcj = F.tlc[qj], lwj = F.tlw[qj];
  double upM = aboveM, upD = aboveD;        // row above, this column
  double ulM = Ln.upM, ulD = Ln.upD;        // row above, previous column
#pragma unroll
  for (int k = 0; k < kStutRows; ++k) {
    const int32_t t = Ln.type[k];
    double M = kStutImpossible, I = kStutImpossible, D = kStutImpossible;
    if (t != ROW_OFF) {
      const double e = (Ln.hc[k] == c) ? lcj : lwj;
      if (t == ROW_FIRST) {  // HapAligner.cpp:36-44
        M = e + Ln.left;
        I = lcj + Ln.left;
        Ln.left += lcj;
      } else if (j == 0) {   // :124-128
        M = e;
        if (t == ROW_NORMAL) {
          I = lcj;
          D = smax(upD + C.d2d, upM + C.d2m);
        }
      } else if (t == ROW_AFTER_STUTTER) {  // :131-138
        M = e + ulM;
      } else {  // :141-156
        const double p0 = Ln.r[k].I + C.m2i, p1 = ulM + C.m2m, p2 = ulD + C.m2d;
        M = e + smax(p0, smax(p1, p2));
        I = lcj + smax(ulM + C.i2m, Ln.r[k].I + C.i2i);
        D = smax(upM + C.d2m, upD + C.d2d);
      }
    }
    Mout[k] = M;
    // this row becomes "the row above" for the next one
    ulM = Ln.r[k].M;
    ulD = Ln.r[k].D;
    Ln.r[k].M = M;
    Ln.r[k].I = I;
    Ln.r[k].D = D;
    upM = M;
    upD = D;
    if (t != ROW_OFF) {
      Ln.outM = M;
      Ln.outD = D;
    }
  }
  Ln.upM = aboveM;
  Ln.upD = aboveD;
}

// Cell-equivalents of one pair for throughput accounting: flank rows count hap_rows x columns,
// the stutter row 13 x block_len per column, both flanks.
#if defined(__CUDACC__)
__host__ __device__
#endif
inline unsigned long long stutter_pair_cells(int32_t n_flank_rows, int32_t B, int32_t read_len) {
  const unsigned long long cols = (unsigned long long)(read_len > 0 ? read_len - 1 : 0);
  return cols * ((unsigned long long)n_flank_rows + 13ull * (unsigned long long)B);
}

}  // namespace ltr
