/* TEST INFRASTRUCTURE ONLY -- see longtr_oracle.h.
 *
 * Plain-C restatement of LongTR's homopolymer / --stutter-align-len path ("short" path):
 *   HapAligner::calc_seed_base            src/SeqAlignment/HapAligner.cpp:467-542
 *   HapAligner::align_seq_to_hap_short    HapAligner.cpp:27-163
 *   StutterAlignerClass::load_read / align_stutter_region_reverse
 *                                         src/SeqAlignment/StutterAlignerClass.cpp:12-166, .h:35-87
 *   RepeatStutterInfo::log_prob_pcr_artifact, StutterModel::log_stutter_pmf
 *                                         RepeatStutterInfo.h:53-61, src/stutter_model.cpp:29-53
 *   HapAligner::compute_aln_logprob       HapAligner.cpp:165-233
 *   fast_log_sum_exp(vector), fasterexp, fasterlog
 *                                         src/mathops.cpp:98-107, src/fastonebigheader.h:206-218, 348-357
 *   BaseQuality tables                    src/base_quality.h:29-75
 * Written from the behaviour of those functions, sequentially and without any of the
 * reference's row-reuse / pointer tricks: every (read, haplotype, flank) matrix is filled from scratch.
 * Build with -ffp-contract=off.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "longtr_oracle.h"

#define S_IMPOSSIBLE (-1000000000.0)
#define S_MIN_SEED_DIST 5

static inline double smax(double a, double b) { return (a < b) ? b : a; }

/* ---- approximate log-sum-exp over a vector (mathops.cpp:98-107) ------------------------------------ */
static float s_fasterexp(float p) {
  float x = 1.442695040f * p;
  float clipp = (x < -126) ? -126.0f : x;
  union { uint32_t i; float f; } v;
  v.i = (uint32_t)((1 << 23) * (clipp + 126.94269504f));
  return v.f;
}
static float s_fasterlog(float x) {
  union { float f; uint32_t i; } vx;
  vx.f = x;
  float y = (float)vx.i;
  y *= 8.2629582881927490e-8f;
  return y - 87.989971088f;
}
static double s_fast_lse(const double* v, int n) {
  const double LOG_THRESH = log(0.001); /* mathops.h:36 */
  double mx = v[0];
  for (int k = 1; k < n; ++k) if (mx < v[k]) mx = v[k];
  double total = 0;
  for (int k = 0; k < n; ++k) {
    double diff = v[k] - mx;
    if (diff > LOG_THRESH) total += s_fasterexp((float)diff);
  }
  return mx + s_fasterlog((float)total);
}
static double s_int_log(int v) { return v == 0 ? -1000.0 : log((double)v); } /* mathops.cpp:16-22 */

/* ---- stutter model (stutter_model.cpp:29-53; the model held by a RepeatBlock is a copy whose period is the
 *      motif length, stutter_model.h:72) ---------------------------------------------------------------- */
typedef struct {
  double in_nostep, in_step, in_up, in_down, equal, out_nostep, out_step, out_up, out_down;
  int motif_len;
} s_model;
static void s_model_init(s_model* m, const double st[6], const char* motif) {
  m->in_step = log(1 - st[0]); m->in_nostep = log(st[0]);
  m->in_up = log(st[1]); m->in_down = log(st[2]);
  m->out_step = log(1 - st[3]); m->out_nostep = log(st[3]);
  m->out_up = log(st[4]); m->out_down = log(st[5]);
  m->equal = log(1 - st[1] - st[2] - st[4] - st[5]);
  m->motif_len = (int)strlen(motif);
}
static double s_pmf(const s_model* m, int sample_bps, int read_bps) {
  int d = read_bps - sample_bps;
  if (d % m->motif_len != 0) {
    int e = d - (d / m->motif_len);
    if (e < 0) return m->out_down + m->out_nostep + m->out_step * (-e - 1);
    return m->out_up + m->out_nostep + m->out_step * (e - 1);
  }
  int r = d / m->motif_len;
  if (r == 0) return m->equal;
  if (r < 0) return m->in_down + m->in_nostep + m->in_step * (-r - 1);
  return m->in_up + m->in_nostep + m->in_step * (r - 1);
}
static double s_pcr_artifact(const s_model* m, int period, int allele_len, int D) { /* RepeatStutterInfo.h:53-61 */
  const int max_ins = 6 * period, max_del = -6 * period;
  int read_size = allele_len + D;
  if (D == 0) return s_pmf(m, allele_len, read_size);
  if (D > 0) return D > max_ins ? -10e6 : s_pmf(m, allele_len, read_size);
  return (D < max_del || read_size < 0) ? -10e6 : s_pmf(m, allele_len, read_size);
}

/* ---- StutterAlignerClass for one allele and one read flank ------------------------------------------ */
typedef struct {
  const char* blk; /* block sequence, FORWARD indexing: blk[0..B-1] */
  int B, period, n_ins, n_del, max_ins, max_del;
  int** um;        /* um[k][pos]: run of matches at lag (k+1)*period ending at pos (StutterAlignerClass.h:35-42) */
  int n_um;
  /* per-read tables (load_read), indexed by read position p (the reference indexes by offset = L-1-p) */
  double *match, *ins, *del;
  const char* seq; const double *lw, *lc; int L;
} s_aligner;

static int* s_upstream(const char* s, int n, int lag) {
  int* ml = (int*)calloc((size_t)(n > 0 ? n : 1), sizeof(int));
  for (int i = lag; i < n; ++i) ml[i] = (s[i - lag] != s[i]) ? 0 : 1 + ml[i - 1];
  return ml;
}
static void s_aligner_init(s_aligner* A, const char* blk, int B, int period) {
  memset(A, 0, sizeof(*A));
  A->blk = blk; A->B = B; A->period = period;
  A->n_ins = 6; A->n_del = 6;
  while (A->n_del * period > B) A->n_del--;
  A->max_ins = period * A->n_ins; A->max_del = -period * A->n_del;
  A->n_um = A->n_del > 0 ? A->n_del : 1; /* max_deletion_ == 0: one table at lag = period (.h:72-73) */
  A->um = (int**)calloc((size_t)A->n_um, sizeof(int*));
  for (int k = 0; k < A->n_um; ++k) A->um[k] = s_upstream(blk, B, (k + 1) * period);
}
static void s_aligner_free(s_aligner* A) {
  for (int k = 0; k < A->n_um; ++k) free(A->um[k]);
  free(A->um); free(A->match); free(A->ins); free(A->del);
}
static inline double s_emit(const s_aligner* A, int p, char c) { return (A->seq[p] == c) ? A->lc[p] : A->lw[p]; }

/* load_read (StutterAlignerClass.cpp:12-53): emission sums of the read, walking backwards from position p,
 * against the block walking backwards from its last base. */
static void s_load_read(s_aligner* A, const char* seq, const double* lw, const double* lc, int L) {
  free(A->match); free(A->ins); free(A->del);
  A->seq = seq; A->lw = lw; A->lc = lc; A->L = L;
  A->match = (double*)calloc((size_t)L, sizeof(double));
  A->ins = (double*)calloc((size_t)L * A->n_ins, sizeof(double));
  A->del = (double*)calloc((size_t)L * (A->n_del > 0 ? A->n_del : 1), sizeof(double));
  const int B = A->B, per = A->period;
  for (int p = L - 1; p >= 0; --p) {
    const int avail = p + 1; /* read bases at or left of p */
    double lp = 0.0;
    int j, k = 0;
    for (j = 0; j < (avail < -A->max_del ? avail : -A->max_del); ++j) {
      lp += s_emit(A, p - j, A->blk[B - 1 - j]);
      if ((j + 1) % per == 0) A->del[(size_t)p * A->n_del + k++] = lp;
    }
    for (; j < (avail < B ? avail : B); ++j) lp += s_emit(A, p - j, A->blk[B - 1 - j]);
    A->match[p] = lp;
    double li = 0.0;
    k = 0;
    for (j = 0; j < (A->max_ins < avail ? A->max_ins : avail); ++j) {
      if (j % per < B) li += s_emit(A, p - j, A->blk[B - 1 - (j % per)]);
      else li += lc[p - j];
      if ((j + 1) % per == 0) A->ins[(size_t)p * A->n_ins + k++] = li;
    }
    for (; j < A->max_ins; ++j)
      if ((j + 1) % per == 0) A->ins[(size_t)p * A->n_ins + k++] = li;
  }
}

/* align_stutter_region_reverse (StutterAlignerClass.cpp:55-166): likelihood of the read segment of base_len
 * bases ending at read position j, given a PCR artifact of D bases somewhere in the block. */
static double s_align_region(const s_aligner* A, int base_len, int j, int D, double* terms) {
  const int B = A->B, per = A->period;
  if (D == 0) return A->match[j];
  int nt = 0;
  if (D > 0) { /* :59-104 */
    const int* um = A->um[0];
    double lp = -s_int_log(B + 1) + A->ins[(size_t)j * A->n_ins + D / per - 1] + (base_len > D ? A->match[j - D] : 0);
    terms[nt++] = lp;
    int lim = base_len - D; if (lim < 0) lim = 0; if (lim > B) lim = B;
    int i = 0;
    for (; i > -lim; i--) {
      if (-i + per < B) {
        const int run = um[B - 1 + i];
        if (run == 0) {
          for (int idx = i - per; idx >= i - D; idx -= per) {
            lp -= s_emit(A, j + idx, A->blk[B - 1 + i]);
            lp += s_emit(A, j + idx, A->blk[B - 1 + i - per]);
          }
          terms[nt++] = lp;
        } else {
          terms[nt++] = s_int_log(run) + lp;
          i -= (run - 1);
        }
      } else
        terms[nt++] = lp;
    }
    if (i > -B) terms[nt++] = s_int_log(B + i) + lp;
    return s_fast_lse(terms, nt);
  }
  /* D < 0, :106-154 */
  const int* um = A->um[-D / per - 1];
  double lp = -s_int_log(B + D + 1);
  if (j - D <= A->L - 1) /* offset + D >= 0 */
    lp += A->match[j - D] - A->del[(size_t)(j - D) * A->n_del + (-D / per - 1)];
  else
    for (int q = 0; q > -base_len; q--) lp += s_emit(A, j + q, A->blk[B - 1 + q + D]);
  terms[nt++] = lp;
  int i;
  for (i = 0; i > -base_len; i--) {
    const int run = um[B - 1 + i];
    if (run == 0) {
      lp -= s_emit(A, j + i, A->blk[B - 1 + i + D]);
      lp += s_emit(A, j + i, A->blk[B - 1 + i]);
      terms[nt++] = lp;
    } else {
      terms[nt++] = s_int_log(run) + lp;
      i -= (run - 1);
    }
  }
  if (-i < B + D) terms[nt++] = s_int_log(B + D + i) + lp;
  return s_fast_lse(terms, nt);
}

/* ---- one flank against one haplotype (HapAligner.cpp:27-163) ---------------------------------------------
 * blocks: b0 (flank), b1 (repeat allele), b2 (flank).  Fills M[hap_row*L + j] for every row the reference
 * fills (rows strictly inside the stutter block stay untouched) and returns left_prob = sum_j lc[j].     */
typedef struct { float i2i, i2m, d2d, d2m, m2m, m2i, m2d; } s_aln;

static double s_align_flank(const char* b0, int n0, const char* b1, int n1, const char* b2, int n2, int period,
                            const s_model* model, int allele_len_for_pmf, const s_aln* P, const char* seq,
                            const double* lw, const double* lc, int L, double* M) {
  const int hapsize = n0 + n1 + n2;
  double* I = (double*)malloc(sizeof(double) * (size_t)hapsize * L);
  double* Dm = (double*)malloc(sizeof(double) * (size_t)hapsize * L);
  double left = 0.0;
  for (int j = 0; j < L; ++j) { /* row 0, :36-44 */
    M[j] = ((seq[j] == b0[0]) ? lc[j] : lw[j]) + left;
    I[j] = lc[j] + left;
    Dm[j] = S_IMPOSSIBLE;
    left += lc[j];
  }
  int row = 1, stutter_R = -1;
  for (int blk = 0; blk < 3; ++blk) {
    const char* bs = blk == 0 ? b0 : (blk == 1 ? b1 : b2);
    const int bn = blk == 0 ? n0 : (blk == 1 ? n1 : n2);
    if (blk == 1) { /* stutter block collapses into its last row, :64-111 */
      s_aligner A;
      s_aligner_init(&A, bs, bn, period);
      s_load_read(&A, seq, lw, lc, L);
      double* terms = (double*)malloc(sizeof(double) * (size_t)(bn + 4));
      const double* prevM = M + (size_t)(row - 1) * L;
      double* outM = M + (size_t)(row + bn - 1) * L;
      for (int j = 0; j < L; ++j) {
        double probs[13];
        int na = 0;
        for (int D = -6 * period; D <= 6 * period; D += period) {
          int base_len = bn + D < j + 1 ? bn + D : j + 1;
          if (base_len >= 0) {
            double prob = s_align_region(&A, base_len, j, D, terms);
            double pre = (j - base_len < 0) ? 0 : prevM[j - base_len];
            probs[na] = s_pcr_artifact(model, period, allele_len_for_pmf, D) + prob + pre;
          } else
            probs[na] = S_IMPOSSIBLE;
          na++;
        }
        outM[j] = s_fast_lse(probs, na);
        I[(size_t)(row + bn - 1) * L + j] = S_IMPOSSIBLE;
        Dm[(size_t)(row + bn - 1) * L + j] = S_IMPOSSIBLE;
      }
      free(terms);
      s_aligner_free(&A);
      stutter_R = row + bn - 1;
      row += bn;
      continue;
    }
    for (int c = (blk == 0 ? 1 : 0); c < bn; ++c, ++row) { /* flank rows, :112-158 */
      const char hc = bs[c];
      double* m = M + (size_t)row * L; double* in = I + (size_t)row * L; double* de = Dm + (size_t)row * L;
      const double* mu = m - L; const double* iu = in - L; const double* du = de - L;
      (void)iu;
      const int after_stutter = (row == stutter_R + 1);
      m[0] = (seq[0] == hc) ? lc[0] : lw[0];
      in[0] = after_stutter ? S_IMPOSSIBLE : lc[0];
      de[0] = after_stutter ? S_IMPOSSIBLE : smax(du[0] + P->d2d, mu[0] + P->d2m);
      if (after_stutter) { /* a stutter block must be followed by a match, :129-138 */
        for (int j = 1; j < L; ++j) {
          m[j] = ((seq[j] == hc) ? lc[j] : lw[j]) + mu[j - 1];
          in[j] = S_IMPOSSIBLE;
          de[j] = S_IMPOSSIBLE;
        }
        continue;
      }
      for (int j = 1; j < L; ++j) { /* :141-156 (parameter names as the reference uses them) */
        const double p0 = in[j - 1] + P->m2i, p1 = mu[j - 1] + P->m2m, p2 = du[j - 1] + P->m2d;
        const double e = (seq[j] == hc) ? lc[j] : lw[j];
        m[j] = e + smax(p0, smax(p1, p2));
        in[j] = lc[j] + smax(mu[j - 1] + P->i2m, in[j - 1] + P->i2i);
        de[j] = smax(mu[j] + P->d2m, du[j] + P->d2d);
      }
    }
  }
  free(I); free(Dm);
  return left;
}

/* ---- seed selection (HapAligner.cpp:467-542) -------------------------------------------------------- */
static void s_best_seed_pos(int32_t rs, int32_t re, int32_t rep_start, int32_t rep_end, int32_t* best_dist, int32_t* best_pos) {
  *best_dist = *best_pos = -1;
  int32_t pos = rs;
  int k = 0;
  while (k < 1 && pos <= re) {
    if (pos < rep_start) {
      int32_t lim = re < rep_start - 1 ? re : rep_start - 1;
      int32_t dist = 1 + (lim - pos) / 2;
      if (dist >= *best_dist) { *best_dist = dist; *best_pos = dist - 1 + pos; }
      pos = rep_end; k++;
    } else if (pos < rep_end) { pos = rep_end; k++; }
    else k++;
  }
  if (pos <= re) {
    int32_t dist = 1 + (re - pos) / 2;
    if (dist >= *best_dist) { *best_dist = dist; *best_pos = dist - 1 + pos; }
  }
}

int32_t ltr_oracle_seed_base(const ltr_flat_locus* L, int32_t read_index) {
  const ltr_flat_read* R = &L->reads[read_index];
  const int32_t first = L->repeat_start - (int32_t)strlen(L->lflank), last = L->repeat_end + (int32_t)strlen(L->rflank);
  int32_t pos = R->start;
  int best_seed = -1, cur = 0, max_dist = S_MIN_SEED_DIST, num = 0;
  for (const char* c = R->cigar; *c; ++c) {
    if (*c >= '0' && *c <= '9') { num = num * 10 + (*c - '0'); continue; }
    switch (*c) {
      case '=': {
        int32_t lo = pos > first ? pos : first, hi = pos + num - 1 < last - 1 ? pos + num - 1 : last - 1;
        if (lo <= hi) {
          int32_t d, dp;
          s_best_seed_pos(lo, hi, L->repeat_start, L->repeat_end, &d, &dp);
          if (d >= max_dist) { max_dist = d; best_seed = cur + (dp - pos); }
        }
        pos += num; cur += num; break;
      }
      case 'I': cur += num; break;
      case 'X': pos += num; cur += num; break;
      case 'D': pos += num; break;
      default: return -2;
    }
    num = 0;
  }
  if (best_seed < -1 || best_seed == 0 || best_seed >= (int)strlen(R->seq) - 1) return -1;
  return best_seed;
}

/* ---- process_read, short_ == 1 (HapAligner.cpp:855-975) for one read ---------------------------------- */
static void s_reverse(char* s, int n) { for (int i = 0; i < n / 2; ++i) { char t = s[i]; s[i] = s[n - 1 - i]; s[n - 1 - i] = t; } }
static char* s_revdup(const char* s, int n) { char* r = (char*)malloc((size_t)n + 1); memcpy(r, s, (size_t)n); r[n] = 0; s_reverse(r, n); return r; }

int ltr_oracle_process_read_short(const ltr_flat_locus* L, int32_t read_index, int32_t seed, double* out_row) {
  const ltr_flat_read* R = &L->reads[read_index];
  const int N = (int)strlen(R->seq);
  if (seed < 1 || seed >= N - 1) return -1;
  /* base_quality.h:29-75 */
  double lcq[256], lwq[256];
  const int maxq = 'J' - '!';
  lcq[0] = -100; lwq[0] = 0;
  for (int i = 1; i <= maxq; ++i) { lcq[i] = log(1.0 - pow(10.0, i / (-10.0))); lwq[i] = log(pow(10.0, i / (-10.0) / 5.0)); }
  double* lw = (double*)malloc(sizeof(double) * N);
  double* lc = (double*)malloc(sizeof(double) * N);
  for (int j = 0; j < N; ++j) {
    const char q = R->qual[j];
    int qi = q - '!';
    if (q < '!') qi = 0; else if (q > 'J') qi = maxq;
    lw[j] = lwq[qi]; lc[j] = lcq[qi];
  }
  const int Lf = seed, Rf = N - seed - 1;
  /* right flank: reversed read suffix and reversed qualities (:887-890) */
  char* rseq = s_revdup(R->seq + seed + 1, Rf);
  double* rlw = (double*)malloc(sizeof(double) * Rf);
  double* rlc = (double*)malloc(sizeof(double) * Rf);
  for (int j = 0; j < Rf; ++j) { rlw[j] = lw[N - 1 - j]; rlc[j] = lc[N - 1 - j]; }
  s_aln P;
  ltr_oracle_params dp;
  ltr_oracle_default_params(&dp);
  if (L->n_aln_params == 7) {
    dp.ins_ins = L->aln_params[0]; dp.ins_match = L->aln_params[1]; dp.del_del = L->aln_params[2];
    dp.del_match = L->aln_params[3]; dp.match_match = L->aln_params[4]; dp.match_ins = L->aln_params[5];
    dp.match_del = L->aln_params[6];
  }
  P.i2i = dp.ins_ins; P.i2m = dp.ins_match; P.d2d = dp.del_del; P.d2m = dp.del_match;
  P.m2m = dp.match_match; P.m2i = dp.match_ins; P.m2d = dp.match_del;
  s_model model;
  s_model_init(&model, L->stutter, L->motif);
  const int n0 = (int)strlen(L->lflank), n2 = (int)strlen(L->rflank);
  char* rev_l = s_revdup(L->lflank, n0);
  char* rev_r = s_revdup(L->rflank, n2);
  int max_allele = 0;
  for (int a = 0; a < L->n_alleles; ++a) { int al = (int)strlen(L->alleles[a]); if (al > max_allele) max_allele = al; }
  const int max_hap = n0 + n2 + max_allele;
  double* Lm = (double*)malloc(sizeof(double) * (size_t)max_hap * Lf);
  double* Rm = (double*)malloc(sizeof(double) * (size_t)max_hap * Rf);
  int rc = 0;
  for (int a = 0; a < L->n_alleles; ++a) {
    if (L->realign_to_hap && !L->realign_to_hap[a]) continue;
    const char* al = L->alleles[a];
    const int n1 = (int)strlen(al);
    if (n1 == 0) { rc = -4; break; } /* empty allele: the reference's stutter row would overwrite its predecessor */
    char* rev_a = s_revdup(al, n1);
    const int hapsize = n0 + n1 + n2;
    const double l_prob = s_align_flank(L->lflank, n0, al, n1, L->rflank, n2, L->period, &model, n1, &P, R->seq, lw, lc, Lf, Lm);
    const double r_prob = s_align_flank(rev_r, n2, rev_a, n1, rev_l, n0, L->period, &model, n1, &P, rseq, rlw, rlc, Rf, Rm);
    free(rev_a);
    /* compute_aln_logprob (:165-233) */
    const char seed_char = R->seq[seed];
    const double sw = lw[seed], sc = lc[seed];
    const double prior = -s_int_log(n0 + n2);
    double* terms = (double*)malloc(sizeof(double) * (size_t)(hapsize + 2));
    int nt = 0;
    terms[nt++] = prior + (seed_char == L->lflank[0] ? sc : sw) + l_prob + Rm[(size_t)Rf * (hapsize - 1) - 1];
    terms[nt++] = prior + (seed_char == L->rflank[n2 - 1] ? sc : sw) + r_prob + Lm[(size_t)Lf * (hapsize - 1) - 1];
    for (int i = 1; i < hapsize - 1; ++i) {
      if (i >= n0 && i < n0 + n1) continue; /* repeat block positions cannot hold the seed */
      const char hc = i < n0 ? L->lflank[i] : L->rflank[i - n0 - n1];
      terms[nt++] = prior + (seed_char == hc ? sc : sw) + Lm[(size_t)Lf * i - 1] + Rm[(size_t)Rf * (hapsize - 1 - i) - 1];
    }
    out_row[a] = s_fast_lse(terms, nt);
    free(terms);
  }
  free(Lm); free(Rm); free(rev_l); free(rev_r); free(rseq); free(rlw); free(rlc); free(lw); free(lc);
  return rc;
}
