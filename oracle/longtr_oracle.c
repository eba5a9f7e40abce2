/* TEST INFRASTRUCTURE ONLY -- see longtr_oracle.h.
 *
 * Plain-C restatement of LongTR's read x haplotype hot path, written from the
 * behaviour of the reference (file:line cited per function), not from its text.
 * Arithmetic notes that matter for bit parity:
 *   - accumulators are double, every model constant is a *float* that is
 *     promoted to double at each use (HapAligner.h:16-22, HapAligner.cpp:260-261);
 *   - MATCH + LOG_MATCH_TO_INS is added in float first (HapAligner.cpp:277);
 *   - the band penalty is int*float -> float, then added to a double (:298);
 *   - build with -ffp-contract=off (the reference's x86-64 build has no FMA).
 */
#include "longtr_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define ORC_IMPOSSIBLE (-1000000000.0) /* HapAligner.cpp:20 */
#define ORC_REF_FLANK 35               /* HapAligner.cpp:245 (HaplotypeGenerator REF_FLANK_LEN) */

static inline double dmax(double a, double b) { return (a < b) ? b : a; } /* std::max */

void ltr_oracle_default_params(ltr_oracle_params* p) {
  /* HapAligner.h:118 */
  p->ins_ins = -1.0f;
  p->ins_match = (float)-0.458675;
  p->del_del = -1.0f;
  p->del_match = (float)-0.458675;
  p->match_match = (float)-0.00005800168;
  p->match_ins = (float)-10.448214728;
  p->match_del = (float)-10.448214728;
  p->indel_flank_len = 5;
}

/* ---------------------------------------------------------------------------
 * HapAligner::align_seq_to_hap  (HapAligner.cpp:236-343)
 * Rolling two-row evaluation of the same cells in the same order.
 * ------------------------------------------------------------------------- */
double ltr_oracle_viterbi_pair_cells(const char* full_hap, int32_t hap_len, const char* read,
                                     int32_t read_len, const ltr_oracle_params* p,
                                     int64_t* cells) {
  const float MISMATCH = -9.0f;               /* :260 */
  const float MATCH = (float)-0.000100005;    /* :261 */
  if (cells) *cells = 0;
  if (hap_len <= 60) return ORC_IMPOSSIBLE;   /* :241-244 */
  const int cut = ORC_REF_FLANK - p->indel_flank_len; /* :246 */
  const char* h = full_hap + cut;
  const int n = hap_len - 2 * cut;
  /* std::string read_seq = seq_0 stops at the first NUL (:240) */
  int m = 0;
  while (m < read_len && read[m] != '\0') m++;
  if (abs(n - m) > 600) return -700.0;        /* :249-252 */
  if (n <= 0 || m <= 0) return ORC_IMPOSSIBLE; /* not reachable from LongTR (n>=1 when hap_len>60 and flank 5; m>=10) */

  /* out-of-range reads the reference performs: h[j] for j>=n (row 0) -> '\0'
   * policy; r[1] when m==1 is the std::string terminator.       */
#define HCH(k) ((k) < n ? h[(k)] : '\0')
#define RCH(k) ((k) < m ? read[(k)] : '\0')

  double* buf = (double*)malloc(sizeof(double) * 6 * (size_t)m);
  double *Mp = buf, *Ip = buf + m, *Dp = buf + 2 * m;
  double *Mc = buf + 3 * m, *Ic = buf + 4 * m, *Dc = buf + 5 * m;

  /* row 0 (:263-272) */
  Dp[0] = ORC_IMPOSSIBLE;
  Ip[0] = ORC_IMPOSSIBLE;
  Mp[0] = (HCH(0) == RCH(0)) ? MATCH : MISMATCH;
  double left = 0.0;
  for (int j = 1; j < m; ++j) {
    Mp[j] = Dp[j - 1] + p->del_match + ((HCH(j) == RCH(0)) ? MATCH : MISMATCH);
    Ip[j] = ORC_IMPOSSIBLE;
    Dp[j] = p->match_del + left;
    left += p->del_del;
  }
  /* column 0 is generated on the fly (:274-280) */
  left = 0.0;
  const float match_plus_m2i = MATCH + p->match_ins; /* float add, :277 */
  const double e_col0 = (HCH(0) == RCH(1)) ? MATCH : MISMATCH;
  const int nm = n - m;
  int64_t ncell = 0;
  for (int i = 1; i < n; ++i) {
    Mc[0] = Ip[0] + p->ins_match + e_col0;
    Ic[0] = match_plus_m2i + left;
    Dc[0] = ORC_IMPOSSIBLE;
    left += p->ins_ins;
    double rowmax = ORC_IMPOSSIBLE;
    const char hc = h[i];
    for (int j = 1; j < m; ++j) {
      const double emit = (hc == read[j]) ? MATCH : MISMATCH;
      const double mv = emit + dmax(Mp[j - 1] + p->match_match,
                                    dmax(Dp[j - 1] + p->del_match, Ip[j - 1] + p->ins_match));
      const double iv = MATCH + dmax(Mp[j] + p->match_ins, Ip[j] + p->ins_ins);
      const double dv = dmax(Mc[j - 1] + p->match_del, Dc[j - 1] + p->del_del);
      Mc[j] = mv;
      Ic[j] = iv;
      Dc[j] = dv;
      const double best = dmax(dv, dmax(iv, mv));
      const float pen = (float)abs(nm - (i - j)) * p->del_del; /* int*float, :298 */
      const double v = best + pen;
      if (v > rowmax) rowmax = v;
    }
    ncell += (m - 1);
    if (rowmax < -600) { /* :300-306 */
      free(buf);
      if (cells) *cells = ncell;
      return -700.0;
    }
    double* t;
    t = Mp; Mp = Mc; Mc = t;
    t = Ip; Ip = Ic; Ic = t;
    t = Dp; Dp = Dc; Dc = t;
  }
  const double res = dmax(Dp[m - 1], dmax(Ip[m - 1], Mp[m - 1])); /* :309 */
  free(buf);
  if (cells) *cells = ncell;
  return res;
#undef HCH
#undef RCH
}

double ltr_oracle_viterbi_pair(const char* full_hap, int32_t hap_len, const char* read,
                               int32_t read_len, const ltr_oracle_params* p) {
  return ltr_oracle_viterbi_pair_cells(full_hap, hap_len, read, read_len, p, NULL);
}

/* ---------------------------------------------------------------------------
 * HapAligner::trim_alignment (HapAligner.cpp:346-465): keep the read bases that
 * the pooled read's CIGAR aligns to [repeat_start-pad, repeat_end+pad).
 * The CIGAR is consumed one base-unit at a time from either end.
 * ------------------------------------------------------------------------- */
typedef struct { char op; int32_t len; } orc_cig;

static int orc_parse_cigar(const char* s, orc_cig** out) {
  int cap = 16, cnt = 0, num = 0;
  orc_cig* v = (orc_cig*)malloc(sizeof(orc_cig) * cap);
  for (; *s; ++s) {
    if (*s >= '0' && *s <= '9') num = num * 10 + (*s - '0');
    else {
      if (cnt == cap) { cap *= 2; v = (orc_cig*)realloc(v, sizeof(orc_cig) * cap); }
      v[cnt].op = *s; v[cnt].len = num; cnt++; num = 0;
    }
  }
  *out = v;
  return cnt;
}

int32_t ltr_oracle_trim_read(const ltr_flat_locus* L, int32_t read_index, char* out) {
  const ltr_flat_read* R = &L->reads[read_index];
  const int32_t pad = L->indel_flank_len;
  const int32_t lo = L->repeat_start - pad, hi = L->repeat_end + pad;
  const int32_t seq_len = (int32_t)strlen(R->seq);
  int32_t start_pos = R->start + 1, end_pos = R->stop + 1;
  int32_t ltrim = 0, rtrim = 0;
  orc_cig* cg;
  int ncg = orc_parse_cigar(R->cigar, &cg);
  int f = 0, b = ncg - 1; /* live window [f, b] of CIGAR elements */
#define LIVE (f <= b)
#define POP_FRONT do { if (cg[f].len == 1) f++; else cg[f].len--; } while (0)
#define POP_BACK do { if (cg[b].len == 1) b--; else cg[b].len--; } while (0)
  /* 1. bases left of the window (:360-383) */
  while (start_pos <= lo && LIVE) {
    switch (cg[f].op) {
      case 'M': case '=': case 'X': ltrim++; start_pos++; break;
      case 'D': start_pos++; break;
      case 'I': case 'S': ltrim++; break;
      case 'H': break;
      default: free(cg); return -1;
    }
    POP_FRONT;
  }
  /* 2. inside the left pad: deletions pull one upstream base back in (:385-410) */
  int32_t mid = start_pos;
  while (mid > lo && mid <= lo + pad && LIVE) {
    switch (cg[f].op) {
      case 'M': case '=': case 'X': mid++; break;
      case 'D': ltrim--; mid++; break;
      case 'I': case 'S': case 'H': break;
      default: free(cg); return -1;
    }
    POP_FRONT;
  }
  /* 3. bases right of the window (:412-435) */
  while (end_pos > hi && LIVE) {
    switch (cg[b].op) {
      case 'M': case '=': case 'X': rtrim++; end_pos--; break;
      case 'D': end_pos--; break;
      case 'I': case 'S': rtrim++; break;
      case 'H': break;
      default: free(cg); return -1;
    }
    POP_BACK;
  }
  /* 4. inside the right pad (:437-460) */
  mid = end_pos;
  while (mid > hi - pad && mid <= hi && LIVE) {
    switch (cg[b].op) {
      case 'M': case '=': case 'X': mid--; break;
      case 'D': rtrim--; mid--; break;
      case 'I': case 'S': case 'H': break;
      default: free(cg); return -1;
    }
    POP_BACK;
  }
#undef LIVE
#undef POP_FRONT
#undef POP_BACK
  free(cg);
  if (ltrim < 0) ltrim = 0;
  if (rtrim < 0) rtrim = 0;
  int32_t keep = seq_len - ltrim - rtrim;
  if (keep < 0) return -1; /* reference asserts (:463) */
  memcpy(out, R->seq + ltrim, (size_t)keep);
  if (keep == 0) {
    /* process_read fallback (:820-823): last 5 of left flank + first 5 of right flank */
    const int32_t ll = (int32_t)strlen(L->lflank);
    memcpy(out, L->lflank + (ll - 5), 5);
    memcpy(out + 5, L->rflank, 5);
    keep = 10;
  }
  out[keep] = '\0';
  return keep;
}

/* ---------------------------------------------------------------------------
 * HapAligner::process_reads / process_read, long path
 * (HapAligner.cpp:545-581, 812-854).  With the three-block layout only the
 * repeat block has options, so the gray-code column order (Haplotype.cpp:157-196)
 * is simply the allele index.
 * ------------------------------------------------------------------------- */
int ltr_oracle_process_reads(const ltr_flat_locus* L, double* out_ll, int32_t* out_seeds) {
  if (L->period == 1 && L->switch_old_align_len != 0) { /* HapAligner.cpp:552, 567-579: homopolymer path */
    for (int r = 0; r < L->n_reads; ++r) {
      if (L->realign_read && !L->realign_read[r]) continue;
      const int32_t seed = ltr_oracle_seed_base(L, r);
      if (seed == -2) return -1;
      out_seeds[r] = seed;
      if (seed == -1) { /* no seed: every haplotype gets LL 0 (:570-574) */
        for (int a = 0; a < L->n_alleles; ++a) out_ll[(size_t)r * L->n_alleles + a] = 0;
        continue;
      }
      int rc = ltr_oracle_process_read_short(L, r, seed, out_ll + (size_t)r * L->n_alleles);
      if (rc != 0) return rc;
    }
    return 0;
  }
  ltr_oracle_params p;
  ltr_oracle_default_params(&p);
  if (L->n_aln_params == 7) {
    p.ins_ins = L->aln_params[0]; p.ins_match = L->aln_params[1];
    p.del_del = L->aln_params[2]; p.del_match = L->aln_params[3];
    p.match_match = L->aln_params[4]; p.match_ins = L->aln_params[5];
    p.match_del = L->aln_params[6];
  }
  p.indel_flank_len = L->indel_flank_len;
  const size_t ll = strlen(L->lflank), rl = strlen(L->rflank);
  const int H = L->n_alleles;
  char** haps = (char**)malloc(sizeof(char*) * H);
  int32_t* hlen = (int32_t*)malloc(sizeof(int32_t) * H);
  for (int a = 0; a < H; ++a) {
    const size_t al = strlen(L->alleles[a]);
    haps[a] = (char*)malloc(ll + al + rl + 1);
    memcpy(haps[a], L->lflank, ll);
    memcpy(haps[a] + ll, L->alleles[a], al);
    memcpy(haps[a] + ll + al, L->rflank, rl + 1);
    hlen[a] = (int32_t)(ll + al + rl);
  }
  int rc = 0;
  for (int r = 0; r < L->n_reads && rc == 0; ++r) {
    if (L->realign_read && !L->realign_read[r]) continue;
    const int32_t sl = (int32_t)strlen(L->reads[r].seq);
    out_seeds[r] = sl - 1; /* :562-563 */
    char* trimmed = (char*)malloc((size_t)sl + 11);
    int32_t tl = ltr_oracle_trim_read(L, r, trimmed);
    if (tl < 0) rc = -1;
    for (int a = 0; a < H && rc == 0; ++a) {
      if (L->realign_to_hap && !L->realign_to_hap[a]) continue; /* :841-845 */
      out_ll[(size_t)r * H + a] = ltr_oracle_viterbi_pair(haps[a], hlen[a], trimmed, tl, &p);
    }
    free(trimmed);
  }
  for (int a = 0; a < H; ++a) free(haps[a]);
  free(haps);
  free(hlen);
  return rc;
}

/* --------------------------------------------------------------------------- */
typedef struct {
  uint32_t l0, l1;
  const uint32_t *lhb, *lrb;
  const uint32_t *hoff, *roff;
  const uint8_t *hb, *rb;
  const ltr_oracle_params* p;
  const uint64_t* ll_off;
  double* out;
  int64_t cells;
} orc_job;

static void* orc_worker(void* arg) {
  orc_job* J = (orc_job*)arg;
  int64_t total = 0;
  for (uint32_t l = J->l0; l < J->l1; ++l) {
    const uint32_t h0 = J->lhb[l], h1 = J->lhb[l + 1], r0 = J->lrb[l], r1 = J->lrb[l + 1];
    const uint32_t H = h1 - h0;
    double* o = J->out + J->ll_off[l];
    for (uint32_t r = r0; r < r1; ++r)
      for (uint32_t h = h0; h < h1; ++h) {
        int64_t c;
        o[(size_t)(r - r0) * H + (h - h0)] = ltr_oracle_viterbi_pair_cells(
            (const char*)J->hb + J->hoff[h], (int32_t)(J->hoff[h + 1] - J->hoff[h]),
            (const char*)J->rb + J->roff[r], (int32_t)(J->roff[r + 1] - J->roff[r]), J->p, &c);
        total += c;
      }
  }
  J->cells = total;
  return NULL;
}

int ltr_oracle_viterbi_batch(uint32_t n_loci, const uint32_t* locus_hap_begin,
                             const uint32_t* locus_read_begin, const uint32_t* hap_off,
                             const uint8_t* hap_bytes, const uint32_t* read_off,
                             const uint8_t* read_bytes, const ltr_oracle_params* p,
                             double* out_ll, int64_t* cells, int n_threads) {
  uint64_t* ll_off = (uint64_t*)malloc(sizeof(uint64_t) * ((size_t)n_loci + 1));
  ll_off[0] = 0;
  for (uint32_t l = 0; l < n_loci; ++l)
    ll_off[l + 1] = ll_off[l] + (uint64_t)(locus_hap_begin[l + 1] - locus_hap_begin[l]) *
                                    (locus_read_begin[l + 1] - locus_read_begin[l]);
  if (n_threads < 1) n_threads = 1;
  if ((uint32_t)n_threads > n_loci && n_loci > 0) n_threads = (int)n_loci;
  orc_job* jobs = (orc_job*)calloc((size_t)n_threads, sizeof(orc_job));
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
  /* contiguous locus shards with (roughly) equal output counts */
  uint32_t l = 0;
  for (int t = 0; t < n_threads; ++t) {
    const uint64_t target = ll_off[n_loci] * (uint64_t)(t + 1) / (uint64_t)n_threads;
    uint32_t e = l;
    while (e < n_loci && (ll_off[e + 1] <= target || t == n_threads - 1)) e++;
    if (t == n_threads - 1) e = n_loci;
    jobs[t].l0 = l; jobs[t].l1 = e;
    jobs[t].lhb = locus_hap_begin; jobs[t].lrb = locus_read_begin;
    jobs[t].hoff = hap_off; jobs[t].roff = read_off;
    jobs[t].hb = hap_bytes; jobs[t].rb = read_bytes;
    jobs[t].p = p; jobs[t].ll_off = ll_off; jobs[t].out = out_ll;
    l = e;
  }
  for (int t = 1; t < n_threads; ++t) pthread_create(&th[t], NULL, orc_worker, &jobs[t]);
  orc_worker(&jobs[0]);
  int64_t total = jobs[0].cells;
  for (int t = 1; t < n_threads; ++t) { pthread_join(th[t], NULL); total += jobs[t].cells; }
  if (cells) *cells = total;
  free(jobs); free(th); free(ll_off);
  return 0;
}

/* ---------------------------------------------------------------------------
 * Genotyper::calc_log_sample_posteriors (genotyper.cpp:45-83)
 * ------------------------------------------------------------------------- */
double ltr_oracle_log_sample_posteriors(int haploid, int32_t n_samples, int32_t n_reads,
                                        int32_t n_alleles, double* ll, const double* log_p1,
                                        const double* log_p2, const int32_t* sample_label,
                                        double* post, double* totals) {
  const int H = n_alleles, HH = H * H;
  const double LOG_ONE_HALF = log(0.5); /* mathops.cpp:10 */
  /* priors (genotyper.cpp:21-43); int_log(k) = log(k) (mathops.cpp:16-22) */
  double hom, het;
  if (haploid) { hom = -log((double)H); het = -1.7976931348623157e308 / 2; }
  else {
    hom = log(2.0) - log((double)H) - log((double)(H + 1));
    het = -log((double)H) - log((double)(H + 1));
  }
  for (int s = 0; s < n_samples; ++s)
    for (int a = 0; a < H; ++a)
      for (int b = 0; b < H; ++b) post[(size_t)s * HH + a * H + b] = (a == b) ? hom : het;
  for (int r = 0; r < n_reads; ++r) {
    double* row = ll + (size_t)r * H;
    double* dst = post + (size_t)HH * sample_label[r];
    for (int a = 0; a < H; ++a)
      for (int b = 0; b < H; ++b) {
        if (row[a] < -600) row[a] = -600; /* in-place clamp, :57-58 */
        if (row[b] < -600) row[b] = -600;
        dst[a * H + b] += log(exp(row[a] + log_p1[r] + LOG_ONE_HALF) +
                              exp(row[b] + log_p2[r] + LOG_ONE_HALF));
      }
  }
  double total = 0.0;
  for (int s = 0; s < n_samples; ++s) {
    double* v = post + (size_t)s * HH;
    /* log_sum_exp(begin,end), mathops.cpp:45-51 */
    double mx = v[0];
    for (int k = 1; k < HH; ++k) if (mx < v[k]) mx = v[k];
    double acc = 0.0;
    for (int k = 0; k < HH; ++k) acc += exp(v[k] - mx);
    const double tot = mx + log(acc);
    totals[s] = tot;
    for (int k = 0; k < HH; ++k) v[k] -= tot;
    total += tot; /* sum(), mathops.cpp:24-29 */
  }
  return total;
}

/* Genotyper::get_optimal_haplotypes (genotyper.cpp:85-100): first strict maximum */
void ltr_oracle_optimal_haplotypes(int32_t n_samples, int32_t n_alleles, const double* post,
                                   int32_t* best) {
  const int H = n_alleles;
  for (int s = 0; s < n_samples; ++s) {
    double bv = -1.7976931348623157e308;
    best[2 * s] = best[2 * s + 1] = -1;
    for (int a = 0; a < H; ++a)
      for (int b = 0; b < H; ++b) {
        const double v = post[(size_t)s * H * H + a * H + b];
        if (v > bv) { bv = v; best[2 * s] = a; best[2 * s + 1] = b; }
      }
  }
}
