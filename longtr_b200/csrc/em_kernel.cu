// em_kernel.cu -- ltr_em_stutter_train: length-based EM of the per-locus stutter model.
//
// Reference: EMStutterGenotyper (src/em_stutter_genotyper.{h,cpp}) as GenotyperBamProcessor::learn_stutter_model drives it
// (src/genotyper_bam_processor.cpp:170-225: allele sizes = distinct observed bp differences with the reference allele 0 first,
// train(MAX_EM_ITER = 100, ABS_LL_CONVERGE = 0.01, FRAC_LL_CONVERGE = 0.001)).  In the reference's CLI this path is switched
// off by the default stutter model (hipstr_main.cpp:140, 362-363); the class itself is complete and is the oracle
// (oracle/em_driver.cpp compiles it in place).
//   train                               em_stutter_genotyper.cpp:170-226   loop, convergence rules
//   init_log_gt_priors                  :10-19     allele frequencies from read counts (+1 pseudo count)
//   calc_hap_aln_probs                  :140-144   LL[read][allele] = StutterModel::log_stutter_pmf (stutter_model.cpp:29-53)
//   Genotyper::calc_log_sample_posteriors  genotyper.cpp:45-83 with the priors of init_log_sample_priors (:128-138 here)
//   recalc_log_read_phase_posteriors    :146-163   approximate two-term log-sum-exp (fastlog / fastexp)
//   recalc_log_gt_priors                :21-55     streaming log-sum-exp per allele
//   recalc_stutter_model                :62-126    seven approximate log-sum-exps (fasterlog / fasterexp) + pseudo counts
//
// One warp per locus, the whole EM loop inside the kernel (no host round trips).  Every sum whose value depends on the order
// of its terms runs in the reference's order (per-sample posterior accumulation over reads, the exact log-sum-exps, the
// streaming ones per allele); the seven big sums of recalc_stutter_model add single-precision values within 2^10 of each
// other into a double -- exact for up to 2^19 terms, so they are reduced in parallel.  fastexp / fastlog / fasterexp /
// fasterlog are reproduced operation by operation with single-precision intrinsics; exp / log in double are CUDA's (<= 1 ulp
// from glibc's), so parameters agree with the reference to ~1e-12 and the iteration counts are the same unless a
// convergence test falls within that distance of its threshold (tests: 1e-9).
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <set>
#include <vector>

#include "ctx.h"
#include "kernels.h"

namespace ltr {

struct EmLocus {
  uint32_t read_begin, read_end;
  uint32_t allele_begin;    // into bps[]
  uint32_t sample_begin;    // into sample_read_begin[] (n_samples + 1 entries, read indices)
  int32_t n_alleles, n_samples, motif_len, haploid;
  unsigned long long scratch_off;  // doubles
};

struct EmArgs {
  const EmLocus* loci;
  uint32_t n_loci;
  const int32_t* aidx;        // [n_reads] allele index of the read's observed size
  const int32_t* sample;      // [n_reads] sample label
  const double* log_p1;
  const double* log_p2;
  const int32_t* bps;         // allele sizes (bp differences), reference allele first
  const uint32_t* sample_read_begin;
  const double* int_logs;     // host libm log(i), [0] = -1000 (mathops.cpp:14-20)
  double* scratch;
  double log_half, log_thresh, log_1p1, tolerance;
  int32_t max_iter;
  double abs_conv, frac_conv;
  double* out_params;         // [6 * n_loci]
  int32_t* out_trained;       // [n_loci]
  int32_t* out_n_iter;        // [n_loci]
  double* out_ll;             // [n_loci]
  double* out_gt_priors;      // [n_loci * prior_stride] or NULL
  uint32_t prior_stride;
};

namespace {

// ---- fastonebigheader.h:188-218, 320-357 in explicit single precision ------------------------------------------------------
__device__ __forceinline__ float em_fastpow2(float p) {
  const float offset = (p < 0.0f) ? 1.0f : 0.0f;
  const float clipp = (p < -126.0f) ? -126.0f : p;
  const int w = __float2int_rz(clipp);
  const float z = __fadd_rn(__fsub_rn(clipp, __int2float_rn(w)), offset);
  const float t = __fsub_rn(__fadd_rn(__fadd_rn(clipp, 121.2740575f), __fdiv_rn(27.7280233f, __fsub_rn(4.84252568f, z))),
                            __fmul_rn(1.49012907f, z));
  return __uint_as_float(__float2uint_rz(__fmul_rn(8388608.0f, t)));
}
__device__ __forceinline__ float em_fastexp(float p) { return em_fastpow2(__fmul_rn(1.442695040f, p)); }
__device__ __forceinline__ float em_fastlog(float x) {
  const uint32_t bits = __float_as_uint(x);
  const float mx = __uint_as_float((bits & 0x007FFFFFu) | 0x3f000000u);
  float y = __uint2float_rn(bits);
  y = __fmul_rn(y, 1.1920928955078125e-7f);
  const float l2 = __fsub_rn(__fsub_rn(__fsub_rn(y, 124.22551499f), __fmul_rn(1.498030302f, mx)),
                             __fdiv_rn(1.72587999f, __fadd_rn(0.3520887068f, mx)));
  return __fmul_rn(0.69314718f, l2);
}
__device__ __forceinline__ float em_fasterexp(float p) {
  const float x = __fmul_rn(1.442695040f, p);
  const float clipp = (x < -126.0f) ? -126.0f : x;
  return __uint_as_float(__float2uint_rz(__fmul_rn(8388608.0f, __fadd_rn(clipp, 126.94269504f))));
}
__device__ __forceinline__ float em_fasterlog(float x) {
  float y = __uint2float_rn(__float_as_uint(x));
  y = __fmul_rn(y, 8.2629582881927490e-8f);
  return __fsub_rn(y, 87.989971088f);
}
// mathops.cpp:87-96
__device__ __forceinline__ double em_fast_lse2(double v1, double v2, double log_thresh) {
  if (v1 > v2) {
    const double diff = v2 - v1;
    return diff < log_thresh ? v1 : v1 + (double)em_fastlog(__fadd_rn(1.0f, em_fastexp(__double2float_rn(diff))));
  }
  const double diff = v1 - v2;
  return diff < log_thresh ? v2 : v2 + (double)em_fastlog(__fadd_rn(1.0f, em_fastexp(__double2float_rn(diff))));
}
// mathops.cpp:53-58, 60-63
__device__ __forceinline__ double em_lse2(double a, double b) {
  return (a > b) ? a + log(1 + exp(b - a)) : b + log(1 + exp(a - b));
}
__device__ __forceinline__ double em_lse3(double a, double b, double c) {
  const double m = fmax(fmax(a, b), c);
  return m + log(exp(a - m) + exp(b - m) + exp(c - m));
}

struct Model {  // StutterModel (stutter_model.h:17-67)
  double in_geom, in_up, in_down, out_geom, out_up, out_down;
  double in_log_step, in_log_nostep, in_log_up, in_log_down, log_equal, out_log_step, out_log_nostep, out_log_up, out_log_down;
};
__device__ void model_set(Model& m, double ig, double iu, double id, double og, double ou, double od) {
  m.in_geom = ig; m.in_up = iu; m.in_down = id; m.out_geom = og; m.out_up = ou; m.out_down = od;
  m.in_log_step = log(1 - ig);
  m.in_log_nostep = log(ig);
  m.in_log_up = log(iu);
  m.in_log_down = log(id);
  m.out_log_step = log(1 - og);
  m.out_log_nostep = log(og);
  m.out_log_up = log(ou);
  m.out_log_down = log(od);
  m.log_equal = log(1 - iu - id - ou - od);
}
// stutter_model.cpp:29-53
__device__ __forceinline__ double model_pmf(const Model& m, int motif_len, int sample_bps, int read_bps) {
  const int bp_diff = read_bps - sample_bps;
  if (bp_diff % motif_len != 0) {
    const int eff = bp_diff - (bp_diff / motif_len);
    if (eff < 0) return m.out_log_down + m.out_log_nostep + m.out_log_step * (-eff - 1);
    return m.out_log_up + m.out_log_nostep + m.out_log_step * (eff - 1);
  }
  const int rep = bp_diff / motif_len;
  if (rep == 0) return m.log_equal;
  if (rep < 0) return m.in_log_down + m.in_log_nostep + m.in_log_step * (-rep - 1);
  return m.in_log_up + m.in_log_nostep + m.in_log_step * (rep - 1);
}

__device__ __forceinline__ double warp_max(double v) {
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {  // only for sums that are exact in double (see the file header)
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

// exact log_sum_exp over x[0..n) in index order (mathops.cpp:45-51); e[] is scratch.  Uniform result on all lanes.
__device__ double warp_lse_ordered(const double* x, double* e, int n, int lane) {
  double mx = -DBL_MAX;
  for (int k = lane; k < n; k += 32) mx = fmax(mx, x[k]);
  mx = warp_max(mx);
  for (int k = lane; k < n; k += 32) e[k] = exp(x[k] - mx);
  __syncwarp();
  double tot = 0.0;
  if (lane == 0)
    for (int k = 0; k < n; ++k) tot += e[k];
  tot = __shfl_sync(0xFFFFFFFFu, tot, 0);
  __syncwarp();
  return mx + log(tot);
}

enum { C_IN_EQ = 0, C_IN_UP, C_IN_DOWN, C_IN_DIFF, C_OUT_UP, C_OUT_DOWN, C_OUT_DIFF, C_N };

}  // namespace

__global__ void __launch_bounds__(128) em_stutter_kernel(EmArgs A) {
  const int lane = threadIdx.x & 31;
  const uint32_t l = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (l >= A.n_loci) return;
  const EmLocus L = A.loci[l];
  const int nA = L.n_alleles, nS = L.n_samples, AA = nA * nA;
  const int R = (int)(L.read_end - L.read_begin);
  const int32_t* aidx = A.aidx + L.read_begin;
  const int32_t* samp = A.sample + L.read_begin;
  const double* lp1 = A.log_p1 + L.read_begin;
  const double* lp2 = A.log_p2 + L.read_begin;
  const int32_t* bps = A.bps + L.allele_begin;
  const uint32_t* srb = A.sample_read_begin + L.sample_begin;  // absolute read indices
  double* T = A.scratch + L.scratch_off;   // [AA] log_stutter_pmf(allele a, allele of the read)
  double* post = T + AA;                   // [nS * AA]
  double* prior = post + (size_t)nS * AA;  // [nA]
  double* ebuf = prior + nA;               // [max(AA, nA)]
  double* tmp = ebuf + (AA > nA ? AA : nA);  // [nA]

  // ---- init_log_gt_priors (:10-19) ----
  for (int a = lane; a < nA; a += 32) {
    double x = 1.0;
    for (int r = 0; r < R; ++r)
      if (aidx[r] == a) {
        const int s = samp[r];
        x += 1.0 / (double)(int)(srb[s + 1] - srb[s]);
      }
    tmp[a] = x;
  }
  __syncwarp();
  {
    double tot = 0.0;
    if (lane == 0)
      for (int a = 0; a < nA; ++a) tot += tmp[a];
    tot = __shfl_sync(0xFFFFFFFFu, tot, 0);
    const double log_total = log(tot);
    for (int a = lane; a < nA; a += 32) prior[a] = log(tmp[a]) - log_total;
  }
  __syncwarp();
  Model M;
  model_set(M, 0.9, 0.1, 0.1, 0.8, 0.01, 0.01);  // init_stutter_model (:57-60)

  int num_iter = 1, n_done = 0;
  double LL = -DBL_MAX, new_LL = 0.0;
  int trained = 0;
  while (num_iter <= A.max_iter) {
    // ---- E step: calc_hap_aln_probs (:140-144) as a table over (allele, allele of the read) ----
    for (int k = lane; k < AA; k += 32) T[k] = model_pmf(M, L.motif_len, bps[k / nA], bps[k % nA]);
    __syncwarp();
    // ---- calc_log_sample_posteriors (genotyper.cpp:45-83), priors from init_log_sample_priors (:128-138) ----
    new_LL = 0.0;
    for (int s = 0; s < nS; ++s) {
      double* ps = post + (size_t)s * AA;
      const int r0 = (int)(srb[s] - L.read_begin), r1 = (int)(srb[s + 1] - L.read_begin);
      for (int k = lane; k < AA; k += 32) {
        const int i = k / nA, j = k % nA;
        double acc = L.haploid ? (i == j ? prior[i] : -DBL_MAX / 2) : prior[i] + prior[j];
        for (int r = r0; r < r1; ++r) {
          const int ar = aidx[r];
          double t1 = T[i * nA + ar], t2 = T[j * nA + ar];
          t1 = (t1 < -600) ? -600 : t1;
          t2 = (t2 < -600) ? -600 : t2;
          acc += log(exp(t1 + lp1[r] + A.log_half) + exp(t2 + lp2[r] + A.log_half));
        }
        ps[k] = acc;
      }
      __syncwarp();
      const double total = warp_lse_ordered(ps, ebuf, AA, lane);
      for (int k = lane; k < AA; k += 32) ps[k] -= total;
      new_LL += total;
      __syncwarp();
    }
    ++n_done;
    if (new_LL < LL + A.tolerance) {  // :196-200
      trained = 1;
      break;
    }
    // ---- M step: recalc_log_gt_priors (:21-55) ----
    for (int a = lane; a < nA; a += 32) {
      double mx = -DBL_MAX / 2, total = 0.0;
      for (int s = 0; s < nS; ++s) {  // first allele of the diplotype: exact log-sum-exp of row (s, a, .)
        const double* row = post + (size_t)s * AA + (size_t)a * nA;
        double rm = row[0];
        for (int j = 1; j < nA; ++j) rm = fmax(rm, row[j]);
        double rt = 0.0;
        for (int j = 0; j < nA; ++j) rt += exp(row[j] - rm);
        const double v = rm + log(rt);
        if (v <= mx) total += exp(v - mx);
        else {
          total *= exp(mx - v);
          total += 1.0;
          mx = v;
        }
      }
      for (int s = 0; s < nS; ++s)  // second allele
        for (int i = 0; i < nA; ++i) {
          const double v = post[(size_t)s * AA + (size_t)i * nA + a];
          if (v <= mx) total += exp(v - mx);
          else {
            total *= exp(mx - v);
            total += 1.0;
            mx = v;
          }
        }
      tmp[a] = mx + log(total);
    }
    __syncwarp();
    {
      const double log_total = warp_lse_ordered(tmp, ebuf, nA, lane);
      for (int a = lane; a < nA; a += 32) prior[a] = tmp[a] - log_total;
    }
    __syncwarp();
    // ---- recalc_stutter_model (:62-126): phase posteriors (:146-163) recomputed on the fly, maxima then sums ----
    double mx[C_N], sm[C_N];
#pragma unroll
    for (int c = 0; c < C_N; ++c) mx[c] = 0.0;  // every list starts with the pseudo count 0.0 ...
    mx[C_IN_DIFF] = A.log_1p1;                  // ... and the two "diffs" lists also hold log(1.1)
    mx[C_OUT_DIFF] = A.log_1p1;
    for (int pass = 0; pass < 2; ++pass) {
      if (pass == 1) {
#pragma unroll
        for (int c = 0; c < C_N; ++c) {
          mx[c] = warp_max(mx[c]);
          sm[c] = 0.0;
        }
      }
      const long long n_trip = (long long)R * AA;
      for (long long q = lane; q < n_trip; q += 32) {
        const int r = (int)(q / AA), k = (int)(q % AA);
        const int i = k / nA, j = k % nA;
        const int ar = aidx[r];
        const double gp = post[(size_t)samp[r] * AA + k];
        const double one = A.log_half + lp1[r] + T[i * nA + ar];
        const double two = A.log_half + lp2[r] + T[j * nA + ar];
        const double tot = em_fast_lse2(one, two, A.log_thresh);
#pragma unroll
        for (int ph = 0; ph < 2; ++ph) {
          const int gt = ph == 0 ? i : j;
          const int bp_diff = bps[ar] - bps[gt];
          const double factor = gp + ((ph == 0 ? one : two) - tot);
          int c_a, c_b = -1;
          double vb = 0.0;
          if (bp_diff == 0) {
            c_a = C_IN_EQ;
          } else if (bp_diff % L.motif_len != 0) {
            int eff = bp_diff - bp_diff / L.motif_len;
            eff = eff < 0 ? -eff : eff;
            c_b = C_OUT_DIFF;
            vb = factor + A.int_logs[eff];
            c_a = bp_diff > 0 ? C_OUT_UP : C_OUT_DOWN;
          } else {
            int eff = bp_diff / L.motif_len;
            eff = eff < 0 ? -eff : eff;
            c_b = C_IN_DIFF;
            vb = factor + A.int_logs[eff];
            c_a = bp_diff > 0 ? C_IN_UP : C_IN_DOWN;
          }
#pragma unroll
          for (int c = 0; c < C_N; ++c) {
            if (pass == 0) {
              if (c == c_a) mx[c] = fmax(mx[c], factor);
              if (c == c_b) mx[c] = fmax(mx[c], vb);
            } else {
              if (c == c_a) {
                const double d = factor - mx[c];
                if (d > A.log_thresh) sm[c] += (double)em_fasterexp(__double2float_rn(d));
              }
              if (c == c_b) {
                const double d = vb - mx[c];
                if (d > A.log_thresh) sm[c] += (double)em_fasterexp(__double2float_rn(d));
              }
            }
          }
        }
      }
    }
    double tot7[C_N];
#pragma unroll
    for (int c = 0; c < C_N; ++c) {
      double s = warp_sum(sm[c]);
      // the pseudo counts (the same terms on every lane: added once)
      double d = 0.0 - mx[c];
      if (d > A.log_thresh) s += (double)em_fasterexp(__double2float_rn(d));
      if (c == C_IN_DIFF || c == C_OUT_DIFF) {
        d = A.log_1p1 - mx[c];
        if (d > A.log_thresh) s += (double)em_fasterexp(__double2float_rn(d));
      }
      tot7[c] = mx[c] + (double)em_fasterlog(__double2float_rn(s));
    }
    const double out_log_total = em_fast_lse2(tot7[C_OUT_UP], tot7[C_OUT_DOWN], A.log_thresh);
    const double in_pgeom = fmin(0.999, exp(em_lse2(tot7[C_IN_UP], tot7[C_IN_DOWN]) - tot7[C_IN_DIFF]));
    const double out_pgeom = fmin(0.999, exp(out_log_total - tot7[C_OUT_DIFF]));
    const double log_total = em_lse2(em_lse3(tot7[C_IN_UP], tot7[C_IN_DOWN], tot7[C_IN_EQ]), out_log_total);
    const double in_pup = exp(tot7[C_IN_UP] - log_total), in_pdown = exp(tot7[C_IN_DOWN] - log_total);
    const double out_pup = exp(tot7[C_OUT_UP] - log_total), out_pdown = exp(tot7[C_OUT_DOWN] - log_total);
    const Model prev = M;
    model_set(M, in_pgeom, in_pup, in_pdown, out_pgeom, out_pup, out_pdown);
    // ---- convergence (:210-222) ----
    const double abs_change = new_LL - LL;
    const double frac_change = -(new_LL - LL) / LL;
    bool converged = false;
    if (abs_change < A.abs_conv && frac_change < A.frac_conv) converged = true;
    else {
      const double md = 0.0001;
      converged = fabs(prev.in_geom - M.in_geom) < md && fabs(prev.in_up - M.in_up) < md && fabs(prev.in_down - M.in_down) < md &&
                  fabs(prev.out_geom - M.out_geom) < md && fabs(prev.out_up - M.out_up) < md && fabs(prev.out_down - M.out_down) < md;
    }
    if (converged) {
      trained = 1;
      break;
    }
    LL = new_LL;
    ++num_iter;
  }
  if (lane == 0) {
    double* p = A.out_params + (size_t)l * 6;
    p[0] = M.in_geom; p[1] = M.in_up; p[2] = M.in_down; p[3] = M.out_geom; p[4] = M.out_up; p[5] = M.out_down;
    A.out_trained[l] = trained;
    A.out_n_iter[l] = n_done;
    A.out_ll[l] = new_LL;
  }
  if (A.out_gt_priors)
    for (int a = lane; a < nA && a < (int)A.prior_stride; a += 32) A.out_gt_priors[(size_t)l * A.prior_stride + a] = prior[a];
}

}  // namespace ltr

using namespace ltr;

namespace {
struct Pool {
  std::vector<DeviceBuffer*> all;
  ~Pool() {
    for (DeviceBuffer* b : all) {
      b->free();
      delete b;
    }
  }
  DeviceBuffer* get() {
    all.push_back(new DeviceBuffer());
    return all.back();
  }
};
template <typename T>
int up(ltr_ctx* ctx, Pool& pool, const std::vector<T>& v, T** dev) {
  DeviceBuffer* b = pool.get();
  LTR_CUDA(ctx, b->alloc(v.size() * sizeof(T) + 16));
  if (!v.empty()) LTR_CUDA(ctx, cudaMemcpyAsync(b->p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->main_stream));
  *dev = b->as<T>();
  return LTR_OK;
}
template <typename T>
int room(ltr_ctx* ctx, Pool& pool, size_t n, T** dev) {
  DeviceBuffer* b = pool.get();
  LTR_CUDA(ctx, b->alloc(n * sizeof(T) + 16));
  *dev = b->as<T>();
  return LTR_OK;
}
}  // namespace

extern "C" void ltr_em_opts_default(ltr_em_opts* o) {
  if (!o) return;
  o->max_iter = 100;         // MAX_EM_ITER       (genotyper_bam_processor.h:107-109)
  o->abs_ll_converge = 0.01; // ABS_LL_CONVERGE
  o->frac_ll_converge = 0.001;  // FRAC_LL_CONVERGE
}

extern "C" int ltr_em_stutter_train(ltr_ctx* ctx, const ltr_em_batch* B, const ltr_em_opts* opts, double* out_params,
                                    int32_t* out_trained, int32_t* out_n_iter, double* out_ll, double* out_log_gt_priors,
                                    uint32_t prior_stride) {
  if (!ctx || !B || !opts || !out_params || !out_trained) return LTR_ERR_INVALID;
  const uint32_t n = B->n_loci;
  if (n == 0) return LTR_OK;
  if (!B->locus_sample_begin || !B->sample_read_begin || !B->locus_motif_len || opts->max_iter < 1) return LTR_ERR_INVALID;
  if (out_log_gt_priors && prior_stride == 0) return LTR_ERR_INVALID;
  const uint32_t n_samples = B->locus_sample_begin[n];
  for (uint32_t l = 0; l < n; ++l)
    if (B->locus_sample_begin[l + 1] < B->locus_sample_begin[l] || B->locus_motif_len[l] < 1) return LTR_ERR_INVALID;
  for (uint32_t s = 0; s < n_samples; ++s)
    if (B->sample_read_begin[s + 1] < B->sample_read_begin[s]) return LTR_ERR_INVALID;
  const uint32_t n_reads = B->sample_read_begin[n_samples];
  if (n_reads && (!B->read_bp_diff || !B->log_p1 || !B->log_p2)) return LTR_ERR_INVALID;
  for (uint32_t r = 0; r < n_reads; ++r)
    if (B->log_p1[r] > 0.0 || B->log_p2[r] > 0.0) return LTR_ERR_INVALID;  // em_stutter_genotyper.h:96
  // allele sizes per locus (em_stutter_genotyper.h:61-83): distinct observed sizes, ascending, the reference allele (0) first
  std::vector<EmLocus> loci(n);
  std::vector<int32_t> bps, aidx(n_reads), sample(n_reads);
  std::vector<uint32_t> srb;
  unsigned long long scratch = 0;
  int32_t max_abs = 1;
  for (uint32_t l = 0; l < n; ++l) {
    const uint32_t s0 = B->locus_sample_begin[l], s1 = B->locus_sample_begin[l + 1];
    const uint32_t r0 = B->sample_read_begin[s0], r1 = B->sample_read_begin[s1];
    std::set<int32_t> sizes(B->read_bp_diff + r0, B->read_bp_diff + r1);
    sizes.erase(0);
    EmLocus& E = loci[l];
    E.read_begin = r0;
    E.read_end = r1;
    E.allele_begin = (uint32_t)bps.size();
    E.sample_begin = (uint32_t)srb.size();
    E.n_samples = (int32_t)(s1 - s0);
    E.motif_len = B->locus_motif_len[l];
    E.haploid = (B->locus_haploid && B->locus_haploid[l]) ? 1 : 0;
    std::map<int32_t, int32_t> index;
    bps.push_back(0);
    index[0] = 0;
    for (int32_t v : sizes) {
      index[v] = (int32_t)(bps.size() - E.allele_begin);
      bps.push_back(v);
    }
    E.n_alleles = (int32_t)(bps.size() - E.allele_begin);
    for (uint32_t s = s0; s <= s1; ++s) srb.push_back(B->sample_read_begin[s]);
    for (uint32_t s = s0; s < s1; ++s)
      for (uint32_t r = B->sample_read_begin[s]; r < B->sample_read_begin[s + 1]; ++r) {
        aidx[r] = index[B->read_bp_diff[r]];
        sample[r] = (int32_t)(s - s0);
      }
    const long long lo = sizes.empty() ? 0 : *sizes.begin(), hi = sizes.empty() ? 0 : *sizes.rbegin();
    const long long span = std::max<long long>(std::max<long long>(hi, 0) - std::min<long long>(lo, 0), 1);
    if (span > 999999) return LTR_ERR_INVALID;  // INT_LOGS has a million entries
    max_abs = std::max<int32_t>(max_abs, (int32_t)span);
    const unsigned long long AA = (unsigned long long)E.n_alleles * E.n_alleles;
    E.scratch_off = scratch;
    scratch += AA + (unsigned long long)E.n_samples * AA + E.n_alleles + std::max<unsigned long long>(AA, E.n_alleles) + E.n_alleles;
  }
  std::vector<double> int_logs((size_t)max_abs + 2);
  int_logs[0] = -1000;
  for (size_t i = 1; i < int_logs.size(); ++i) int_logs[i] = log((double)i);
  LTR_CUDA(ctx, cudaSetDevice(ctx->device));
  AllocScope alloc_scope(ctx->main_stream);
  Pool pool;
  EmArgs A;
  memset(&A, 0, sizeof(A));
  EmLocus* d_loci;
  int32_t *d_aidx, *d_sample, *d_bps;
  uint32_t* d_srb;
  double *d_p1, *d_p2, *d_logs;
  int rc;
  if ((rc = up(ctx, pool, loci, &d_loci)) != LTR_OK) return rc;
  if ((rc = up(ctx, pool, aidx, &d_aidx)) != LTR_OK) return rc;
  if ((rc = up(ctx, pool, sample, &d_sample)) != LTR_OK) return rc;
  if ((rc = up(ctx, pool, bps, &d_bps)) != LTR_OK) return rc;
  if ((rc = up(ctx, pool, srb, &d_srb)) != LTR_OK) return rc;
  if ((rc = up(ctx, pool, int_logs, &d_logs)) != LTR_OK) return rc;
  std::vector<double> p1(B->log_p1, B->log_p1 + n_reads), p2(B->log_p2, B->log_p2 + n_reads);
  if ((rc = up(ctx, pool, p1, &d_p1)) != LTR_OK) return rc;
  if ((rc = up(ctx, pool, p2, &d_p2)) != LTR_OK) return rc;
  if ((rc = room(ctx, pool, (size_t)scratch, &A.scratch)) != LTR_OK) return rc;
  if ((rc = room(ctx, pool, (size_t)n * 6, &A.out_params)) != LTR_OK) return rc;
  if ((rc = room(ctx, pool, (size_t)n, &A.out_trained)) != LTR_OK) return rc;
  if ((rc = room(ctx, pool, (size_t)n, &A.out_n_iter)) != LTR_OK) return rc;
  if ((rc = room(ctx, pool, (size_t)n, &A.out_ll)) != LTR_OK) return rc;
  if (out_log_gt_priors) {
    if ((rc = room(ctx, pool, (size_t)n * prior_stride, &A.out_gt_priors)) != LTR_OK) return rc;
    LTR_CUDA(ctx, cudaMemsetAsync(A.out_gt_priors, 0, (size_t)n * prior_stride * sizeof(double), ctx->main_stream));
  }
  A.loci = d_loci; A.n_loci = n; A.aidx = d_aidx; A.sample = d_sample; A.log_p1 = d_p1; A.log_p2 = d_p2; A.bps = d_bps;
  A.sample_read_begin = d_srb; A.int_logs = d_logs; A.prior_stride = prior_stride;
  A.log_half = log(0.5);       // LOG_ONE_HALF, host libm like the reference's (mathops.cpp:10)
  A.log_thresh = log(0.001);   // LOG_THRESH (mathops.h:36)
  A.log_1p1 = log(1.1);        // pseudo count of recalc_stutter_model (:67-68)
  A.tolerance = 1e-10;         // TOLERANCE (mathops.cpp:11)
  A.max_iter = opts->max_iter; A.abs_conv = opts->abs_ll_converge; A.frac_conv = opts->frac_ll_converge;
  const int warps = 4;
  em_stutter_kernel<<<(n + warps - 1) / warps, warps * 32, 0, ctx->main_stream>>>(A);
  LTR_CUDA(ctx, cudaGetLastError());
  LTR_CUDA(ctx, cudaMemcpyAsync(out_params, A.out_params, (size_t)n * 6 * sizeof(double), cudaMemcpyDeviceToHost, ctx->main_stream));
  LTR_CUDA(ctx, cudaMemcpyAsync(out_trained, A.out_trained, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->main_stream));
  if (out_n_iter)
    LTR_CUDA(ctx, cudaMemcpyAsync(out_n_iter, A.out_n_iter, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->main_stream));
  if (out_ll) LTR_CUDA(ctx, cudaMemcpyAsync(out_ll, A.out_ll, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->main_stream));
  if (out_log_gt_priors)
    LTR_CUDA(ctx, cudaMemcpyAsync(out_log_gt_priors, A.out_gt_priors, (size_t)n * prior_stride * sizeof(double),
                                  cudaMemcpyDeviceToHost, ctx->main_stream));
  LTR_CUDA(ctx, cudaStreamSynchronize(ctx->main_stream));
  return LTR_OK;
}
