// posterior_kernel.cu -- genotype posteriors from the read x haplotype LL matrix.
//
// Restates Genotyper::calc_log_sample_posteriors (reference src/genotyper.cpp:45-83):
//   post[s][a][b] = prior(a==b) + sum_{reads r of sample s, in storage order}
//                   log( exp(LL[r][a] + log_p1[r] + log(1/2)) + exp(LL[r][b] + log_p2[r] + log(1/2)) )
// with LL clamped to >= -600 (:57-58), then per sample total = log_sum_exp over the H*H
// entries in storage order (mathops.cpp:45-51) and post -= total.
// One warp per locus; each lane owns whole (s,a,b) entries so every sum runs in the
// reference's order.  Priors use the host's libm log table (genotyper.cpp:21-33 use INT_LOGS).
// Floating point: CUDA exp/log are within 1 ulp of libm's -> parity tolerance 1e-12 relative.
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "kernels.h"

namespace ltr {

__global__ void __launch_bounds__(128) posterior_validate_kernel(const DevPosterior P, uint32_t* err) {
  bool bad = false;
  for (uint32_t l = blockIdx.x; l < P.n_loci; l += gridDim.x) {
    const uint32_t S = P.locus_n_samples[l];
    const uint32_t np = P.locus_read_begin[l + 1] - P.locus_read_begin[l];
    for (uint32_t r = P.locus_sread_begin[l] + threadIdx.x; r < P.locus_sread_begin[l + 1]; r += blockDim.x)
      bad |= (P.pool_index[r] >= np) || (P.sample_label[r] < 0) || ((uint32_t)P.sample_label[r] >= S);
  }
  if (bad) atomicOr(err, 4u);
}

// LL of (sample-read r, allele a): its pool's row, or -- when reads come in mate pairs (consecutive reads with the same
// name, P.second_mate) -- what src/seq_stutter_genotyper.cpp:546-559 leaves in the two rows: the sum of both mates,
// accumulated in read order along a run of flagged reads.
__device__ __forceinline__ double read_ll(const DevPosterior& P, const double* ll, uint32_t H, uint32_t r0, uint32_t r1,
                                          uint32_t r, uint32_t a) {
  if (P.second_mate == nullptr) return ll[(size_t)P.pool_index[r] * H + a];
  uint32_t c0 = r;
  while (c0 > r0 && P.second_mate[c0]) --c0;  // first read of the run
  uint32_t last = r;                          // the row of read r is last rewritten when read r+1 is a second mate
  if (r + 1 < r1 && P.second_mate[r + 1]) last = r + 1;
  double s = ll[(size_t)P.pool_index[c0] * H + a];
  for (uint32_t j = c0 + 1; j <= last; ++j) s = s + ll[(size_t)P.pool_index[j] * H + a];
  return s;
}

// One warp per locus.  exp(LL[r][a] + log_p1[r] + log 1/2) depends on (read, allele) only and the term
// log(e1[r][a] + e2[r][b]) on (read, a, b) only, so the warp first tabulates the 2 R H exponentials and the R H^2 logarithms
// with all lanes (shared memory, kPostDoubles per warp), and every (sample, a, b) entry is then summed by one lane over its
// sample's reads in storage order -- the reference's order of summation, on values that are bit for bit what the
// straightforward loop computes.  Loci whose tables do not fit fall back to fewer tables and finally to that loop.
static constexpr int kPostWarps = 4;
static constexpr uint32_t kPostDoubles = 1024;

__global__ void __launch_bounds__(kPostWarps * 32) posterior_kernel(const DevPosterior P) {
  extern __shared__ __align__(16) double post_smem[];
  if (P.err && *P.err != 0u) return;
  const uint32_t lane = threadIdx.x & 31u;
  double* buf = post_smem + (size_t)(threadIdx.x >> 5) * kPostDoubles;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t l = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; l < P.n_loci; l += warps) {
    const uint32_t H = P.locus_hap_begin[l + 1] - P.locus_hap_begin[l];
    const uint32_t S = P.locus_n_samples[l];
    const uint32_t r0 = P.locus_sread_begin[l], r1 = P.locus_sread_begin[l + 1];
    const uint32_t R = r1 - r0;
    const bool haploid = P.locus_haploid ? (P.locus_haploid[l] != 0) : false;
    const double* ll = P.ll + P.ll_off[l];
    double* post = P.post + P.post_off[l];
    double* tot = P.totals + P.tot_off[l];
    const uint32_t HH = H * H;
    if (H == 0 || S == 0) continue;
    double hom, het;
    const double lH = P.int_logs[H < P.n_int_logs ? H : 0], lH1 = P.int_logs[H + 1 < P.n_int_logs ? H + 1 : 0];
    if (haploid) {
      hom = -lH;
      het = -DBL_MAX / 2;
    } else {
      hom = P.int_logs[2] - lH - lH1;
      het = -lH - lH1;
    }
    const unsigned long long n_e = 2ull * R * H, n_t = (unsigned long long)R * HH;
    const bool have_e = n_e <= kPostDoubles, have_t = have_e && (n_e + n_t <= kPostDoubles);
    double* E1 = buf;
    double* E2 = buf + (size_t)R * H;
    double* T = buf + 2 * (size_t)R * H;
    __syncwarp();
    if (have_e) {
      for (uint32_t i = lane; i < R * H; i += 32u) {
        const uint32_t r = i / H, a = i - r * H;
        double v = read_ll(P, ll, H, r0, r1, r0 + r, a);
        v = (v < -600.0) ? -600.0 : v;  // genotyper.cpp:57-58
        E1[i] = exp(v + P.log_p1[r0 + r] + P.log_one_half);
        E2[i] = exp(v + P.log_p2[r0 + r] + P.log_one_half);
      }
      __syncwarp();
    }
    if (have_t) {
      for (uint32_t i = lane; i < R * HH; i += 32u) {
        const uint32_t r = i / HH, ab = i - r * HH, a = ab / H, b = ab - a * H;
        T[i] = log(E1[r * H + a] + E2[r * H + b]);
      }
      __syncwarp();
    }
    for (uint32_t idx = lane; idx < S * HH; idx += 32u) {
      const uint32_t s = idx / HH, ab = idx - s * HH, a = ab / H, b = ab - a * H;
      double acc = (a == b) ? hom : het;
      for (uint32_t r = 0; r < R; ++r) {
        if ((uint32_t)P.sample_label[r0 + r] != s) continue;
        if (have_t) {
          acc += T[r * HH + ab];
        } else if (have_e) {
          acc += log(E1[r * H + a] + E2[r * H + b]);
        } else {
          double la = read_ll(P, ll, H, r0, r1, r0 + r, a), lb = read_ll(P, ll, H, r0, r1, r0 + r, b);
          la = (la < -600.0) ? -600.0 : la;
          lb = (lb < -600.0) ? -600.0 : lb;
          acc += log(exp(la + P.log_p1[r0 + r] + P.log_one_half) + exp(lb + P.log_p2[r0 + r] + P.log_one_half));
        }
      }
      post[idx] = acc;
    }
    __syncwarp();
    for (uint32_t s = lane; s < S; s += 32u) {  // log_sum_exp in storage order (mathops.cpp:45-51)
      const double* v = post + (size_t)s * HH;
      double mx = v[0];
      for (uint32_t k = 1; k < HH; ++k) mx = (mx < v[k]) ? v[k] : mx;
      double sum = 0.0;
      for (uint32_t k = 0; k < HH; ++k) sum += exp(v[k] - mx);
      tot[s] = mx + log(sum);
    }
    __syncwarp();
    for (uint32_t idx = lane; idx < S * HH; idx += 32u) post[idx] -= tot[idx / HH];
    if (P.kept_mask == nullptr) continue;
    // ---- what SeqStutterGenotyper::genotype does next (src/seq_stutter_genotyper.cpp:636-645): non-reference alleles in
    // no voting sample's optimal pair are dropped and the posteriors recomputed on the surviving ones ------------------
    __syncwarp();
    const uint32_t h0 = P.locus_hap_begin[l];
    uint8_t* kept = P.kept_mask + h0;
    uint32_t* kidx = P.kept_index + h0;
    for (uint32_t a = lane; a < H; a += 32u) kept[a] = (a == 0) ? 1 : 0;
    __syncwarp();
    for (uint32_t s = lane; s < S; s += 32u) {
      bool votes = false;  // the sample has an aligned read (get_unused_alleles, :262-265)
      for (uint32_t r = r0; r < r1 && !votes; ++r)
        votes = ((uint32_t)P.sample_label[r] == s) && (P.read_aligned == nullptr || P.read_aligned[r] != 0);
      if (!votes) continue;
      const double* v = post + (size_t)s * HH;  // get_optimal_haplotypes (genotyper.cpp:85-100): first maximum
      double best = -DBL_MAX;
      uint32_t arg = 0;
      for (uint32_t k = 0; k < HH; ++k)
        if (v[k] > best) {
          best = v[k];
          arg = k;
        }
      kept[arg / H] = 1;
      kept[arg % H] = 1;
    }
    __syncwarp();
    uint32_t K = 0;
    if (lane == 0) {
      for (uint32_t a = 0; a < H; ++a)
        if (kept[a]) kidx[K++] = a;
    }
    K = __shfl_sync(0xFFFFFFFFu, K, 0);
    __syncwarp();
    if (K == H) continue;  // nothing removed: the first pass stands
    const uint32_t KK = K * K;
    const double lK = P.int_logs[K < P.n_int_logs ? K : 0], lK1 = P.int_logs[K + 1 < P.n_int_logs ? K + 1 : 0];
    const double hom2 = haploid ? -lK : P.int_logs[2] - lK - lK1;
    const double het2 = haploid ? -DBL_MAX / 2 : -lK - lK1;
    // every lane computes its entries from the tables / the LL rows first, the compact array is written afterwards
    // (it overlaps the first pass' array, which the LL rows do not depend on)
    for (uint32_t base = 0; base < S * KK; base += 32u) {
      const uint32_t idx = base + lane;
      double acc = 0.0;
      if (idx < S * KK) {
        const uint32_t s = idx / KK, ab = idx - s * KK, a2 = ab / K, b2 = ab - a2 * K;
        const uint32_t a = kidx[a2], b = kidx[b2], abo = a * H + b;
        acc = (a2 == b2) ? hom2 : het2;
        for (uint32_t r = 0; r < R; ++r) {
          if ((uint32_t)P.sample_label[r0 + r] != s) continue;
          if (have_t) {
            acc += T[r * HH + abo];
          } else if (have_e) {
            acc += log(E1[r * H + a] + E2[r * H + b]);
          } else {
            double la = read_ll(P, ll, H, r0, r1, r0 + r, a), lb = read_ll(P, ll, H, r0, r1, r0 + r, b);
            la = (la < -600.0) ? -600.0 : la;
            lb = (lb < -600.0) ? -600.0 : lb;
            acc += log(exp(la + P.log_p1[r0 + r] + P.log_one_half) + exp(lb + P.log_p2[r0 + r] + P.log_one_half));
          }
        }
      }
      __syncwarp();
      if (idx < S * KK) post[idx] = acc;
    }
    __syncwarp();
    for (uint32_t s = lane; s < S; s += 32u) {
      const double* v = post + (size_t)s * KK;
      double mx = v[0];
      for (uint32_t k = 1; k < KK; ++k) mx = (mx < v[k]) ? v[k] : mx;
      double sum = 0.0;
      for (uint32_t k = 0; k < KK; ++k) sum += exp(v[k] - mx);
      tot[s] = mx + log(sum);
    }
    __syncwarp();
    for (uint32_t idx = lane; idx < S * KK; idx += 32u) post[idx] -= tot[idx / KK];
  }
}

// 8 lanes per pooled read walk its H haplotype columns (H is 2-12 in practice): coalesced within the row.
__global__ void __launch_bounds__(256) expand_ll_kernel(const ExpandArgs E) {
  const uint32_t sub = threadIdx.x & 7u;
  if (E.err && *E.err != 0u) return;
  for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; r < E.n_reads; r += (gridDim.x * blockDim.x) >> 3) {
    const uint32_t l = E.read_locus[r];
    const uint32_t H = E.locus_hap_begin[l + 1] - E.locus_hap_begin[l];
    const double* src = E.uniq_ll + E.ull_off[l] + (size_t)(E.read_to_uread[r] - E.locus_uread_begin[l]) * H;
    double* dst = E.out_ll + E.ll_off[l] + (size_t)(r - E.locus_read_begin[l]) * H;
    for (uint32_t h = sub; h < H; h += 8u) dst[h] = src[h];
  }
}

cudaError_t launch_expand_ll(const ExpandArgs& E, cudaStream_t stream) {
  if (E.n_reads == 0) return cudaSuccess;
  const uint64_t want = ((uint64_t)E.n_reads * 8u + 255u) / 256u;
  const uint32_t grid = (uint32_t)(want < 148u * 32u ? want : 148u * 32u);
  expand_ll_kernel<<<grid, 256, 0, stream>>>(E);
  return cudaGetLastError();
}

cudaError_t launch_posterior_validate(const DevPosterior& P, uint32_t* err, cudaStream_t stream) {
  if (P.n_loci == 0) return cudaSuccess;
  const uint32_t grid = P.n_loci < 148u * 16u ? P.n_loci : 148u * 16u;
  posterior_validate_kernel<<<grid, 128, 0, stream>>>(P, err);
  return cudaGetLastError();
}

cudaError_t launch_posteriors(const DevPosterior& P, cudaStream_t stream) {
  if (P.n_loci == 0) return cudaSuccess;
  const size_t smem = (size_t)kPostWarps * kPostDoubles * sizeof(double);
  cudaFuncSetAttribute(posterior_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  // per device
  const uint32_t want = (P.n_loci + kPostWarps - 1) / kPostWarps;
  const uint32_t grid = want < 148u * 8u ? want : 148u * 8u;
  posterior_kernel<<<grid, kPostWarps * 32, smem, stream>>>(P);
  return cudaGetLastError();
}

}  // namespace ltr
