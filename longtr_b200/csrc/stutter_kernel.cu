// stutter_kernel.cu -- sm_100a kernel of the homopolymer / --stutter-align-len path (kernel 2).
//
// One warp per (pooled read, candidate haplotype) pair; see stutter_core.cuh for the per-lane algorithm and the
// reference lines it follows (HapAligner.cpp:27-233, 855-975; StutterAlignerClass.cpp).  Each warp stages the
// read flank (bases + per-base log P(correct/error)), the allele, its periodicity tables and load_read's match[]
// in shared memory, runs the two wavefront phases and the stutter row for the left flank, the same for the
// reversed right flank, and joins the two at the seed base.  FP64 max-plus + a float bit-trick log-sum-exp; no
// tensor cores (not a contraction).  The matrices of the reference are never materialised: only one row
// (hand-off between phases / strips) and the last column (consumed by the seed join) are kept.
//
// Code layout follows two measurements (profiles/r1i, r1j): (1) the phases are single noinline functions so
// that the kernel stays inside the instruction caches (the fully inlined first version stalled ~10 cycles per
// issue on instruction fetch); (2) nothing is passed to them by reference on the thread stack -- with ~55 KB of
// shared memory per CTA there is almost no L1 left, so local-memory traffic goes to L2 (5.5 cycles of
// long-scoreboard stall per issue in the second version).  All per-warp context lives in shared memory and is
// copied into registers at the top of each phase; the constants and the log table are staged per CTA.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"
#include "stutter_core.cuh"

namespace ltr {

static constexpr unsigned kFull = 0xFFFFFFFFu;
static constexpr int kStutWarps = 4;

// Everything a phase touches lives in shared memory, but the phases are separate (noinline) functions that receive
// their pointers through structs, so the compiler sees generic pointers: generic LD + 64-bit address arithmetic
// (profiles/r2a: 8.5 % LD, 9 % R2UR, no LDS in the hot loops).  Telling it the address space turns them into LDS / STS
// with 32-bit addresses.
#define LTR_ASSUME_SHARED(p) __builtin_assume(__isShared(p))
#define LTR_ASSUME_GLOBAL(p) __builtin_assume(__isGlobal(p))
__device__ __forceinline__ void assume_shared(const FlankView& F) {
  LTR_ASSUME_SHARED(F.seq); LTR_ASSUME_SHARED(F.qual); LTR_ASSUME_SHARED(F.tlc); LTR_ASSUME_SHARED(F.tlw);
  LTR_ASSUME_SHARED(F.blk); LTR_ASSUME_SHARED(F.um); LTR_ASSUME_SHARED(F.match); LTR_ASSUME_SHARED(F.art_lp);
}
__device__ __forceinline__ void assume_shared(const StutConsts& C) {
  LTR_ASSUME_SHARED(C.int_logs); LTR_ASSUME_SHARED(C.qual_lc); LTR_ASSUME_SHARED(C.qual_lw);
}

struct SideGeom {  // one pair
  const uint8_t* read;  // whole read
  const uint8_t* qual;
  int32_t N, seed;
  const uint8_t* lflank;
  const uint8_t* rflank;
  const uint8_t* allele;
  int32_t n0, n2, B;  // |lflank|, |rflank|, |allele|
};

struct WarpCtx {  // per warp, in shared memory
  double* match;
  double* lineM;
  double* lineD;
  double* lastL;
  double* lastR;
  double* probs;  // [7][32] artifact terms D = -6..0 of the stutter row, one column per lane
  int32_t* um;
  uint8_t* seq;
  uint8_t* qual;
  uint8_t* blk;
  double art_lp[13];
  SideGeom G;
  FlankView F;
};

__host__ __device__ inline size_t stutter_warp_smem_bytes(uint32_t max_flank, uint32_t max_block, uint32_t max_hap) {
  size_t b = (sizeof(WarpCtx) + 15) / 16 * 16;
  b += 3 * (size_t)max_flank * sizeof(double);  // match lineM lineD
  b += 2 * (size_t)max_hap * sizeof(double);    // lastL lastR
  b += 7 * 32 * sizeof(double);                 // probs (deletion sizes and D = 0; the insertion sizes stay in registers)
  b += 6 * (size_t)max_block * sizeof(int32_t); // um
  b += 2 * (((size_t)max_flank + 15) / 16 * 16);  // seq, qual
  b += ((size_t)max_block + 15) / 16 * 16;      // blk
  return (b + 15) / 16 * 16;
}
__host__ __device__ inline size_t stutter_cta_smem_bytes(uint32_t n_logs) {  // constants, log table, 2 quality tables
  return (sizeof(StutConsts) + ((size_t)n_logs + 512) * sizeof(double) + 15) / 16 * 16;
}

__device__ __forceinline__ WarpCtx* carve(unsigned char* base, uint32_t max_flank, uint32_t max_block, uint32_t max_hap,
                                          int lane) {
  WarpCtx* X = reinterpret_cast<WarpCtx*>(base);
  if (lane == 0) {
    double* d = reinterpret_cast<double*>(base + (sizeof(WarpCtx) + 15) / 16 * 16);
    X->match = d; d += max_flank;
    X->lineM = d; d += max_flank;
    X->lineD = d; d += max_flank;
    X->lastL = d; d += max_hap;
    X->lastR = d; d += max_hap;
    X->probs = d; d += 7 * 32;
    X->um = reinterpret_cast<int32_t*>(d);
    uint8_t* u = reinterpret_cast<uint8_t*>(X->um + 6 * (size_t)max_block);
    X->seq = u;
    X->qual = u + ((size_t)max_flank + 15) / 16 * 16;
    X->blk = X->qual + ((size_t)max_flank + 15) / 16 * 16;
  }
  return X;
}

// Wavefront over `nrows` consecutive flank rows whose first hap row is row0.  The haplotype character of row r of
// the block is chars[r * dir] (dir = -1 walks a reversed block).  first_type is the type of the very first row
// (ROW_FIRST / ROW_AFTER_STUTTER); every other row is ROW_NORMAL.  lineM/lineD hold the row above on entry and the
// last row on exit; last[] receives M at the last column of every row.  Returns left_prob for ROW_FIRST.
__device__ __noinline__ double wavefront_rows(const StutConsts* Cs, const WarpCtx* X, double* last, int32_t row0,
                                              int32_t nrows, int32_t first_type, const uint8_t* chars, int32_t dir,
                                              int lane) {
  LTR_ASSUME_SHARED(Cs); LTR_ASSUME_SHARED(X); LTR_ASSUME_SHARED(last); LTR_ASSUME_GLOBAL(chars);
  const StutConsts C = *Cs;  // registers
  const FlankView F = X->F;
  assume_shared(C);
  assume_shared(F);
  double* lineM = X->lineM;
  double* lineD = X->lineD;
  LTR_ASSUME_SHARED(lineM); LTR_ASSUME_SHARED(lineD);
  double left_prob = 0.0;
  const int32_t per_strip = 32 * kStutRows;
  for (int32_t s0 = 0; s0 < nrows; s0 += per_strip) {
    const int32_t rows = (nrows - s0 < per_strip) ? (nrows - s0) : per_strip;
    const int32_t t_last = (rows - 1) / kStutRows;
    FlankLane Ln;
    flank_lane_reset(Ln);
#pragma unroll
    for (int k = 0; k < kStutRows; ++k) {
      const int32_t r = lane * kStutRows + k;
      if (r < rows) {
        Ln.type[k] = (s0 + r == 0) ? first_type : ROW_NORMAL;
        Ln.hc[k] = (int32_t)chars[(s0 + r) * dir];
      }
    }
    const int32_t nsteps = F.L + t_last;
    for (int32_t step = 0; step < nsteps; ++step) {
      double aM = __shfl_up_sync(kFull, Ln.outM, 1);
      double aD = __shfl_up_sync(kFull, Ln.outD, 1);
      const int32_t j = step - lane;
      if (lane <= t_last && j >= 0 && j < F.L) {
        if (lane == 0) {
          aM = lineM[j];
          aD = lineD[j];
        }
        double Mout[kStutRows];
        flank_lane_column(Ln, C, F, j, aM, aD, Mout);
        if (lane == t_last) {  // hand-off line for the next strip / phase (written after lane 0 consumed entry j)
          lineM[j] = Ln.outM;
          lineD[j] = Ln.outD;
        }
        if (j == F.L - 1) {
#pragma unroll
          for (int k = 0; k < kStutRows; ++k)
            if (Ln.type[k] != ROW_OFF) last[row0 + s0 + lane * kStutRows + k] = Mout[k];
        }
      }
      // lane 0 read entry `step` of the hand-off line in this step, lane t_last overwrites it t_last steps from now: the
      // shuffles keep the lanes in step, this orders the shared-memory accesses as well (compute-sanitizer racecheck)
      __syncwarp();
    }
    if (first_type == ROW_FIRST && s0 == 0) left_prob = __shfl_sync(kFull, Ln.left, 0);
    __syncwarp();
  }
  return left_prob;
}

// The stutter row (HapAligner.cpp:64-111): one read column per lane, 13 artifact sizes each; prevM = lineM,
// result -> lineD.
__device__ __noinline__ void stutter_row(const StutConsts* Cs, const WarpCtx* X, int lane) {
  LTR_ASSUME_SHARED(Cs); LTR_ASSUME_SHARED(X);
  const StutConsts C = *Cs;
  const FlankView F = X->F;
  assume_shared(C);
  assume_shared(F);
  const double* prevM = X->lineM;
  double* out = X->lineD;
  double* probs = X->probs + lane;  // stride 32: conflict free
  LTR_ASSUME_SHARED(prevM); LTR_ASSUME_SHARED(out); LTR_ASSUME_SHARED(probs);
  for (int32_t j = lane; j < F.L; j += 32) {
    double vi[6];  // insertions: the six sizes share one walk over the block (stutter_insertion_terms); kept in registers
    {
      double ins_ll[6];
      stutter_insertion_lls(C, F, j, ins_ll);
#pragma unroll
      for (int32_t d = 0; d < 6; ++d) {
        const int32_t D = d + 1;
        int32_t base_len = F.B + D;
        base_len = (base_len < j + 1) ? base_len : (j + 1);
        const double pre = (j - base_len < 0) ? 0.0 : prevM[j - base_len];
        vi[d] = (F.art_lp[7 + d] + ins_ll[d]) + pre;
      }
    }
#pragma unroll 1
    for (int32_t a = 0; a <= 6; ++a) {  // deletions and the artifact-free alignment: one copy of the code, values in smem
      const int32_t D = a - 6;
      int32_t base_len = F.B + D;
      base_len = (base_len < j + 1) ? base_len : (j + 1);
      double v = kStutImpossible;
      if (base_len >= 0) {
        const double prob = stutter_region_ll(C, F, base_len, j, D);
        const double pre = (j - base_len < 0) ? 0.0 : prevM[j - base_len];
        v = (F.art_lp[a] + prob) + pre;
      }
      probs[a * 32] = v;
    }
    double mx = probs[0];
#pragma unroll
    for (int32_t a = 1; a <= 6; ++a) mx = smax(mx, probs[a * 32]);
#pragma unroll
    for (int32_t d = 0; d < 6; ++d) mx = smax(mx, vi[d]);
    double total = 0.0;
#pragma unroll
    for (int32_t a = 0; a <= 6; ++a) total += lse_term(C, probs[a * 32], mx);
#pragma unroll
    for (int32_t d = 0; d < 6; ++d) total += lse_term(C, vi[d], mx);
    out[j] = lse_finish(mx, total);
  }
}

// Runs one flank (side 0 = left of the seed against the forward haplotype, side 1 = right of the seed, reversed,
// against the reversed haplotype).  Fills last[] (last-column M of every reachable hap row), returns left_prob.
__device__ __noinline__ double run_side(const StutConsts* Cs, WarpCtx* X, int side, int lane) {
  LTR_ASSUME_SHARED(Cs); LTR_ASSUME_SHARED(X);
  const SideGeom G = X->G;
  LTR_ASSUME_GLOBAL(G.read); LTR_ASSUME_GLOBAL(G.qual); LTR_ASSUME_GLOBAL(G.allele); LTR_ASSUME_GLOBAL(G.lflank);
  LTR_ASSUME_GLOBAL(G.rflank);
  const int32_t L = side == 0 ? G.seed : (G.N - G.seed - 1);
  const int32_t B = G.B;
  double* last = side ? X->lastR : X->lastL;
  LTR_ASSUME_SHARED(last);
  // ---- stage the flank, the allele and the tables ---------------------------------------------------------
  {
    uint8_t* seq = X->seq;
    uint8_t* qual = X->qual;
    LTR_ASSUME_SHARED(seq); LTR_ASSUME_SHARED(qual);
    for (int32_t j = lane; j < L; j += 32) {
      const int32_t p = side == 0 ? j : (G.N - 1 - j);
      seq[j] = G.read[p];
      qual[j] = G.qual[p];
    }
    uint8_t* blk = X->blk;
    LTR_ASSUME_SHARED(blk);
    for (int32_t i = lane; i < B; i += 32) blk[i] = side == 0 ? G.allele[i] : G.allele[B - 1 - i];
  }
  __syncwarp();
  const int32_t n_del = B < 6 ? B : 6;
  if (lane < n_del) {  // num_upstream_matches at lag lane+1 (StutterAlignerClass.h:35-42)
    const int32_t lag = lane + 1;
    const uint8_t* blk = X->blk;
    int32_t* ml = X->um + (size_t)lane * B;
    LTR_ASSUME_SHARED(blk); LTR_ASSUME_SHARED(ml);
    int32_t run = 0;
    for (int32_t i = 0; i < B; ++i) {
      if (i < lag) run = 0;
      else run = (blk[i - lag] != blk[i]) ? 0 : run + 1;
      ml[i] = run;
    }
  }
  if (lane == 0) {
    FlankView F;
    F.seq = X->seq;
    F.qual = X->qual;
    F.tlc = Cs->qual_lc;
    F.tlw = Cs->qual_lw;
    F.L = L;
    F.blk = X->blk;
    F.B = B;
    F.um = X->um;
    F.n_del = n_del;
    F.match = X->match;
    F.art_lp = X->art_lp;
    X->F = F;
  }
  __syncwarp();
  {
    const FlankView F = X->F;
    assume_shared(F);
    double* match = X->match;
    LTR_ASSUME_SHARED(match);
    for (int32_t p = lane; p < L; p += 32) match[p] = stutter_match_prob(F, p);
  }
  __syncwarp();
  // ---- phase A: rows of the first flank block ----------------------------------------------------------------
  const int32_t na = side == 0 ? G.n0 : G.n2, nc = side == 0 ? G.n2 : G.n0;
  const uint8_t* fa = side == 0 ? G.lflank : G.rflank;
  const uint8_t* fc = side == 0 ? G.rflank : G.lflank;
  const int32_t dir = side == 0 ? 1 : -1;
  const double left_prob = wavefront_rows(Cs, X, last, 0, na, ROW_FIRST, side == 0 ? fa : fa + (na - 1), dir, lane);
  // ---- phase B: the stutter row ----------------------------------------------------------------------------------
  stutter_row(Cs, X, lane);
  __syncwarp();
  if (lane == 0) {
    double* t = X->lineM;  // the stutter row becomes the row above phase C; its I and D are IMPOSSIBLE (:104-105)
    X->lineM = X->lineD;
    X->lineD = t;
    last[na + B - 1] = X->lineM[L - 1];
  }
  __syncwarp();
  // ---- phase C: rows of the second flank block -----------------------------------------------------------------
  wavefront_rows(Cs, X, last, na + B, nc, ROW_AFTER_STUTTER, side == 0 ? fc : fc + (nc - 1), dir, lane);
  return left_prob;
}

__global__ void __launch_bounds__(kStutWarps * 32) stutter_pair_kernel(const StutConsts C, const StutterDevBatch Bt,
                                                                        uint32_t n_logs) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // ---- per-CTA constants: StutConsts with the log table redirected into shared memory -----------------------------
  StutConsts* Cs = reinterpret_cast<StutConsts*>(smem);
  double* logs = reinterpret_cast<double*>(smem + sizeof(StutConsts));
  double* tlc = logs + n_logs;
  double* tlw = tlc + 256;
  for (uint32_t i = threadIdx.x; i < n_logs; i += blockDim.x) logs[i] = C.int_logs[i];
  for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
    tlc[i] = C.qual_lc[i];
    tlw[i] = C.qual_lw[i];
  }
  if (threadIdx.x == 0) {
    StutConsts c = C;
    c.int_logs = logs;
    c.qual_lc = tlc;
    c.qual_lw = tlw;
    *Cs = c;
  }
  __syncthreads();
  const uint32_t ti = blockIdx.x * kStutWarps + warp;
  if (ti >= Bt.n_tasks) return;
  const StutterTask T = Bt.tasks[ti];
  WarpCtx* X = carve(smem + stutter_cta_smem_bytes(n_logs) +
                         (size_t)warp * stutter_warp_smem_bytes(Bt.max_flank, Bt.max_block, Bt.max_hap),
                     Bt.max_flank, Bt.max_block, Bt.max_hap, lane);
  if (lane == 0) {
    SideGeom G;
    const uint32_t ro = Bt.read_off[T.read];
    G.read = Bt.read_bytes + ro;
    G.qual = Bt.qual_bytes + ro;
    G.N = (int32_t)(Bt.read_off[T.read + 1] - ro);
    G.seed = Bt.read_seed[T.read];
    G.lflank = Bt.lflank_bytes + Bt.lflank_off[T.locus];
    G.n0 = (int32_t)(Bt.lflank_off[T.locus + 1] - Bt.lflank_off[T.locus]);
    G.rflank = Bt.rflank_bytes + Bt.rflank_off[T.locus];
    G.n2 = (int32_t)(Bt.rflank_off[T.locus + 1] - Bt.rflank_off[T.locus]);
    G.allele = Bt.allele_bytes + Bt.allele_off[T.allele];
    G.B = (int32_t)(Bt.allele_off[T.allele + 1] - Bt.allele_off[T.allele]);
    X->G = G;
  }
  if (lane < 13) X->art_lp[lane] = Bt.allele_artifact_lp[(size_t)T.allele * 13 + lane];
  __syncwarp();

  const double l_prob = run_side(Cs, X, 0, lane);
  __syncwarp();
  const double r_prob = run_side(Cs, X, 1, lane);
  __syncwarp();

  // ---- seed join (compute_aln_logprob, HapAligner.cpp:165-233) ----------------------------------------------------
  const SideGeom G = X->G;
  LTR_ASSUME_GLOBAL(G.read); LTR_ASSUME_GLOBAL(G.qual); LTR_ASSUME_GLOBAL(G.lflank); LTR_ASSUME_GLOBAL(G.rflank);
  const double* lastL = X->lastL;
  const double* lastR = X->lastR;
  LTR_ASSUME_SHARED(lastL); LTR_ASSUME_SHARED(lastR);
  const double log_thresh = Cs->log_thresh;
  StutConsts Cj;  // only log_thresh is used by lse_term
  Cj.log_thresh = log_thresh;
  const int32_t hapsize = G.n0 + G.B + G.n2;
  const int32_t seed_char = (int32_t)G.read[G.seed];
  const uint8_t sq = G.qual[G.seed];
  const double sc = Cs->qual_lc[sq], sw = Cs->qual_lw[sq];
  const double prior = -logs[G.n0 + G.n2];
  // term index u: 0 and 1 are the two "flank entirely outside" configurations, u >= 2 <-> hap position u-1
  double mx = 0.0;
  bool any = false;
  for (int pass = 0; pass < 2; ++pass) {
    double total = 0.0;
    for (int32_t u = lane; u < hapsize; u += 32) {
      double v;
      if (u == 0) {
        v = ((prior + (seed_char == (int32_t)G.lflank[0] ? sc : sw)) + l_prob) + lastR[hapsize - 2];
      } else if (u == 1) {
        v = ((prior + (seed_char == (int32_t)G.rflank[G.n2 - 1] ? sc : sw)) + r_prob) + lastL[hapsize - 2];
      } else {
        const int32_t i = u - 1;  // 1 .. hapsize-2
        if (i >= G.n0 && i < G.n0 + G.B) continue;
        const int32_t hc = (int32_t)(i < G.n0 ? G.lflank[i] : G.rflank[i - G.n0 - G.B]);
        v = ((prior + (seed_char == hc ? sc : sw)) + lastL[i - 1]) + lastR[hapsize - 2 - i];
      }
      if (pass == 0) {
        mx = any ? smax(mx, v) : v;
        any = true;
      } else {
        total += lse_term(Cj, v, mx);
      }
    }
    if (pass == 0) {
      // a lane may hold no term (all its positions inside the repeat block): reduce with flags
      for (int off = 16; off > 0; off >>= 1) {
        const double omx = __shfl_xor_sync(kFull, mx, off);
        const int oany = __shfl_xor_sync(kFull, (int)any, off);
        if (oany) {
          mx = any ? smax(mx, omx) : omx;
          any = true;
        }
      }
    } else {
      for (int off = 16; off > 0; off >>= 1) total += __shfl_xor_sync(kFull, total, off);  // exact: see stutter_core.cuh
      if (lane == 0) Bt.out_ll[T.out_index] = lse_finish(mx, total);
    }
  }
}

size_t stutter_block_smem_bytes(uint32_t max_flank, uint32_t max_block, uint32_t max_hap) {
  return stutter_cta_smem_bytes(max_hap + 16) + (size_t)kStutWarps * stutter_warp_smem_bytes(max_flank, max_block, max_hap);
}

cudaError_t launch_stutter(const StutConsts& C, const StutterDevBatch& B, cudaStream_t stream) {
  if (B.n_tasks == 0) return cudaSuccess;
  const uint32_t n_logs = B.max_hap + 16;  // log table entries used: <= max(block)+2 and <= hap length (stutter_abi.cu sizes it)
  const size_t smem = stutter_block_smem_bytes(B.max_flank, B.max_block, B.max_hap);
  cudaError_t e = cudaFuncSetAttribute(stutter_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const uint32_t grid = (B.n_tasks + kStutWarps - 1) / kStutWarps;
  stutter_pair_kernel<<<grid, kStutWarps * 32, smem, stream>>>(C, B, n_logs);
  return cudaGetLastError();
}

}  // namespace ltr
