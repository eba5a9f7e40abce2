// stutter_kernel.cu -- sm_100a kernel of the homopolymer / --stutter-align-len path (kernel 2).
//
// One warp per (pooled read, candidate haplotype) pair; see stutter_core.cuh for the per-lane algorithm and the
// reference lines it follows (HapAligner.cpp:27-233, 855-975; StutterAlignerClass.cpp).  Each warp stages the
// read flank (bases + per-base log P(correct/error)), the allele, its periodicity tables and load_read's match[]
// in shared memory, runs the two wavefront phases and the stutter row for the left flank, the same for the
// reversed right flank, and joins the two at the seed base.  FP64 max-plus + a float bit-trick log-sum-exp; no
// tensor cores (not a contraction).  The matrices of the reference are never materialised: only one row
// (hand-off between phases / strips) and the last column (consumed by the seed join) are kept.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"
#include "stutter_core.cuh"

namespace ltr {

static constexpr unsigned kFull = 0xFFFFFFFFu;
static constexpr int kStutWarps = 4;

struct WarpSmem {  // carved out of dynamic shared memory, per warp
  double* lc;
  double* lw;
  double* match;
  double* lineM;
  double* lineD;
  double* lastL;
  double* lastR;
  int32_t* um;
  uint8_t* seq;
  uint8_t* blk;
};

__host__ __device__ inline size_t stutter_warp_smem_bytes(uint32_t max_flank, uint32_t max_block, uint32_t max_hap) {
  size_t b = 0;
  b += 5 * (size_t)max_flank * sizeof(double);  // lc lw match lineM lineD
  b += 2 * (size_t)max_hap * sizeof(double);    // lastL lastR
  b += 6 * (size_t)max_block * sizeof(int32_t); // um
  b += ((size_t)max_flank + 15) / 16 * 16;      // seq
  b += ((size_t)max_block + 15) / 16 * 16;      // blk
  return (b + 15) / 16 * 16;
}

__device__ __forceinline__ WarpSmem carve(unsigned char* base, uint32_t max_flank, uint32_t max_block, uint32_t max_hap) {
  WarpSmem W;
  double* d = reinterpret_cast<double*>(base);
  W.lc = d; d += max_flank;
  W.lw = d; d += max_flank;
  W.match = d; d += max_flank;
  W.lineM = d; d += max_flank;
  W.lineD = d; d += max_flank;
  W.lastL = d; d += max_hap;
  W.lastR = d; d += max_hap;
  W.um = reinterpret_cast<int32_t*>(d);
  uint8_t* u = reinterpret_cast<uint8_t*>(W.um + 6 * (size_t)max_block);
  W.seq = u;
  W.blk = u + ((size_t)max_flank + 15) / 16 * 16;
  return W;
}

// Wavefront over `nrows` consecutive flank rows whose first hap row is row0; chars come from hapc(row).
// first_type is the type of the very first row (ROW_FIRST / ROW_AFTER_STUTTER); every other row is ROW_NORMAL.
// lineM/lineD hold the row above on entry (unused for ROW_FIRST / ROW_AFTER_STUTTER beyond M) and the last
// row on exit.  last[] receives M at the last column of every row.  Returns left_prob when first_type == ROW_FIRST.
// The haplotype character of row r of the block is chars[r * dir] (dir = -1 walks a reversed block).
// One copy in the binary (noinline): the whole kernel has to stay inside the instruction caches -- the first
// version inlined this four times and stalled ~10 cycles per issue on instruction fetch (profiles/r1i).
__device__ __noinline__ double wavefront_rows(const StutConsts& C, const FlankView& F, WarpSmem& W, double* last,
                                              int32_t row0, int32_t nrows, int32_t first_type, const uint8_t* chars,
                                              int32_t dir, int lane) {
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
          aM = W.lineM[j];
          aD = W.lineD[j];
        }
        double Mout[kStutRows];
        flank_lane_column(Ln, C, F, j, aM, aD, Mout);
        if (lane == t_last) {  // hand-off line for the next strip / phase (written after lane 0 consumed entry j)
          W.lineM[j] = Ln.outM;
          W.lineD[j] = Ln.outD;
        }
        if (j == F.L - 1) {
#pragma unroll
          for (int k = 0; k < kStutRows; ++k)
            if (Ln.type[k] != ROW_OFF) last[row0 + s0 + lane * kStutRows + k] = Mout[k];
        }
      }
    }
    if (first_type == ROW_FIRST && s0 == 0) left_prob = __shfl_sync(kFull, Ln.left, 0);
    __syncwarp();
  }
  return left_prob;
}

struct SideGeom {  // one flank of one pair
  const uint8_t* read;  // whole read
  const uint8_t* qual;
  int32_t N, seed;
  const uint8_t* lflank;
  const uint8_t* rflank;
  const uint8_t* allele;
  int32_t n0, n2, B;  // |lflank|, |rflank|, |allele|
};

// Runs one flank (side 0 = left of the seed against the forward haplotype, side 1 = right of the seed, reversed,
// against the reversed haplotype).  Fills last[] (last-column M of every reachable hap row), returns left_prob.
__device__ __noinline__ double run_side(const StutConsts& C, const SideGeom& G, int side, const double* art_lp,
                                        WarpSmem& W, double* last, int lane) {
  const int32_t L = side == 0 ? G.seed : (G.N - G.seed - 1);
  const int32_t B = G.B;
  // ---- stage the flank, the allele and the tables ---------------------------------------------------------
  for (int32_t j = lane; j < L; j += 32) {
    const int32_t p = side == 0 ? j : (G.N - 1 - j);
    const uint8_t q = G.qual[p];
    W.seq[j] = G.read[p];
    W.lc[j] = C.qual_lc[q];
    W.lw[j] = C.qual_lw[q];
  }
  for (int32_t i = lane; i < B; i += 32) W.blk[i] = side == 0 ? G.allele[i] : G.allele[B - 1 - i];
  __syncwarp();
  const int32_t n_del = B < 6 ? B : 6;
  if (lane < n_del) {  // num_upstream_matches at lag lane+1 (StutterAlignerClass.h:35-42)
    const int32_t lag = lane + 1;
    int32_t* ml = W.um + (size_t)lane * B;
    int32_t run = 0;
    for (int32_t i = 0; i < B; ++i) {
      if (i < lag) run = 0;
      else run = (W.blk[i - lag] != W.blk[i]) ? 0 : run + 1;
      ml[i] = run;
    }
  }
  FlankView F;
  F.seq = W.seq;
  F.lc = W.lc;
  F.lw = W.lw;
  F.L = L;
  F.blk = W.blk;
  F.B = B;
  F.um = W.um;
  F.n_del = n_del;
  F.match = W.match;
  F.art_lp = art_lp;
  for (int32_t p = lane; p < L; p += 32) W.match[p] = stutter_match_prob(F, p);
  __syncwarp();
  // ---- phase A: rows of the first flank block ----------------------------------------------------------------
  const int32_t na = side == 0 ? G.n0 : G.n2, nc = side == 0 ? G.n2 : G.n0;
  const uint8_t* fa = side == 0 ? G.lflank : G.rflank;
  const uint8_t* fc = side == 0 ? G.rflank : G.lflank;
  const int32_t dir = side == 0 ? 1 : -1;
  const uint8_t* chars_a = side == 0 ? fa : fa + (na - 1);
  const uint8_t* chars_c = side == 0 ? fc : fc + (nc - 1);
  const double left_prob = wavefront_rows(C, F, W, last, 0, na, ROW_FIRST, chars_a, dir, lane);
  // ---- phase B: the stutter row, one column per lane (HapAligner.cpp:64-111) ----------------------------------
  for (int32_t j = lane; j < L; j += 32) W.lineD[j] = stutter_row_cell(C, F, W.lineM, j);
  __syncwarp();
  {
    double* t = W.lineM;  // the stutter row becomes the row above phase C; its I and D are IMPOSSIBLE (:104-105)
    W.lineM = W.lineD;
    W.lineD = t;
  }
  if (lane == 0) last[na + B - 1] = W.lineM[L - 1];
  // ---- phase C: rows of the second flank block -----------------------------------------------------------------
  wavefront_rows(C, F, W, last, na + B, nc, ROW_AFTER_STUTTER, chars_c, dir, lane);
  return left_prob;
}

__global__ void __launch_bounds__(kStutWarps * 32) stutter_pair_kernel(const StutConsts C, const StutterDevBatch Bt) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t ti = blockIdx.x * kStutWarps + warp;
  if (ti >= Bt.n_tasks) return;
  const StutterTask T = Bt.tasks[ti];
  WarpSmem W = carve(smem + (size_t)warp * stutter_warp_smem_bytes(Bt.max_flank, Bt.max_block, Bt.max_hap), Bt.max_flank,
                     Bt.max_block, Bt.max_hap);
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
  const double* art_lp = Bt.allele_artifact_lp + (size_t)T.allele * 13;

  double side_prob[2];
#pragma unroll 1
  for (int side = 0; side < 2; ++side) side_prob[side] = run_side(C, G, side, art_lp, W, side ? W.lastR : W.lastL, lane);
  const double l_prob = side_prob[0], r_prob = side_prob[1];
  __syncwarp();

  // ---- seed join (compute_aln_logprob, HapAligner.cpp:165-233) ----------------------------------------------------
  const int32_t hapsize = G.n0 + G.B + G.n2;
  const int32_t seed_char = (int32_t)G.read[G.seed];
  const uint8_t sq = G.qual[G.seed];
  const double sc = C.qual_lc[sq], sw = C.qual_lw[sq];
  const double prior = -C.int_logs[G.n0 + G.n2];
  // term index u: 0 and 1 are the two "flank entirely outside" configurations, u >= 2 <-> hap position u-1
  double mx = 0.0;
  bool any = false;
  for (int pass = 0; pass < 2; ++pass) {
    double total = 0.0;
    for (int32_t u = lane; u < hapsize; u += 32) {
      double v;
      if (u == 0) {
        v = ((prior + (seed_char == (int32_t)G.lflank[0] ? sc : sw)) + l_prob) + W.lastR[hapsize - 2];
      } else if (u == 1) {
        v = ((prior + (seed_char == (int32_t)G.rflank[G.n2 - 1] ? sc : sw)) + r_prob) + W.lastL[hapsize - 2];
      } else {
        const int32_t i = u - 1;  // 1 .. hapsize-2
        if (i >= G.n0 && i < G.n0 + G.B) continue;
        const int32_t hc = (int32_t)(i < G.n0 ? G.lflank[i] : G.rflank[i - G.n0 - G.B]);
        v = ((prior + (seed_char == hc ? sc : sw)) + W.lastL[i - 1]) + W.lastR[hapsize - 2 - i];
      }
      if (pass == 0) {
        mx = any ? smax(mx, v) : v;
        any = true;
      } else {
        total += lse_term(C, v, mx);
      }
    }
    if (pass == 0) {
      // every lane < min(32, hapsize) has at least one term only if it is not inside the repeat block: reduce with flags
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
  return (size_t)kStutWarps * stutter_warp_smem_bytes(max_flank, max_block, max_hap);
}

cudaError_t launch_stutter(const StutConsts& C, const StutterDevBatch& B, cudaStream_t stream) {
  if (B.n_tasks == 0) return cudaSuccess;
  const size_t smem = stutter_block_smem_bytes(B.max_flank, B.max_block, B.max_hap);
  cudaError_t e = cudaFuncSetAttribute(stutter_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const uint32_t grid = (B.n_tasks + kStutWarps - 1) / kStutWarps;
  stutter_pair_kernel<<<grid, kStutWarps * 32, smem, stream>>>(C, B);
  return cudaGetLastError();
}

}  // namespace ltr
