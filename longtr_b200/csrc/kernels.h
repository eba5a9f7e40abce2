// kernels.h -- launch interface between the C-ABI implementation (abi.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "band_core.cuh"
#include "plan_device.cuh"
#include "viterbi_core.cuh"

namespace ltr {

int viterbi_max_rows_per_lane();
int viterbi_block_threads();
int viterbi_blocks_per_sm(int k, int mode);
uint32_t viterbi_scratch_entries(uint32_t q);
cudaError_t launch_viterbi(int k, int mode, int grid_blocks, cudaStream_t stream, const VitConsts& C,
                           const DevBatch& B, const Task* tasks, const uint32_t* ntasks_ptr,
                           uint32_t task_cap, uint32_t* cursor, const FailSink& fail, XY* sxy,
                           uint32_t* sb, uint32_t scratch_stride);

// Banded anti-diagonal Viterbi (band_kernel.cu, band_core.cuh): rounds of 32 / G pairs per warp.
struct BandArgs {
  const uint2* pairs;       // (haplotype, unique read) pairs of ALL band classes, class after class
  const uint32_t* info;     // device words {first pair, number of pairs} of this launch's band class (written by the
                            // host plan's upload or by the device plan: the launch needs no count on the host)
  uint32_t* cursor;         // round cursor of the persistent grid
  uint32_t* counters;       // [0] pairs evaluated, [1] of which not certified (shared by all band classes of a job)
  unsigned long long* cells_evaluated;  // interior band cells of the pairs evaluated (statistics)
  BandGap gap;              // unavoidable gap costs of the certificate (band_core.cuh)
  uint32_t abandon_after;   // 0: never; else stop banding once this many pairs were evaluated and most failed
};
struct BandCollect {        // where band_collect_kernel appends the uncertified runs: task list of each row class
  Task* tasks[17];
  uint32_t* count[17];
  uint32_t cap[17];
  int kmax;
  unsigned long long* n_uncertified;
  unsigned long long* cells_uncertified;
  uint32_t* bucket_count;   // [17*32] runs per (row class, floor(log2 cost)); zeroed before the collect
  uint32_t* bucket_base;    // [17*32] list position of each bucket (band_bucket_scan_kernel)
  uint32_t* bucket_fill;    // [17*32] zeroed before the collect
  // second chance in a wider band (band_retry_class); retry_pairs == NULL: every uncertified pair goes to the stream kernel
  uint2* retry_pairs;       // pair lists of the second band round, class after class
  uint32_t retry_cap;
  uint32_t* retry_count;    // [8] zeroed before the collect
  uint32_t* retry_fill;     // [8] zeroed before the collect
  uint32_t* retry_info;     // [16] {first pair, number of pairs} per band class (band_bucket_scan_kernel)
  unsigned long long* n_retried;       // statistics: pairs given a second band round ...
  unsigned long long* n_retry_failed;  // ... and not certified by it (expected: 0)
  BandGap gap;
  int retry_rho_pct;        // a retry must cost less than this share of the full matrix
};
int band_block_threads();
int band_blocks_per_sm(int cls);  // cls: band class index (band_core.cuh)
cudaError_t launch_band(int cls, int grid_blocks, cudaStream_t stream, const VitConsts& C, const DevBatch& B,
                        const BandArgs& A);
cudaError_t launch_band_expand(const BandTask* tasks, const uint32_t* cum, uint32_t n_tasks, uint2* pairs,
                               cudaStream_t stream);
// n_tasks_ptr: device word holding the number of band tasks (<= task_cap)
cudaError_t launch_band_collect(const VitConsts& C, const DevBatch& B, const BandTask* tasks, const uint32_t* n_tasks_ptr,
                                uint32_t task_cap, const BandCollect& S, int sm_count, cudaStream_t stream);
cudaError_t launch_band_retry_check(const VitConsts& C, const DevBatch& B, const BandCollect& S, int sm_count,
                                    cudaStream_t stream);

// The job plan on the device (plan_kernels.cu, plan_device.cuh): five small kernels in stream order.
cudaError_t launch_device_plan(const PlanDev& P, int sm_count, cudaStream_t stream);

// Fan-out of the unique LL matrices (one row per distinct trimmed read of a locus) to the reference's
// aln_probs[read*H + hap] layout (HapAligner.cpp:549): out[ll_off(l) + (p-rb0)*H + h] = uniq[ull_off(l) + u(p)*H + h].
struct ExpandArgs {
  uint32_t n_reads;
  const uint32_t* read_locus;
  const uint32_t* read_to_uread;
  const uint32_t* locus_hap_begin;
  const uint32_t* locus_read_begin;
  const uint32_t* locus_uread_begin;
  const unsigned long long* ll_off;
  const unsigned long long* ull_off;
  const double* uniq_ll;
  double* out_ll;
  const uint32_t* err;  // device plan: != 0 -> malformed batch, nothing to expand (may be NULL)
};
cudaError_t launch_expand_ll(const ExpandArgs& E, cudaStream_t stream);

// Posterior step for a batch of loci (Genotyper::calc_log_sample_posteriors, genotyper.cpp:45-83).
struct DevPosterior {
  uint32_t n_loci;
  const uint32_t* locus_hap_begin;    // H_l
  const uint32_t* locus_read_begin;   // pooled reads of locus l (bounds of pool_index)
  const uint32_t* locus_sread_begin;  // sample-reads of locus l
  const uint32_t* pool_index;         // pooled read (relative to the locus) of each sample-read
  const int32_t* sample_label;
  const double* log_p1;
  const double* log_p2;
  const uint32_t* locus_n_samples;
  const uint8_t* locus_haploid;       // may be NULL
  const unsigned long long* ll_off;   // [n_loci+1] offsets of the pooled LL matrices
  const unsigned long long* post_off; // [n_loci+1] offsets into post (sum S*H*H)
  const unsigned long long* tot_off;  // [n_loci+1] offsets into totals (sum S)
  const double* ll;                   // pooled LL matrices (not modified; clamp applied on the fly)
  const double* int_logs;             // log(k), k < n_int_logs, computed on the host with libm
  uint32_t n_int_logs;
  double log_one_half;                // host libm log(0.5) (mathops.cpp:10)
  double* post;
  double* totals;
  const uint32_t* err;                // != 0: malformed batch, nothing is computed (may be NULL)
  const uint8_t* second_mate;         // [n_sreads] or NULL: read r is the second mate of read r-1 (LL rows are summed)
  const uint8_t* read_aligned;        // [n_sreads] or NULL (all): seed_positions >= 0 -- whether the read lets its sample vote
  uint8_t* kept_mask;                 // [n_haps] or NULL: != NULL turns on the removal of uncalled alleles + second pass;
  uint32_t* kept_index;               //   kept_mask[h] = allele survives; kept_index: scratch, [n_haps]
};
cudaError_t launch_posteriors(const DevPosterior& P, cudaStream_t stream);
// Sets bit 2 of *err when a sample-read points outside its locus' pooled reads or names a sample that does not exist
// (jobs planned on the device: the host never walks the per-read arrays).
cudaError_t launch_posterior_validate(const DevPosterior& P, uint32_t* err, cudaStream_t stream);

// Homopolymer / --stutter-align-len path (stutter_kernel.cu): one warp per (read, allele) pair.
struct StutConsts;
struct StutterTask {
  uint32_t locus, read, allele;
  unsigned long long out_index;  // position in out_ll
};
struct StutterDevBatch {
  uint32_t n_tasks;
  const StutterTask* tasks;
  const uint32_t* read_off;     // [n_reads+1] whole (untrimmed) pooled reads
  const uint8_t* read_bytes;
  const uint8_t* qual_bytes;    // Phred+33, same offsets
  const int32_t* read_seed;     // [n_reads] HapAligner::calc_seed_base
  const uint32_t* lflank_off;   // [n_loci+1]
  const uint8_t* lflank_bytes;
  const uint32_t* rflank_off;   // [n_loci+1]
  const uint8_t* rflank_bytes;
  const uint32_t* allele_off;   // [n_alleles+1]
  const uint8_t* allele_bytes;
  const double* allele_artifact_lp;  // [n_alleles*13] log_prob_pcr_artifact(allele, D), D = -6..6
  double* out_ll;
  uint32_t max_flank, max_block, max_hap;  // shared-memory sizing
};
size_t stutter_block_smem_bytes(uint32_t max_flank, uint32_t max_block, uint32_t max_hap);
cudaError_t launch_stutter(const StutConsts& C, const StutterDevBatch& B, cudaStream_t stream);

// Thresholded unit-cost edit distances and greedy clustering (edit_kernel.cu, edit_core.cuh): one warp per pair.
struct EditArgs {
  const uint8_t* seq_bytes;
  const uint32_t* seq_off;      // [n_seqs+1]
  const uint32_t* pair_a;       // rows    (cent_seq of HaplotypeGenerator::needleman_wunsch)
  const uint32_t* pair_b;       // columns (read_seq)
  const int32_t* pair_T;        // threshold of each pair
  uint32_t n_pairs;             // used when n_pairs_ptr == NULL
  const uint32_t* n_pairs_ptr;  // device word holding the number of pairs
  int32_t* out;                 // [n_pairs]
  uint32_t* cursor;             // pair ticket of the persistent grid (zero before the launch)
  int8_t* lines;                // Myers kernel: line_stride bytes per warp (only read when a string exceeds 1024 bases)
  int32_t* dp_lines;            // exact kernel: line_stride words per warp
  uint32_t line_stride;
  uint32_t* flagged;            // pairs with ED == T, listed by the Myers kernel for the exact kernel (may be NULL)
  uint32_t* n_flagged;
};
uint32_t edit_myers_warps(int sm_count);
uint32_t edit_exact_warps(int sm_count);
cudaError_t launch_edit_myers(const EditArgs& A, uint32_t pair_cap, int sm_count, cudaStream_t stream);
cudaError_t launch_edit_exact(const EditArgs& A, uint32_t pair_cap, int sm_count, cudaStream_t stream);

struct ClusterDev {
  uint32_t n_sets, n_items;
  const uint32_t* set_begin;   // [n_sets+1] items of each set
  const int32_t* set_T;        // [n_sets]
  const uint32_t* item_set;    // [n_items]
  const uint32_t* item_seq;    // [n_items] sequence of each item
  int32_t* best_score;         // [n_items]
  int32_t* centroid_of;        // [n_items] centroid of the item, as an index into its set
  uint32_t* cur_centroid;      // [n_sets] item that is the centroid of the current round
  uint32_t* next_centroid;     // [n_sets]
  int32_t* n_centroids;        // [n_sets]
  uint8_t* state;              // [n_sets] 0 running, 1 finished, 2 more than 15 centroids (greedy_clustering returns false)
  uint32_t* pair_item;         // [n_items] round pair list (pair_a / pair_b / pair_T / score live in EditArgs)
  uint32_t* pair_a;
  uint32_t* pair_b;
  int32_t* pair_T;
  const int32_t* score;
  uint32_t* n_pairs;
  uint32_t* cursor;
  const uint32_t* seq_off;       // statistics only
  unsigned long long* stat;      // [2] comparisons, their n*m cells (may be NULL)
};
cudaError_t launch_cluster(const ClusterDev& C, const EditArgs& A, int sm_count, cudaStream_t stream);

}  // namespace ltr
