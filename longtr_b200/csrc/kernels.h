// kernels.h -- launch interface between the C-ABI implementation (abi.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

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

// Posterior step for a batch of loci (Genotyper::calc_log_sample_posteriors, genotyper.cpp:45-83).
struct DevPosterior {
  uint32_t n_loci;
  const uint32_t* locus_hap_begin;    // H_l
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
};
cudaError_t launch_posteriors(const DevPosterior& P, cudaStream_t stream);

}  // namespace ltr
