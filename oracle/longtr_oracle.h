/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the LongTR hot path.
 *
 * Nothing in the product (longtr_b200/) may include, link or call this; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker or the CPU baseline.
 *
 * Parity status: PINNED.  The reference ships no golden vectors for this path
 * (SURVEY.md section 4), so the restatement is pinned against outputs of the
 * reference itself: oracle/_ref (the unmodified reference sources compiled in
 * place, see build_ref.sh) reproduces SURVEY Appendix A1-A3 bit for bit, and
 * tests/test_oracle_vs_ref.py + tests/golden/ hold this file to oracle/_ref
 * bit-exactly on seeded random loci.
 */
#ifndef LONGTR_ORACLE_H_
#define LONGTR_ORACLE_H_

#include <stdint.h>
#include "longtr_b200_locus.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ltr_oracle_params {
  /* AlignmentModel, src/SeqAlignment/HapAligner.h:16-22; defaults :118 */
  float ins_ins, ins_match, del_del, del_match, match_match, match_ins, match_del;
  int32_t indel_flank_len; /* HapAligner::INDEL_FLANK_LEN (default 5) */
} ltr_oracle_params;

void ltr_oracle_default_params(ltr_oracle_params* p);

/* HapAligner::align_seq_to_hap, src/SeqAlignment/HapAligner.cpp:236-343.
 * full_hap = Haplotype::get_seq() (flanks included), read = trimmed read. */
double ltr_oracle_viterbi_pair(const char* full_hap, int32_t hap_len, const char* read,
                               int32_t read_len, const ltr_oracle_params* p);

/* Same recurrence, also reports the number of DP cells evaluated before the
 * reference would have returned (row bail-out) -- used for GCUPS accounting. */
double ltr_oracle_viterbi_pair_cells(const char* full_hap, int32_t hap_len, const char* read,
                                     int32_t read_len, const ltr_oracle_params* p,
                                     int64_t* cells);

/* HapAligner::trim_alignment, HapAligner.cpp:346-465 (+ the empty-read fallback of
 * process_read, :820-823).  out must hold strlen(seq)+11 bytes. Returns length. */
int32_t ltr_oracle_trim_read(const ltr_flat_locus* L, int32_t read_index, char* out);

/* HapAligner::calc_seed_base, HapAligner.cpp:467-542: seed index, -1 = none, -2 = bad CIGAR. */
int32_t ltr_oracle_seed_base(const ltr_flat_locus* L, int32_t read_index);

/* process_read with short_ == 1 (HapAligner.cpp:855-975) for one read with a valid seed: writes
 * out_row[a] for every allele flagged for realignment (longtr_oracle_short.c).              */
int ltr_oracle_process_read_short(const ltr_flat_locus* L, int32_t read_index, int32_t seed, double* out_row);

/* HapAligner::process_reads, HapAligner.cpp:545-581 + process_read :812-991 (long path and, for
 * period-1 loci with switch_old_align_len != 0, the homopolymer path).  Returns 0 on success.  */
int ltr_oracle_process_reads(const ltr_flat_locus* L, double* out_ll, int32_t* out_seeds);

/* Flattened batch of (trimmed read, full haplotype) loci; same layout as
 * ltr_viterbi_batch in include/longtr_b200.h.  n_threads>1 shards loci over
 * pthreads (the README's "split the BED" parallelisation).                     */
int ltr_oracle_viterbi_batch(uint32_t n_loci, const uint32_t* locus_hap_begin,
                             const uint32_t* locus_read_begin, const uint32_t* hap_off,
                             const uint8_t* hap_bytes, const uint32_t* read_off,
                             const uint8_t* read_bytes, const ltr_oracle_params* p,
                             double* out_ll, int64_t* cells, int n_threads);

/* Genotyper::calc_log_sample_posteriors, src/genotyper.cpp:45-83 (priors :21-43).
 * ll is [n_reads*H] and is clamped IN PLACE to >= -600 like the reference.
 * sample_label[r] in [0,S). post is [S*H*H], totals [S]. Returns sum of totals. */
double ltr_oracle_log_sample_posteriors(int haploid, int32_t n_samples, int32_t n_reads,
                                        int32_t n_alleles, double* ll, const double* log_p1,
                                        const double* log_p2, const int32_t* sample_label,
                                        double* post, double* totals);

/* Genotyper::get_optimal_haplotypes, src/genotyper.cpp:85-100. best is [2*S]. */
void ltr_oracle_optimal_haplotypes(int32_t n_samples, int32_t n_alleles, const double* post,
                                   int32_t* best);

#ifdef __cplusplus
}
#endif
#endif
