/* longtr_synth.h -- deterministic synthetic TR loci of BASELINE.json configs 3 and 4, emitted as
 * flattened batches.  Benchmark / test data only: built into its own library (libltr_synth.so, plain C++, no CUDA) so
 * that a process which only needs the workload -- the reference arm of bench.py -- maps no product code. */
#ifndef LONGTR_SYNTH_H_
#define LONGTR_SYNTH_H_

#include "longtr_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ltr_synth_batch {
  ltr_viterbi_batch vit;      /* flattened loci                                   */
  ltr_posterior_batch post;   /* 30 sample-reads per locus, one sample            */
  uint32_t n_haps, n_reads, n_sreads;
  uint64_t hap_nbytes, read_nbytes;
} ltr_synth_batch;
/* config 3 = HiFi STRs, 4 = VNTRs (ONT-like), 5 = homopolymers; loci [first_locus,
 * first_locus+n_loci) of the job seeded with base_seed (mt19937_64(base_seed + locus)).      */
int ltr_synth_generate(int config, uint64_t base_seed, uint32_t first_locus, uint32_t n_loci,
                       int n_threads, ltr_synth_batch** out);
void ltr_synth_params(int config, ltr_params* p);
void ltr_synth_free(ltr_synth_batch* b);

/* The same loci in raw form (configs 3 and 4): whole reads with CIGARs, flank blocks and candidate alleles, ready for
 * ltr_genotyper_run (include/longtr_b200.h).  Same seeds and read bases as ltr_synth_generate.                      */
typedef struct ltr_synth_loci {
  ltr_locus_batch batch;
  uint32_t n_alleles, n_reads, n_cigar_ops;
  uint64_t allele_nbytes, read_nbytes;
} ltr_synth_loci;
int ltr_synth_generate_loci(int config, uint64_t base_seed, uint32_t first_locus, uint32_t n_loci, int n_threads,
                            ltr_synth_loci** out);
void ltr_synth_loci_free(ltr_synth_loci* loci);


#ifdef __cplusplus
}
#endif
#endif
