/* TEST INFRASTRUCTURE ONLY -- flat description of one IO-less locus run through the reference's
 * SeqStutterGenotyper (ctor -> genotype -> write_vcf_record), SURVEY.md Appendix A4. */
#ifndef LTR_FULL_LOCUS_H_
#define LTR_FULL_LOCUS_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct ltr_full_read {
  int32_t start, stop;   /* 0-based reference start / inclusive stop */
  int32_t rev_strand;
  int32_t sample;        /* reads are sample-major */
  const char* name;
  const char* seq;
  const char* qual;
  const char* aln;       /* sequence with '-' for deleted reference bases */
  const char* cigar;     /* =XID */
  double log_p1, log_p2; /* phasing terms (src/snp_bam_processor.h:16-18) */
} ltr_full_read;
typedef struct ltr_full_locus {
  const char* chrom_name;
  const char* chrom_seq;
  int32_t region_start, region_stop;
  const char* motif;
  const char* region_name;
  int32_t n_samples;
  const char* const* sample_names;
  const int32_t* n_p1s;  /* [n_samples] reads per haplotype shown in PDP */
  const int32_t* n_p2s;
  int32_t n_reads;
  const ltr_full_read* reads;
  double stutter[6];
  const char* stutter_motif;
  int32_t stutter_period;
  int32_t haploid;
  int32_t indel_flank_len;
  int32_t switch_old_align_len;
  int32_t n_aln_params;
  float aln_params[7];
} ltr_full_locus;
/* Returns the length of the VCF record text written to out (0 if genotype() returned false), <0 on error. */
int32_t ltr_ref_full_locus(const ltr_full_locus* L, char* out, int32_t cap);
#ifdef __cplusplus
}
#endif
#endif
