/* longtr_b200_locus.h -- flat (POD, C-ABI) description of one TR locus.
 *
 * This is the language-neutral form of the objects LongTR's hot path is handed:
 *   - the three haplotype blocks that HaplotypeGenerator::fuse_haplotype_blocks
 *     produces (reference: src/SeqAlignment/HaplotypeGenerator.cpp:580-607) and
 *     that `Haplotype(std::vector<HapBlock*>&)` wraps (Haplotype.h:34-50):
 *     left flank HapBlock, RepeatBlock with its candidate alleles, right flank;
 *   - the pooled reads (`Alignment`, src/SeqAlignment/AlignmentData.h:32-60)
 *     that `HapAligner::process_reads` consumes (HapAligner.h:137-138).
 *
 * The same struct is accepted by
 *   - the product (`ltr_process_reads_flat`, include/longtr_b200.h),
 *   - the CPU restatement under oracle/ (test infrastructure), and
 *   - the driver around the unmodified reference sources in oracle/_ref
 *     (test infrastructure),
 * so that parity tests feed all three the very same bytes.
 */
#ifndef LONGTR_B200_LOCUS_H_
#define LONGTR_B200_LOCUS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ltr_flat_read {
  int32_t start;        /* Alignment::get_start(): 0-based reference start          */
  int32_t stop;         /* Alignment::get_stop():  0-based inclusive reference stop */
  const char* seq;      /* NUL-terminated bases                                      */
  const char* qual;     /* NUL-terminated Phred+33 string, same length as seq        */
  const char* cigar;    /* NUL-terminated, ops from "=XIDSHM" e.g. "110=4I90="       */
} ltr_flat_read;

typedef struct ltr_flat_locus {
  /* block 0: HapBlock(repeat_start - strlen(lflank), repeat_start, lflank) */
  const char* lflank;
  /* block 1: RepeatBlock(repeat_start, repeat_end, alleles[0], period, stutter) +
   *          add_alternate(alleles[1..])                                          */
  int32_t repeat_start;
  int32_t repeat_end;
  int32_t period;
  int32_t n_alleles;
  const char* const* alleles;
  /* block 2: HapBlock(repeat_end, repeat_end + strlen(rflank), rflank) */
  const char* rflank;
  /* StutterModel(inframe_geom, inframe_up, inframe_down,
   *              outframe_geom, outframe_up, outframe_down, motif)
   * (src/stutter_model.h:34-60)                                                    */
  double stutter[6];
  const char* motif;
  /* pooled reads */
  int32_t n_reads;
  const ltr_flat_read* reads;
  /* HapAligner ctor arguments (HapAligner.h:94-95) */
  int32_t indel_flank_len;        /* INDEL_FLANK_LEN, default 5                   */
  int32_t switch_old_align_len;   /* SWITCH_OLD_ALIGN_LEN (--stutter-align-len)   */
  int32_t n_aln_params;           /* 0 (use defaults) or 7                        */
  float aln_params[7];            /* ins_ins, ins_match, del_del, del_match,
                                     match_match, match_ins, match_del            */
  /* optional masks; NULL = all true */
  const uint8_t* realign_to_hap;  /* [n_alleles] */
  const uint8_t* realign_read;    /* [n_reads]   */
} ltr_flat_locus;

#ifdef __cplusplus
}
#endif
#endif
