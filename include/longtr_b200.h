/* longtr_b200.h -- C ABI of the B200-native LongTR hot path.
 *
 * Everything the GPU does for LongTR is reached through the entry points below:
 * plain pointers and sizes, no C++/torch types, no exceptions, no exit().
 * One ltr_ctx per (host thread, GPU).  Return value: 0 = LTR_OK, negative = error
 * (ltr_strerror).  There is NO CPU fallback: if no CUDA device is usable,
 * ltr_ctx_create fails with LTR_ERR_NO_DEVICE.
 *
 * What each entry point replaces in the reference (paths relative to the LongTR
 * tree):
 *   ltr_viterbi_ll        the (pooled read x candidate haplotype) loop of
 *                         HapAligner::process_reads / process_read, long path
 *                         (src/SeqAlignment/HapAligner.cpp:545-581, 812-854) with
 *                         align_seq_to_hap (:236-343) as the per-pair kernel;
 *   ltr_posteriors        Genotyper::calc_log_sample_posteriors
 *                         (src/genotyper.cpp:45-83, priors :21-43);
 *   ltr_job_*             the same two steps for a whole batch of loci kept resident
 *                         on the device (what SeqStutterGenotyper::genotype does per
 *                         locus at src/seq_stutter_genotyper.cpp:634-635, for many
 *                         loci at once);
 *   ltr_process_reads_flat  HapAligner::process_reads on one flat locus, through the
 *                         host-side mirror of the HapAligner class
 *                         (longtr_b200/csrc/host/), used by bindings and tests;
 *   ltr_process_reads_flat_batch  the same for many loci in one GPU job.
 */
#ifndef LONGTR_B200_H_
#define LONGTR_B200_H_

#include <stddef.h>
#include <stdint.h>

#include "longtr_b200_locus.h"

#ifdef __cplusplus
extern "C" {
#endif

#define LTR_OK 0
#define LTR_ERR_NO_DEVICE (-1)
#define LTR_ERR_CUDA (-2)
#define LTR_ERR_INVALID (-3)
#define LTR_ERR_OOM (-4)
#define LTR_ERR_UNSUPPORTED (-5)

typedef struct ltr_ctx ltr_ctx;
typedef struct ltr_job ltr_job;

/* AlignmentModel (HapAligner.h:12-37) + the constants align_seq_to_hap hard-codes. */
typedef struct ltr_params {
  float ins_ins, ins_match, del_del, del_match, match_match, match_ins, match_del;
  int32_t indel_flank_len; /* INDEL_FLANK_LEN; haplotypes are cut by 35-indel_flank_len
                              on each side (HapAligner.cpp:245-246).  5 .. 35; values below
                              5 (where the reference's substr length underflows for
                              haplotypes of 61 .. 2*(35-flank) bases) are answered with
                              LTR_ERR_UNSUPPORTED                                     */
} ltr_params;

/* Fills the Dindel defaults of HapAligner.h:118 and indel_flank_len = 5. */
void ltr_params_default(ltr_params* p);

/* A batch of loci, flattened.  Locus l owns haplotypes [locus_hap_begin[l],
 * locus_hap_begin[l+1]) and pooled reads [locus_read_begin[l], locus_read_begin[l+1]).
 * hap_bytes holds Haplotype::get_seq() of every candidate haplotype (flanks included,
 * column order = gray-code order of Haplotype::next(), Haplotype.cpp:157-196);
 * read_bytes holds the reads already trimmed by HapAligner::trim_alignment
 * (HapAligner.cpp:346-465).  Offsets are byte offsets, arrays have n+1 entries.
 * The log-likelihood of (read p, haplotype h) of locus l is written to
 *   out_ll[ ll_off(l) + (p - locus_read_begin[l]) * H_l + (h - locus_hap_begin[l]) ],
 * ll_off(l) = sum_{l'<l} P_l' * H_l'   -- i.e. the reference's aln_probs[read*H + hap]
 * matrices (HapAligner.cpp:549), one after the other.                              */
typedef struct ltr_viterbi_batch {
  uint32_t n_loci;
  const uint32_t* locus_hap_begin;  /* [n_loci+1] */
  const uint32_t* locus_read_begin; /* [n_loci+1] */
  const uint32_t* hap_off;          /* [n_haps+1]  */
  const uint8_t* hap_bytes;
  const uint32_t* read_off;         /* [n_reads+1] */
  const uint8_t* read_bytes;
} ltr_viterbi_batch;

/* Per-read inputs of the posterior step for the same batch (optional, see ltr_job_create).
 * Sample-reads of locus l are [locus_sread_begin[l], locus_sread_begin[l+1]); each maps to
 * a pooled read of its locus (ReadPooler, src/read_pooler.cpp:3-20) through pool_index
 * (index RELATIVE to the locus' first pooled read), carries the phasing terms log_p1/log_p2
 * (src/snp_bam_processor.h:16-18) and the index of its sample within the locus.       */
typedef struct ltr_posterior_batch {
  const uint32_t* locus_sread_begin; /* [n_loci+1] */
  const uint32_t* pool_index;        /* [n_sreads]  */
  const int32_t* sample_label;       /* [n_sreads], 0..n_samples(l)-1 */
  const double* log_p1;              /* [n_sreads]  */
  const double* log_p2;              /* [n_sreads]  */
  const uint32_t* locus_n_samples;   /* [n_loci]    */
  const uint8_t* locus_haploid;      /* [n_loci] or NULL (all diploid) */
  /* -- optional (zero / NULL = off), appended in version 0.2 -- */
  const uint8_t* second_mate;        /* [n_sreads] or NULL: sample-read r is the second mate of read r-1 (same read name,
                                        src/seq_stutter_genotyper.cpp:494): the LL rows of the two are summed before the
                                        posteriors are formed (:546-559)                                              */
  const uint8_t* read_aligned;       /* [n_sreads] or NULL (all): seed_positions >= 0, i.e. the read lets its sample vote
                                        when uncalled alleles are removed (get_unused_alleles, :262-265)              */
  int32_t prune_uncalled;            /* != 0: after the posteriors, non-reference alleles that are in no voting sample's
                                        optimal pair are dropped and the posteriors recomputed on the K surviving alleles
                                        (SeqStutterGenotyper::genotype, :636-645).  post of locus l is then [S][K][K],
                                        compact, at the locus' usual offset; the surviving alleles come back through
                                        ltr_job_outputs.kept_mask                                                      */
} ltr_posterior_batch;

typedef struct ltr_job_stats {
  uint64_t n_pairs;        /* (read, haplotype) pairs                               */
  uint64_t n_cells;        /* sum over pairs of n*m (SURVEY 8d GCUPS definition)    */
  uint64_t n_fallback;     /* pairs re-run by the exact row-bail-out kernel          */
  uint64_t h2d_bytes, d2h_bytes;
  uint32_t n_launches;     /* kernels launched by the last ltr_job_run               */
  float kernel_ms;         /* device time of the last ltr_job_run (CUDA events)      */
  float viterbi_ms;        /* ... of which: Viterbi kernels                          */
  uint64_t n_pairs_computed; /* pairs actually aligned: identical trimmed reads of a locus are aligned
                                once and fanned out (results are a pure function of the two strings) */
  uint64_t n_cells_computed; /* ... and the cells evaluated for them: n*m over the full matrix, the cells
                                inside the band for pairs certified by the banded kernel                 */
  uint64_t n_band_pairs;       /* pairs sent to the banded kernel (band_core.cuh) ...                    */
  uint64_t n_band_uncertified; /* ... of which the band could not certify (re-run over the full matrix)  */
  float plan_ms;               /* device time of the plan kernels (de-duplication of the trimmed reads, band classes,
                                  task lists: plan_kernels.cu) inside kernel_ms; 0 when the plan was built on the host */
  uint64_t n_band_retried;     /* of n_band_uncertified: pairs whose banded score proved that a wider band class certifies
                                  them and that ran a second time there instead of over the full matrix             */
} ltr_job_stats;

/* ---- context --------------------------------------------------------------------- */
int ltr_ctx_create(int device, ltr_ctx** out);
void ltr_ctx_destroy(ltr_ctx* ctx);
/* Banded evaluation of the Viterbi score (exact: a pair is accepted only when its banded score proves the band was
 * wide enough, otherwise it is re-run over the full matrix).  half_width < 0: off; 0 (default): automatic margin;
 * > 0: minimum number of diagonals kept on either side of the diagonals 0 .. m-n.  Results do not depend on it. */
int ltr_ctx_set_band(ltr_ctx* ctx, int32_t half_width);
/* Where the plan of a job (de-duplication of the trimmed reads of each locus, band classes, task lists) is built:
 * 0 (default) on the device for batches of >= 4096 pooled reads and on the host below that (per-locus calls: fewer
 * launches), 1 always on the host, 2 always on the device.  Results do not depend on it. */
int ltr_ctx_set_plan(ltr_ctx* ctx, int32_t mode);
/* How read_bytes of the batches submitted afterwards is encoded (ltr_job_create / ltr_job_submit*): 0 = one byte per base
 * (default), 1 = ONE 4-bit stream over all reads, BAM's own sequence encoding ("=ACMGRSVTWYHKDBN"): base b of the batch sits
 * in byte b / 2, even b in the high nibble; read_off keeps counting bases.  Halves the largest host-to-device copy; the
 * plan's first kernel expands the stream on the device.  Results do not depend on it.                                     */
int ltr_ctx_set_read_encoding(ltr_ctx* ctx, int32_t encoding);
const char* ltr_strerror(int code);
const char* ltr_last_error(const ltr_ctx* ctx); /* CUDA error text of the last failure */
const char* ltr_version(void);

/* ---- one-shot calls (host buffers in, host buffers out) -------------------------- */
/* out_ll must hold sum_l P_l*H_l doubles.  Sentinels as in the reference: -1e9 when
 * the haplotype is <= 60 bp (HapAligner.cpp:241-244), -700 for |n-m| > 600 (:249-252)
 * and for the per-row bail-out (:300-306).                                            */
int ltr_viterbi_ll(ltr_ctx* ctx, const ltr_params* params, const ltr_viterbi_batch* batch,
                   double* out_ll, ltr_job_stats* stats);

/* Genotyper::calc_log_sample_posteriors for one locus.  ll is [n_reads*n_alleles] and is
 * clamped IN PLACE to >= -600 like the reference (genotyper.cpp:57-58); post receives
 * [n_samples*H*H] log posteriors, totals [n_samples]; *total_ll their sum.            */
int ltr_posteriors(ltr_ctx* ctx, int haploid, int32_t n_samples, int32_t n_reads,
                   int32_t n_alleles, double* ll, const double* log_p1, const double* log_p2,
                   const int32_t* sample_label, double* post, double* totals, double* total_ll);

/* ---- homopolymer / --stutter-align-len path ------------------------------------------- */
/* What HapAligner::process_reads does when block 1 has period 1 and SWITCH_OLD_ALIGN_LEN != 0
 * (HapAligner.cpp:552, 567-579, 855-975): per (pooled read, haplotype) two quality-aware flank
 * alignments around a seed base (align_seq_to_hap_short, :27-163) with the repeat block marginalised over
 * PCR stutter artifacts of -6..+6 bases (StutterAlignerClass), joined by compute_aln_logprob (:165-233).
 * Reads are the WHOLE pooled reads (not trimmed) with their Phred+33 qualities and the seed index of
 * HapAligner::calc_seed_base (ltr_seed_base_flat); read_seed < 0 gives the reference's all-zero row.
 * Loci have one left flank, one right flank and their repeat-block alleles (all non-empty); the stutter
 * model of locus l is stutter[6*l..6*l+5] = (inframe_geom, inframe_up, inframe_down, outframe_geom,
 * outframe_up, outframe_down) with motif length motif_len[l] (StutterModel, src/stutter_model.h:34-60).
 * Output layout as ltr_viterbi_batch: out_ll[ll_off(l) + p*H_l + h].                                   */
typedef struct ltr_stutter_batch {
  uint32_t n_loci;
  const uint32_t* locus_allele_begin; /* [n_loci+1] */
  const uint32_t* locus_read_begin;   /* [n_loci+1] */
  const uint32_t* lflank_off;         /* [n_loci+1] */
  const uint8_t* lflank_bytes;
  const uint32_t* rflank_off;         /* [n_loci+1] */
  const uint8_t* rflank_bytes;
  const uint32_t* allele_off;         /* [n_alleles+1] */
  const uint8_t* allele_bytes;
  const double* stutter;              /* [6*n_loci] */
  const int32_t* motif_len;           /* [n_loci]   */
  const uint32_t* read_off;           /* [n_reads+1] */
  const uint8_t* read_bytes;
  const uint8_t* qual_bytes;
  const int32_t* read_seed;           /* [n_reads]  */
  const uint8_t* realign_allele;      /* [n_alleles] or NULL (all): 0 leaves the column untouched */
  const uint8_t* realign_read;        /* [n_reads]   or NULL (all): 0 leaves the row untouched    */
} ltr_stutter_batch;
/* params->indel_flank_len is ignored on this path.  stats->n_cells counts cell-equivalents as defined in
 * SURVEY.md section 8d (flank rows x columns + 13 x block length per column, both flanks).            */
int ltr_stutter_ll(ltr_ctx* ctx, const ltr_params* params, const ltr_stutter_batch* batch, double* out_ll,
                   ltr_job_stats* stats);
/* The same, but a locus that cannot be processed fails ALONE: locus_status[l] (non-NULL, [n_loci]) receives LTR_OK or the
 * locus' error -- LTR_ERR_UNSUPPORTED for an empty (<DEL>) allele (in the reference the stutter row of an empty block
 * overwrites the row before it, HapAligner.cpp:64-111 with bn = 0) or for read flanks / alleles too long for the kernel's
 * shared-memory staging, LTR_ERR_INVALID for a malformed stutter model, flank or seed -- its rows of out_ll stay untouched,
 * and every other locus is computed.  ltr_stutter_ll is this call with locus_status = NULL: the first such locus fails
 * the whole call and out_ll is not modified.                                                                          */
int ltr_stutter_ll_status(ltr_ctx* ctx, const ltr_params* params, const ltr_stutter_batch* batch, double* out_ll,
                          int32_t* locus_status, ltr_job_stats* stats);

/* ---- resident jobs: upload once, run many times, download ------------------------- */
/* post may be NULL (Viterbi only).  Host arrays may be released after the call.       */
int ltr_job_create(ltr_ctx* ctx, const ltr_params* params, const ltr_viterbi_batch* batch,
                   const ltr_posterior_batch* post, ltr_job** out);
/* Launches every kernel of the job on the context's streams; returns after completion. */
int ltr_job_run(ltr_ctx* ctx, ltr_job* job);
/* Sizes of the result arrays (in elements). */
void ltr_job_sizes(const ltr_job* job, uint64_t* n_ll, uint64_t* n_post, uint64_t* n_totals);
/* Any output pointer may be NULL. out_post is [sum_l S_l*H_l*H_l], out_totals [sum_l S_l]. */
int ltr_job_download(ltr_ctx* ctx, ltr_job* job, double* out_ll, double* out_post,
                     double* out_totals);
/* kept_mask [n_haps] of a job created with post->prune_uncalled (all ones otherwise). */
int ltr_job_download_kept(ltr_ctx* ctx, ltr_job* job, uint8_t* out_kept_mask);
void ltr_job_get_stats(const ltr_job* job, ltr_job_stats* stats);
void ltr_job_destroy(ltr_ctx* ctx, ltr_job* job);
/* The posterior stage alone for many loci whose LL matrices the caller already holds (e.g. the output of ltr_stutter_ll):
 * ll = the pooled aln_probs matrices back to back (P_l x H_l, layout of ltr_viterbi_batch), P_l = locus_read_begin[l+1] -
 * locus_read_begin[l], H_l likewise; post as for ltr_job_create, including mate pairs and the removal of uncalled
 * alleles.  out_kept_mask may be NULL.  Genotyper::calc_log_sample_posteriors (src/genotyper.cpp:45-83) per locus.  */
int ltr_posteriors_batch(ltr_ctx* ctx, uint32_t n_loci, const uint32_t* locus_hap_begin, const uint32_t* locus_read_begin,
                         const double* ll, const ltr_posterior_batch* post, double* out_post, double* out_totals,
                         uint8_t* out_kept_mask);

/* ---- asynchronous jobs: many batches in flight from ONE host thread ------------------------------------------------ */
/* ltr_job_submit enqueues a whole job -- upload of the batch, plan, kernels, posteriors, download of the results into
 * out_ll / out_post / out_totals (any may be NULL) -- on the streams of the context and returns without waiting; jobs
 * submitted one after the other overlap (the upload of job k+1 and the download of job k-1 run beside the kernels of
 * job k).  The caller's input AND output arrays must stay valid and untouched until ltr_job_wait returns; page-locked
 * (pinned) host memory makes the copies truly asynchronous, pageable memory works but blocks inside submit.
 * ltr_job_wait blocks until the job is complete and returns its status: a malformed batch that is only detected on the
 * device (read offsets, sample-read indices) is reported here as LTR_ERR_INVALID.  ltr_job_poll: 1 = complete,
 * 0 = still running, negative = error.  Statistics are valid after ltr_job_wait; finish with ltr_job_destroy.
 * (the submit / wait pair the pipelined host of SeqStutterGenotyper::genotype calls at
 * src/seq_stutter_genotyper.cpp:634-635.)                                                                            */
int ltr_job_submit(ltr_ctx* ctx, const ltr_params* params, const ltr_viterbi_batch* batch,
                   const ltr_posterior_batch* post, double* out_ll, double* out_post, double* out_totals,
                   ltr_job** out);
/* The same with every output named in one struct (any may be NULL); kept_mask [n_haps] receives 1 for every candidate
 * allele that survives post->prune_uncalled (all ones when it is off). */
typedef struct ltr_job_outputs {
  double* ll;          /* [sum_l P_l*H_l]   */
  double* post;        /* [sum_l S_l*H_l*H_l] (locus l uses the first S_l*K_l*K_l entries of its slice when pruned) */
  double* totals;      /* [sum_l S_l]       */
  uint8_t* kept_mask;  /* [n_haps]          */
} ltr_job_outputs;
int ltr_job_submit_outputs(ltr_ctx* ctx, const ltr_params* params, const ltr_viterbi_batch* batch,
                           const ltr_posterior_batch* post, const ltr_job_outputs* outputs, ltr_job** out);
int ltr_job_wait(ltr_ctx* ctx, ltr_job* job);
int ltr_job_poll(ltr_ctx* ctx, ltr_job* job);

/* ---- reference-facing convenience (host mirror of HapAligner) --------------------- */
/* HapAligner(haplotype, realign_to_hap, INDEL_FLANK_LEN, SWITCH_OLD_ALIGN_LEN, params)
 *   .process_reads(alns, 0, &base_quality, realign_read, out_ll, out_seeds)
 * (HapAligner.h:94-95, 137-138) on a flat locus.  out_ll is [n_reads*n_alleles]; slots of
 * reads / haplotypes that are not realigned are left untouched, as in the reference.   */
int ltr_process_reads_flat(ltr_ctx* ctx, const ltr_flat_locus* locus, double* out_ll,
                           int32_t* out_seeds);
/* The same for n_loci flat loci in one call: what a host that keeps many regions open (the pipelined replacement of
 * the serial loop in src/genotyper_bam_processor.cpp:227-351) hands over.  The long-path loci
 * that share their alignment parameters become ONE flattened GPU job (one plan, one upload, one set of launches);
 * loci on the homopolymer path (HapAligner.cpp:552) are processed one by one.  out_ll[l] / out_seeds[l] are locus l's
 * arrays as in ltr_process_reads_flat; results are identical to n_loci separate calls.                              */
int ltr_process_reads_flat_batch(ltr_ctx* ctx, int32_t n_loci, const ltr_flat_locus* loci, double* const* out_ll,
                                 int32_t* const* out_seeds);

/* The flattening half of the batch call on its own (host only, no GPU): the haplotypes (column order, only those flagged
 * for realignment) and trimmed reads (HapAligner::trim_alignment, only those flagged) of n_loci long-path flat loci as
 * the arrays of an ltr_viterbi_batch, ready for ltr_job_create together with an ltr_posterior_batch of the caller.
 * hap_col[h] / read_row[r] give the column / row of flattened haplotype h / read r in its locus' aln_probs matrix;
 * params_out (optional) receives the alignment parameters of the loci (they must all agree).  Loci on the homopolymer
 * path give LTR_ERR_UNSUPPORTED (ltr_stutter_ll has its own batch form).  Free with ltr_flat_batch_free.            */
typedef struct ltr_flat_batch {
  ltr_viterbi_batch vit;
  const int32_t* hap_col;   /* [n_haps]  */
  const int32_t* read_row;  /* [n_reads] */
  uint32_t n_haps, n_reads;
} ltr_flat_batch;
int ltr_flatten_loci(int32_t n_loci, const ltr_flat_locus* loci, ltr_params* params_out, ltr_flat_batch** out);
void ltr_flat_batch_free(ltr_flat_batch* batch);

/* ---- many loci in flight for a host that produces loci one at a time ------------------------------------------ */
/* LongTR's region loop (src/genotyper_bam_processor.cpp:227-351) decodes a region, builds its haplotypes and calls
 * HapAligner::process_reads before it looks at the next one.  The pipeline lets that loop hand each locus over and carry
 * on: ltr_pipeline_submit deep-copies the flat locus (the caller may release it at once); loci are collected into
 * batches of batch_loci and run through ltr_process_reads_flat_batch on `slots` worker threads, each with its own
 * ltr_ctx on `device`, so that decoding the next regions overlaps the plan / upload / kernels of the previous ones.
 * Results come back by tag in completion order:
 *   ltr_pipeline_next returns 1 and the locus' aln_probs (n_reads x n_alleles, slots that are not realigned hold `fill`),
 *   seed positions and status (LTR_OK or the error of that locus); the pointers stay valid until the next call of
 *   ltr_pipeline_next / ltr_pipeline_destroy.  It returns 0 when nothing is finished (with wait != 0: when nothing is
 *   pending either; waiting also sends a partially filled batch on its way).
 * submit blocks while 2*slots batches are queued (back-pressure).  One producer / consumer thread at a time per call
 * family is assumed (the functions are internally locked).  No CPU fallback: ltr_pipeline_create fails without a device. */
typedef struct ltr_pipeline ltr_pipeline;
int ltr_pipeline_create(int device, int32_t batch_loci, int32_t slots, ltr_pipeline** out);
int ltr_pipeline_submit(ltr_pipeline* p, const ltr_flat_locus* locus, uint64_t tag, const int32_t* in_seeds, double fill);
int ltr_pipeline_flush(ltr_pipeline* p);
int ltr_pipeline_next(ltr_pipeline* p, int wait, uint64_t* tag, int32_t* n_reads, int32_t* n_alleles, const double** ll,
                      const int32_t** seeds, int* status);
void ltr_pipeline_destroy(ltr_pipeline* p);

/* ---- many raw loci -> genotype calls: the front half of SeqStutterGenotyper::genotype, batched -------------------------- */
/* What LongTR does per locus between "the reads of the region are decoded" and "the record can be written"
 * (src/seq_stutter_genotyper.cpp:485-497, 599-645): pool the reads by sequence (ReadPooler, src/read_pooler.cpp:3-20), trim
 * every pool to the repeat +- INDEL_FLANK_LEN with its CIGAR (HapAligner::trim_alignment, HapAligner.cpp:346-465), align
 * the pools to every candidate haplotype (GPU), scatter the pool rows to the reads, form the posteriors (GPU), drop the
 * uncalled alleles and recompute (GPU), extract GT / Q / PQ / GL / PL (Genotyper::extract_genotypes_and_likelihoods,
 * src/genotyper.cpp:132-256) -- for a whole batch of loci in structure-of-arrays form, with no per-read objects.
 * Candidate alleles come from the caller (HaplotypeGenerator, out of scope); loci are independent.
 * Reads of a locus are SAMPLE-MAJOR (all reads of sample 0, then sample 1, ...), as in the reference.
 * CIGARs are BAM-encoded (length << 4 | op, op in MIDNSHP=X = 0..8; N and P are rejected like the reference does).   */
typedef struct ltr_locus_batch {
  uint32_t n_loci;
  const uint32_t* lflank_off;          /* [n_loci+1] block 0 of the haplotype (left flank, 35 bp in LongTR)            */
  const uint8_t* lflank_bytes;
  const uint32_t* rflank_off;          /* [n_loci+1] block 2                                                             */
  const uint8_t* rflank_bytes;
  const uint32_t* locus_allele_begin;  /* [n_loci+1] block 1: candidate alleles, reference allele first                  */
  const uint32_t* allele_off;          /* [n_alleles+1]                                                                  */
  const uint8_t* allele_bytes;
  const int32_t* repeat_start;         /* [n_loci] reference coordinates of block 1 (RepeatBlock start / end)            */
  const int32_t* repeat_end;
  const uint32_t* locus_read_begin;    /* [n_loci+1] raw reads                                                           */
  const int32_t* read_start;           /* [n_reads] Alignment::get_start(): 0-based reference start                      */
  const int32_t* read_stop;            /* [n_reads] Alignment::get_stop(): 0-based inclusive reference stop              */
  const uint32_t* read_off;            /* [n_reads+1]                                                                    */
  const uint8_t* read_bytes;
  const uint32_t* cigar_off;           /* [n_reads+1] offsets into cigar_ops                                             */
  const uint32_t* cigar_ops;
  const int32_t* read_sample;          /* [n_reads] 0 .. locus_n_samples[l]-1, non-decreasing inside a locus             */
  const double* log_p1;                /* [n_reads] phasing terms (src/snp_bam_processor.h:16-18)                        */
  const double* log_p2;
  const uint8_t* second_mate;          /* [n_reads] or NULL                                                              */
  const uint32_t* locus_n_samples;     /* [n_loci]                                                                       */
  const uint8_t* locus_haploid;        /* [n_loci] or NULL                                                               */
} ltr_locus_batch;

/* Calls of a batch, structure of arrays, owned by the library (ltr_batch_calls_free).  Sample (l, s) has the global
 * index locus_sample_begin[l] + s.  Allele indices are indices into the CALLER's allele list of the locus.            */
typedef struct ltr_batch_calls {
  uint32_t n_loci;
  const int32_t* status;                   /* [n_loci] LTR_OK, or why this locus (alone) was not genotyped               */
  const uint32_t* locus_sample_begin;      /* [n_loci+1]                                                                 */
  const uint32_t* locus_allele_begin;      /* [n_loci+1]                                                                 */
  const uint8_t* kept_mask;                /* [n_alleles] allele survives the removal of uncalled alleles                */
  const int32_t* n_kept;                   /* [n_loci]                                                                   */
  const int32_t* n_pools;                  /* [n_loci] distinct read sequences (ReadPooler::num_pools)                   */
  const int32_t* gts;                      /* [2 * n_samples] optimal allele pair (GT)                                   */
  const double* log_phased_posteriors;     /* [n_samples] -> PQ = exp(.)                                                 */
  const double* log_unphased_posteriors;   /* [n_samples] -> Q = exp(.)                                                  */
  const double* gl_diffs;                  /* [n_samples] GLDIFF                                                         */
  const double* sample_total_lls;          /* [n_samples]                                                                */
  const int32_t* n_reads;                  /* [n_samples] DP                                                             */
  const uint64_t* gl_begin;                /* [n_samples+1] slice of gls / pls; K(K+1)/2 (haploid: K) entries are used   */
  const double* gls;                       /* log10 genotype likelihoods over the KEPT alleles (GL)                      */
  const int32_t* pls;                      /* PL                                                                         */
  double prep_ms, gpu_wait_ms, post_ms, total_ms;  /* host timing of the call                                           */
  double submit_ms;                        /* of which inside ltr_job_submit_outputs                                     */
  uint32_t n_chunks;                       /* jobs the batch was cut into                                                */
  const int32_t* read_allele;              /* [n_reads] or NULL (ltr_genotyper_set_read_alleles): the allele of its sample's
                                              genotype each read supports -- what MALLREADS counts                         */
  const uint64_t* pgl_begin;               /* [n_samples+1] or NULL (ltr_genotyper_set_phased_gls): slice of phased_gls;
                                              K*K (haploid: K) entries are used, [a*K + b] over the KEPT alleles          */
  const double* phased_gls;                /* log10 likelihoods of the phased genotypes (PHASEDGL)                       */
} ltr_batch_calls;

typedef struct ltr_genotyper ltr_genotyper;
/* devices: CUDA device indices (one context each; loci are sharded over them in chunks of chunk_loci, results come back in
 * input order -- the host-side gather of SURVEY.md section 8e); host_threads <= 0: all hardware threads; chunk_loci <= 0:
 * default (20 000).  No CPU fallback: fails with LTR_ERR_NO_DEVICE when a device cannot be used.                       */
int ltr_genotyper_create(const int32_t* devices, int32_t n_devices, int32_t host_threads, int32_t chunk_loci,
                         ltr_genotyper** out);
void ltr_genotyper_destroy(ltr_genotyper* g);
/* on != 0: the calls of later runs carry read_allele (the LL matrices are then downloaded as well: ~8 bytes per pooled read and
 * allele).  As write_vcf_record assigns reads (src/seq_stutter_genotyper.cpp:954-970): the first allele of the genotype unless
 * log_p2 + LL[second] >= log_p1 + LL[first].                                                                              */
int ltr_genotyper_set_read_alleles(ltr_genotyper* g, int32_t on);
/* on != 0: later runs also return phased_gls / pgl_begin (the PHASEDGL field behind --output-phased-gls). */
int ltr_genotyper_set_phased_gls(ltr_genotyper* g, int32_t on);
/* Genotypes every locus of the batch.  A malformed locus (CIGAR the reference dies on, reads that do not fit their CIGAR,
 * missing flanks) fails alone: its status is the error, the rest of the batch is unaffected.                          */
int ltr_genotyper_run(ltr_genotyper* g, const ltr_params* params, const ltr_locus_batch* batch, ltr_batch_calls** out);
void ltr_batch_calls_free(ltr_batch_calls* calls);
/* Host only: HapAligner::trim_alignment on read r of the batch (its locus is looked up); writes the trimmed read (or the
 * 10 bp pseudo read of HapAligner.cpp:820-823 when nothing is left) into out[cap] and returns its length, negative =
 * error.  What the genotyper aligns; exposed for parity tests.                                                        */
int32_t ltr_locus_batch_trim_read(const ltr_locus_batch* batch, const ltr_params* params, uint32_t locus, uint32_t read,
                                  uint8_t* out, int32_t cap);

/* Genotype calls of one locus from its read x haplotype LL matrix: what SeqStutterGenotyper::genotype +
 * write_vcf_record obtain from Genotyper::calc_log_sample_posteriors (GPU) followed by
 * Genotyper::extract_genotypes_and_likelihoods (src/genotyper.cpp:132-256; host, integer / small vectors)
 * with hap_to_allele = identity (one multi-allele block).  Reads are sample-major: sample s owns
 * reads_per_sample[s] consecutive rows of ll / log_p1 / log_p2.  ll is clamped in place (genotyper.cpp:57-58).
 * Output arrays may be NULL.  n_gl = H(H+1)/2 (haploid: H), n_pgl = H*H (haploid: H).                      */
typedef struct ltr_locus_calls {
  int32_t* best_gts;                   /* [2*S]  allele pair per sample (GT)                 */
  double* log_phased_posteriors;       /* [S]    -> PQ = exp(.)                              */
  double* log_unphased_posteriors;     /* [S]    -> Q  = exp(.)                              */
  double* hap_log_phased_posteriors;   /* [S]                                                */
  double* hap_log_unphased_posteriors; /* [S]                                                */
  double* gls;                         /* [S*n_gl]  log10 genotype likelihoods (GL)          */
  int32_t* pls;                        /* [S*n_gl]  PL                                       */
  double* phased_gls;                  /* [S*n_pgl] PHASEDGL                                 */
  double* gl_diffs;                    /* [S]    GLDIFF                                      */
  double* sample_total_lls;            /* [S]                                                */
  double* log_sample_posteriors;       /* [S*H*H]                                            */
  double total_ll;
} ltr_locus_calls;
int ltr_genotype_locus(ltr_ctx* ctx, int haploid, int32_t n_samples, const int32_t* reads_per_sample,
                       int32_t n_alleles, double* ll, const double* log_p1, const double* log_p2,
                       ltr_locus_calls* out);
/* The same followed by what SeqStutterGenotyper::genotype does next (src/seq_stutter_genotyper.cpp:636-645):
 * non-reference alleles that are in no sample's optimal haplotype pair are dropped (get_unused_alleles, :250-311;
 * samples whose reads all have seed_positions < 0 do not vote; seed_positions may be NULL = all aligned), the LL
 * columns of the kept alleles are carried over (:317-409) and the posteriors are recomputed on them.
 * kept_alleles[0..*n_kept) receives the original indices of the kept alleles (always includes 0); the arrays of
 * `out` are filled for H' = *n_kept alleles (allocate them for H).                                       */
int ltr_genotype_locus_pruned(ltr_ctx* ctx, int haploid, int32_t n_samples, const int32_t* reads_per_sample,
                              int32_t n_alleles, double* ll, const double* log_p1, const double* log_p2,
                              const int32_t* seed_positions, int32_t* kept_alleles, int32_t* n_kept,
                              ltr_locus_calls* out);
/* The host half of the above on posteriors that are already computed (no GPU involved). */
int ltr_extract_calls(int haploid, int32_t n_samples, int32_t n_alleles, const double* post, const double* totals,
                      ltr_locus_calls* out);

/* Host-only pieces of HapAligner on a flat locus (integer work, no GPU involved):
 * ltr_trim_read_flat  = HapAligner::trim_alignment (HapAligner.cpp:346-465): writes the NUL-terminated trimmed
 *                       read into out[cap] and returns its length (0 = the caller substitutes the 10 bp
 *                       pseudo read of HapAligner.cpp:820-823), negative = error;
 * ltr_seed_base_flat  = HapAligner::calc_seed_base (HapAligner.cpp:493-542): seed index or -1 (none);
 *                       LTR_ERR_INVALID for a CIGAR operation the reference dies on.                       */
int32_t ltr_trim_read_flat(const ltr_flat_locus* locus, int32_t read_index, char* out, int32_t cap);
/* ltr_pool_reads      = ReadPooler::add_alignment for every read + ReadPooler::pool (src/read_pooler.cpp:3-20,
 *                       src/read_pooler.h:42-48): pool_index[r] = pool of read r (reads with identical sequence share a
 *                       pool, pools are numbered by first appearance), *n_pools, and the pooled reads' qualities -- the
 *                       per-position upper median of BaseQuality::median_base_qualities (src/base_quality.cpp:11-28) --
 *                       back to back in pool order in pooled_quals[cap] (may be NULL).  Returns the bytes of pooled
 *                       qualities, negative = error.                                                                */
int64_t ltr_pool_reads(int32_t n_reads, const char* const* seqs, const char* const* quals, int32_t* pool_index,
                       int32_t* n_pools, char* pooled_quals, int64_t cap);
int32_t ltr_seed_base_flat(const ltr_flat_locus* locus, int32_t read_index);

/* ---- candidate-haplotype clustering ------------------------------------------------
 * ltr_edit_distances  HaplotypeGenerator::needleman_wunsch (src/SeqAlignment/HaplotypeGenerator.cpp:201-235) for
 *                     many pairs at once: out_score[p] is what that function leaves in `score` for
 *                     cent_seq = sequence pair_a[p], read_seq = sequence pair_b[p], T = pair_T[p] -- the unit-cost
 *                     edit distance when it is below T, T + 1 when it is above or the lengths differ by more than T,
 *                     and T or T + 1 as the reference's row test decides when it equals T.  Sequence s is
 *                     seq_bytes[seq_off[s] .. seq_off[s+1]); bytes are compared as bytes.  0 <= T <= 999.
 * ltr_cluster_greedy  HaplotypeGenerator::greedy_clustering (:238-271) for many sets at once.  Set k is the item list
 *                     set_items[set_begin[k] .. set_begin[k+1]) of sequence indices (the reference's `seqs`, in its
 *                     order; the same sequences may appear in several sets, e.g. one set per threshold of :403) and
 *                     set_T[k] its threshold.  out_centroid_of[i] = position, inside its set, of the centroid the
 *                     item was assigned to (centroids name themselves): clusters[seqs[c]] of the reference is the
 *                     items with out_centroid_of == c, in item order.  out_n_centroids[k]; out_ok[k] = 0 when the
 *                     reference returns false (more than 15 centroids), in which case the set's assignments are
 *                     unspecified.                                                                                  */
int ltr_edit_distances(ltr_ctx* ctx, const uint8_t* seq_bytes, const uint32_t* seq_off, uint32_t n_seqs,
                       const uint32_t* pair_a, const uint32_t* pair_b, const int32_t* pair_T, uint32_t n_pairs,
                       int32_t* out_score, ltr_job_stats* stats);
int ltr_cluster_greedy(ltr_ctx* ctx, const uint8_t* seq_bytes, const uint32_t* seq_off, uint32_t n_seqs,
                       const uint32_t* set_begin, const uint32_t* set_items, const int32_t* set_T, uint32_t n_sets,
                       int32_t* out_centroid_of, int32_t* out_n_centroids, uint8_t* out_ok, ltr_job_stats* stats);

/* ---- alignment input ------------------------------------------------------------------
 * A from-scratch BGZF / BAM / BAI reader (zlib only) for what LongTR reads through htslib (src/bam_io.h:313-420
 * BamCramReader; src/bam_io.cpp:94-214 sam_open / sam_hdr_read / sam_index_load / sam_itr_querys / sam_itr_next).
 * ltr_bam_open     maps the file, parses header and reference dictionary and loads index_path (NULL: "<path>.bai" when it
 *                  exists; without an index region queries scan the file).
 * ltr_bam_fetch    the records that overlap [beg, end) (0-based, half open) of reference tid, in file order, as a structure
 *                  of arrays owned by the library; tid < 0: every record.  keep_raw != 0 also keeps each record's BAM bytes
 *                  (what follows block_size: the layout of htslib's bam1_t core + data).  Thread safe on one ltr_bam.  */
typedef struct ltr_bam ltr_bam;
typedef struct ltr_bam_reads {
  uint32_t n;
  const int32_t* tid;         /* [n] reference id, -1 unplaced                                         */
  const int32_t* pos;         /* [n] 0-based leftmost reference position                               */
  const int32_t* end;         /* [n] one past the last reference base (htslib bam_endpos)              */
  const uint16_t* flag;       /* [n]                                                                   */
  const uint8_t* mapq;        /* [n]                                                                   */
  const int32_t* mate_tid;    /* [n]                                                                   */
  const int32_t* mate_pos;    /* [n]                                                                   */
  const uint32_t* name_off;   /* [n+1] names: NUL-terminated, back to back                             */
  const char* names;
  const uint32_t* seq_off;    /* [n+1] bases (ASCII, "=ACMGRSVTWYHKDBN") and qualities (Phred+33)      */
  const uint8_t* seq;
  const uint8_t* qual;
  const uint32_t* cigar_off;  /* [n+1] BAM-encoded operations (length << 4 | op, op in MIDNSHP=X)      */
  const uint32_t* cigar_ops;
  const int32_t* hp;          /* [n] integer value of the HP tag, 0 when absent                        */
  const uint32_t* raw_off;    /* [n+1] raw record bytes (empty unless keep_raw)                        */
  const uint8_t* raw;
  void* owner;
} ltr_bam_reads;
int ltr_bam_open(const char* path, const char* index_path, ltr_bam** out);
void ltr_bam_close(ltr_bam* bam);
int32_t ltr_bam_n_refs(const ltr_bam* bam);
const char* ltr_bam_ref_name(const ltr_bam* bam, int32_t tid);
int64_t ltr_bam_ref_len(const ltr_bam* bam, int32_t tid);
int32_t ltr_bam_ref_id(const ltr_bam* bam, const char* name);
const char* ltr_bam_header_text(const ltr_bam* bam);
int ltr_bam_has_index(const ltr_bam* bam);
/* Builds the index of a coordinate-sorted file in memory (one scan) when no .bai was loaded; not to be called while other
 * threads fetch from the same handle.  LTR_ERR_UNSUPPORTED: the file is not coordinate sorted.                         */
int ltr_bam_build_index(ltr_bam* bam);
int ltr_bam_fetch(const ltr_bam* bam, int32_t tid, int64_t beg, int64_t end, int32_t keep_raw, ltr_bam_reads** out);
void ltr_bam_reads_free(ltr_bam_reads* reads);

/* ltr_region_collect   the reads of ONE region from one BAM file per sample, prepared the way LongTR's region loop hands
 *                      them to its genotyper -- single-end (long) reads only:
 *                      window of the query (BamProcessor::process_regions, src/bam_processor.cpp:584-596), read filters and
 *                      the order of reads and samples (read_and_filter_reads, :188-487), phasing terms from the HP tag
 *                      (SNPBamProcessor::process_phased_reads, src/snp_bam_processor.cpp:141-232), spanning test, cut to
 *                      +- flank_size bp around the region and '=XID' CIGAR against the reference sequence
 *                      (GenotyperBamProcessor::left_align_reads, src/genotyper_bam_processor.cpp:38-168;
 *                      BamAlignment::TrimAlignment, src/bam_io.cpp:267-372).  start / stop: 0-based region as LongTR's
 *                      Region holds it.  ref_seq[0] is reference position ref_seq_start of `chrom` (the whole chromosome or
 *                      a slice that covers the reads).  The arrays have the layout of ltr_locus_batch's read fields.     */
typedef struct ltr_region_params {
  int32_t max_mate_dist;     /* MAX_MATE_DIST (1000): the fetched window is [start - d, stop + d]          */
  double min_mean_qual;      /* --min-mean-qual (30): mean Phred quality below it -> LOW_BASE_QUALS         */
  double min_mapq;           /* --min-mapq (20)                                                              */
  int32_t require_spanning;  /* REQUIRE_SPANNING (1)                                                         */
  int32_t min_flank;         /* MIN_FLANK (5): reads covering less are not used for haplotype generation     */
  int32_t flank_size;        /* FLANK_SIZE (200, src/bam_io.h:28)                                            */
  int32_t phased_bam;        /* --phased-bam: phasing terms from the HP tag, otherwise 0 / 0                 */
  int32_t check_hard_clips;  /* BASE_QUAL_TRIM > ' ' (default): hard-clipped reads are dropped               */
} ltr_region_params;
typedef struct ltr_region_reads {
  uint32_t n_samples;                 /* samples that contributed a read, in the reference's order (first appearance
                                         when its read list is emptied from the back: src/bam_processor.cpp:453-483) */
  const uint32_t* sample_file;        /* [n_samples] index into bams                                           */
  const uint32_t* sample_read_begin;  /* [n_samples+1]                                                         */
  uint32_t n_reads;
  const int32_t* read_start;          /* [n_reads] Alignment::get_start()                                      */
  const int32_t* read_stop;           /* [n_reads] Alignment::get_stop(), inclusive                            */
  const uint32_t* read_off;           /* [n_reads+1] bases (upper case) and qualities (Phred+33)               */
  const uint8_t* read_bytes;
  const uint8_t* qual_bytes;
  const uint32_t* cigar_off;          /* [n_reads+1] BAM-encoded '=XID' operations                             */
  const uint32_t* cigar_ops;
  const int32_t* read_sample;         /* [n_reads]                                                             */
  const double* log_p1;               /* [n_reads]                                                             */
  const double* log_p2;
  const uint8_t* hap_gen_ok;          /* [n_reads] usable for haplotype generation (passes_filters)            */
  const uint8_t* deleted;             /* [n_reads] the repeat is deleted in the read (BamAlignment::GetDeleted) */
  const uint32_t* name_off;           /* [n_reads+1] NUL-terminated read names                                 */
  const char* names;
  /* the counters read_and_filter_reads / left_align_reads log */
  uint32_t n_overlapping, n_hard_clipped, n_has_n, n_low_qual, n_low_mapq, n_not_spanning, n_not_unique, n_passed,
      n_trim_failed;
  const int32_t* read_hp;             /* [n_reads] value of the HP tag, 0 without one (PDP of the VCF record)   */
  void* owner;
} ltr_region_reads;
void ltr_region_params_default(ltr_region_params* p);
int ltr_region_collect(const ltr_bam* const* bams, int32_t n_bams, const char* chrom, int32_t start, int32_t stop,
                       const uint8_t* ref_seq, int64_t ref_seq_start, int64_t ref_seq_len, const ltr_region_params* params,
                       ltr_region_reads** out);
void ltr_region_reads_free(ltr_region_reads* reads);

/* ltr_candidate_alleles  the candidate haplotype block of one region from the reads ltr_region_collect prepared:
 *                        HaplotypeGenerator::add_haplotype_block + fuse_haplotype_blocks without --ref-vcf alleles
 *                        (src/SeqAlignment/HaplotypeGenerator.cpp:14-164, 296-481, 521-607; called from
 *                        SeqStutterGenotyper::build_haplotype, src/seq_stutter_genotyper.cpp:416-476).  Alleles: reference
 *                        allele first, then by (length, sequence); block_start / block_end: the RepeatBlock's region after
 *                        the trim; lflank / rflank: the reference-only blocks on either side (lflank starts at lflank_start).
 *                        When some sample leaves more than a quarter of its reads without a candidate the reference
 *                        clusters them (greedy_clustering, thresholds 20 ... 700: :238-271, :403-407), replaces every cluster
 *                        by a partial-order consensus (spoa: :167-199), merges clusters with close consensus sequences
 *                        (:274-292) and adds the consensus of every well-supported cluster as an "inexact" allele
 *                        (:397-471, INEXACT_ALLELE in the VCF record).  ltr_candidate_alleles does the same on the calling
 *                        thread; the consensus restates spoa's published algorithm (spoa is un-vendored and unpinned in the
 *                        reference: parity of the consensus itself is unpinned, csrc/host/poa.cpp); clusters of 30 or more
 *                        sequences, which the reference samples with std::random_device, are sampled with a fixed seed.
 *                        allele_inexact[a] marks consensus alleles, n_consensus counts the consensus computations,
 *                        assembly_threshold is the highest clustering threshold a sample ended on (0: no assembly).
 *                        cluster_* lists, per sample that triggered the assembly, the sequences that were clustered in the
 *                        reference's order with their read counts.
 * ltr_candidate_alleles_flags  the same with LTR_CAND_FLAG_NO_ASSEMBLY: stop before the assembly with status
 *                        LTR_CAND_NEEDS_ASSEMBLY; the alleles found so far are returned and cluster_* holds the sets to
 *                        cluster (e.g. with ltr_cluster_greedy on the device, all thresholds in one call).
 * ltr_poa_consensus      the consensus alone: sequences in the order they are to be added; LTR_ERR_INVALID with *out_len set
 *                        when out_capacity is too small.                                                                 */
#define LTR_CAND_OK 0
#define LTR_CAND_NEAR_CHROM_END 1  /* "Haplotype blocks are too near to the chromosome ends" */
#define LTR_CAND_NO_SPANNING 2     /* "No spanning alignments"                               */
#define LTR_CAND_NEEDS_ASSEMBLY 3
#define LTR_CAND_FLAG_NO_ASSEMBLY 1u
typedef struct ltr_candidates {
  int32_t status;
  int32_t block_start, block_end;
  int32_t n_alleles;
  const uint32_t* allele_off;   /* [n_alleles+1] */
  const uint8_t* allele_bytes;
  int32_t lflank_start;
  const char* lflank;           /* NUL-terminated */
  const char* rflank;
  uint32_t n_cluster_samples;
  const uint32_t* cluster_sample_begin;  /* [n_cluster_samples+1] */
  const uint32_t* cluster_off;           /* [n_cluster_seqs+1]    */
  const uint8_t* cluster_bytes;
  const int32_t* cluster_count;          /* [n_cluster_seqs] reads carrying the sequence */
  const uint8_t* allele_inexact;         /* [n_alleles] 1: consensus of a read cluster   */
  uint32_t n_consensus;
  int32_t assembly_threshold;
  void* owner;
} ltr_candidates;
int ltr_candidate_alleles(const ltr_region_reads* reads, int32_t region_start, int32_t region_stop, int32_t period,
                          const uint8_t* ref_seq, int64_t ref_seq_start, int64_t ref_seq_len, int32_t indel_flank_len,
                          ltr_candidates** out);
int ltr_candidate_alleles_flags(const ltr_region_reads* reads, int32_t region_start, int32_t region_stop, int32_t period,
                                const uint8_t* ref_seq, int64_t ref_seq_start, int64_t ref_seq_len, int32_t indel_flank_len,
                                uint32_t flags, ltr_candidates** out);
int ltr_poa_consensus(const uint8_t* seq_bytes, const uint32_t* seq_off, uint32_t n_seqs, uint8_t* out, uint32_t out_capacity,
                      uint32_t* out_len);
void ltr_candidates_free(ltr_candidates* c);

/* ltr_regions_run   BAM files (one per sample) + regions of ONE chromosome -> genotype calls: LongTR's region loop
 *                   (BamProcessor::process_regions, src/bam_processor.cpp:536-628; GenotyperBamProcessor::
 *                   analyze_reads_and_phasing, src/genotyper_bam_processor.cpp:227-351) with the default stutter model, re-cut
 *                   for the GPU: host threads prepare reads (ltr_region_collect) and candidate alleles
 *                   (ltr_candidate_alleles) of all regions, the survivors become one ltr_locus_batch, ltr_genotyper_run does
 *                   the rest.  status[r] says why a region did not enter the batch; locus_index[r] is its locus in `calls`
 *                   (or -1).  Alleles, block coordinates and the sample order (file index per sample: the order differs from
 *                   region to region exactly as in the reference) are kept per region for whoever writes the records.     */
#define LTR_REGION_OK 0
#define LTR_REGION_INVALID 1
#define LTR_REGION_TOO_LONG 2         /* reference allele longer than --max-tr-len                       */
#define LTR_REGION_NEAR_CONTIG_END 3
#define LTR_REGION_TOO_FEW_READS 4    /* fewer than --min-reads reads passed the filters                  */
#define LTR_REGION_NO_SPANNING 5
#define LTR_REGION_NEEDS_ASSEMBLY 6   /* only with opts->no_assembly: alleles would come from the assembly */
#define LTR_REGION_PAIRED_READS 7     /* paired-end reads: mate logic not reproduced                      */
#define LTR_REGION_DELETED_READ 8     /* a read in which the whole window is deleted (empty sequence)     */
typedef struct ltr_region {
  int32_t start, stop;  /* 0-based, as LongTR's Region holds a BED line */
  int32_t period;
} ltr_region;
typedef struct ltr_regions_opts {
  int32_t host_threads;     /* <= 0: all hardware threads */
  int32_t max_tr_len;       /* --max-tr-len (1000)        */
  int32_t min_total_reads;  /* --min-reads (10)           */
  int32_t no_assembly;      /* 1: report regions that need consensus alleles as LTR_REGION_NEEDS_ASSEMBLY (default 0) */
  int32_t vcf_records;      /* 1: also compose the VCF record of every genotyped region (ltr_vcf_record), one sample column
                               per BAM file; needs region_motifs                                                          */
  const char* const* region_names;   /* [n_regions] or NULL ("."): ID column                                              */
  const char* const* region_motifs;  /* [n_regions] MOTIF / PERIOD of the record                                          */
  int32_t haploid;          /* 1: the chromosome is haploid in every sample (--haploid-chrs): haploid priors and records  */
  uint32_t vcf_switches;    /* LTR_VCF_* output switches of the records (ltr_regions_opts_default: LTR_VCF_DEFAULT)       */
} ltr_regions_opts;
typedef struct ltr_regions_result {
  uint32_t n_regions;
  const int32_t* status;               /* [n_regions] LTR_REGION_*                                             */
  const int32_t* locus_index;          /* [n_regions] locus in calls, -1 when the region was not genotyped     */
  uint32_t n_loci;
  ltr_batch_calls* calls;              /* NULL when no region was genotyped                                    */
  const int32_t* block_start;          /* [n_regions] the RepeatBlock's coordinates                            */
  const int32_t* block_end;
  const uint32_t* region_allele_begin; /* [n_regions+1] candidate alleles (reference allele first)             */
  const uint32_t* allele_off;
  const uint8_t* allele_bytes;
  const uint32_t* region_sample_begin; /* [n_regions+1]                                                        */
  const uint32_t* sample_file;         /* index into bams of each sample of the region                         */
  const uint8_t* allele_inexact;       /* per allele (indexed like allele_off): 1 = consensus of a read cluster */
  uint32_t n_assembled;                /* regions whose candidate alleles went through the assembly branch      */
  const uint32_t* record_off;          /* [n_regions+1] or NULL (opts->vcf_records): record r = records[record_off[r] ..
                                          record_off[r+1]), empty for a region that was not genotyped; no newlines       */
  const char* records;
  void* owner;
  double prepare_ms, layout_ms, genotype_ms, records_ms;  /* host timing of the call: region preparation on the host threads
                                          (BAM fetch, filters, trimming, candidate alleles), laying the batch out, ltr_genotyper_run,
                                          composing the records                                                          */
} ltr_regions_result;
void ltr_regions_opts_default(ltr_regions_opts* o);
int ltr_regions_run(ltr_genotyper* g, const ltr_params* params, const ltr_bam* const* bams, int32_t n_bams, const char* chrom,
                    const ltr_region* regions, uint32_t n_regions, const uint8_t* ref_seq, int64_t ref_seq_start,
                    int64_t ref_seq_len, const ltr_region_params* rp, const ltr_regions_opts* opts, ltr_regions_result** out);
void ltr_regions_result_free(ltr_regions_result* r);

/* ---- reference sequence, region file, the whole run ---------------------------------
 * ltr_fasta_*   indexed FASTA access in place of htslib's faidx behind the reference's FastaReader (src/fasta_reader.{h,cpp}):
 *               `path` is one FASTA file or a directory whose *.fa files are all loaded; the .fai index next to a file is
 *               used when present and built in memory otherwise; a sequence name occurring twice is LTR_ERR_INVALID,
 *               compressed FASTA LTR_ERR_UNSUPPORTED.  ltr_fasta_fetch copies bases [start, end) as they stand in the file.
 * ltr_bed_read  readRegions + orderRegions (src/region.cpp:26-75): lines CHROM START STOP MOTIF [NAME] with 1-based START;
 *               regions come back 0-based, sorted by (chromosome, start, stop), grouped by chromosome; period = the motif
 *               length, -1 when comma-separated motifs differ in length (Region::computePeriod, src/region.h:36-43);
 *               chrom_limit (may be NULL) keeps one chromosome; max_regions 0 = no limit.  A malformed line is
 *               LTR_ERR_INVALID (the reference exits).
 * ltr_run_bed   BamProcessor::process_regions (src/bam_processor.cpp:536-628): chromosomes of the region file are checked
 *               against the FASTA and the alignment files (verify_chromosomes :490-531, LTR_ERR_INVALID when one is
 *               missing), then every chromosome's sequence is fetched once and its regions go through ltr_regions_run.
 *               per_chrom[c] covers bed->regions[chrom_region_begin[c] .. chrom_region_begin[c+1]).                   */
typedef struct ltr_fasta ltr_fasta;
int ltr_fasta_open(const char* path, ltr_fasta** out);
void ltr_fasta_close(ltr_fasta* fa);
int32_t ltr_fasta_n_seqs(const ltr_fasta* fa);
const char* ltr_fasta_seq_name(const ltr_fasta* fa, int32_t i);
int64_t ltr_fasta_seq_len(const ltr_fasta* fa, const char* name); /* -1: no such sequence */
int ltr_fasta_fetch(const ltr_fasta* fa, const char* name, int64_t start, int64_t end, uint8_t* out);
typedef struct ltr_bed {
  uint32_t n_regions;
  const ltr_region* regions;      /* [n_regions] sorted, grouped by chromosome */
  const int32_t* region_chrom;    /* [n_regions] index into chroms             */
  const char* const* names;       /* [n_regions] "" when the line has no name  */
  const char* const* motifs;      /* [n_regions]                               */
  uint32_t n_chroms;
  const char* const* chroms;      /* [n_chroms] in order of appearance         */
  void* owner;
} ltr_bed;
int ltr_bed_read(const char* path, uint32_t max_regions, const char* chrom_limit, ltr_bed** out);
void ltr_bed_free(ltr_bed* b);
typedef struct ltr_bed_run_result {
  uint32_t n_chroms;
  ltr_regions_result* const* per_chrom;  /* [n_chroms]   */
  const uint32_t* chrom_region_begin;    /* [n_chroms+1] */
  void* owner;
} ltr_bed_run_result;
int ltr_run_bed(ltr_genotyper* g, const ltr_params* params, const ltr_bam* const* bams, int32_t n_bams, const ltr_fasta* fasta,
                const ltr_bed* bed, const ltr_region_params* rp, const ltr_regions_opts* opts, ltr_bed_run_result** out);
void ltr_bed_run_result_free(ltr_bed_run_result* r);
/* ltr_run_bed in bounded memory (whole-genome region files): a chromosome's regions go through ltr_regions_run
 * chunk_regions at a time (<= 0: 4096) and every chunk's result is handed to `sink` -- chromosome index, index of the chunk's
 * first region in bed->regions, the result (valid during the call only; the library frees it) -- in region order, the way the
 * reference's region loop hands its records to the VCFWriter one region after the other (src/bam_processor.cpp:563-627).
 * A sink that returns non-zero stops the run (LTR_ERR_INVALID).  Calls and records equal ltr_run_bed's.                     */
typedef int (*ltr_regions_sink)(void* user, uint32_t chrom, uint32_t first_region, const ltr_regions_result* result);
int ltr_run_bed_stream(ltr_genotyper* g, const ltr_params* params, const ltr_bam* const* bams, int32_t n_bams, const ltr_fasta* fasta,
                       const ltr_bed* bed, const ltr_region_params* rp, const ltr_regions_opts* opts, int32_t chunk_regions,
                       ltr_regions_sink sink, void* user);

/* ---- VCF records (the output side of the path) ---------------------------------------------------
 * ltr_vcf_record   the text of one record as SeqStutterGenotyper::write_vcf_record composes it (src/seq_stutter_genotyper.cpp:
 *                  894-1402; get_alleles :688-781, reorder_alleles :667-686) with the reference's default output switches
 *                  (ALLREADS, MALLREADS on; GL / PL / PHASEDGL / FILTER off) on the long-read path: CHROM POS ID REF ALT . .
 *                  INFO (START END MOTIF PERIOD NSKIP NFILT INEXACT_ALLELE BPDIFFS DP DSNP DFLANKINDEL AN REFAC AC) FORMAT and
 *                  one column per entry of column_sample ("." for -1 or a sample without reads).  Alleles are the candidates
 *                  that survive (kept_mask; NULL = all), genotypes are candidate indices; read_allele (needed when a sample is
 *                  heterozygous) is the allele each read is assigned to (ltr_batch_calls.read_allele); read_bp_diff is
 *                  ltr_extract_cigar_bp_diff of the read over [region_start - 5, region_stop + 5], INT32_MIN where it returns
 *                  0.  No trailing newline; LTR_ERR_INVALID with *out_len set when the buffer is too small; empty ("<DEL>")
 *                  alleles are LTR_ERR_UNSUPPORTED.  The header lines of the file are not produced.
 * ltr_extract_cigar_bp_diff  ExtractCigar (src/extract_indels.cpp:18-93) on BAM-encoded CIGAR operations.                  */
typedef struct ltr_vcf_locus {
  const char* chrom;
  const char* name;               /* NULL or "": "."                                                    */
  const char* motif;              /* the region's motif column (comma separated list)                   */
  int32_t region_start, region_stop;  /* 0-based start as LongTR's Region holds it                      */
  const uint8_t* chrom_seq;       /* reference bases [chrom_seq_start, chrom_seq_start + chrom_seq_len) */
  int64_t chrom_seq_start, chrom_seq_len;
  int32_t block_start, block_end; /* the candidate block (ltr_candidates)                               */
  int32_t n_alleles;
  const uint32_t* allele_off;     /* [n_alleles+1] */
  const uint8_t* allele_bytes;
  const uint8_t* allele_inexact;  /* [n_alleles] or NULL */
  const uint8_t* kept_mask;       /* [n_alleles] or NULL */
  int32_t haploid;
  int32_t n_samples;              /* samples of the locus, in the locus' order                          */
  const int32_t* gts;             /* [2 * n_samples] candidate indices                                  */
  const double* log_unphased_posteriors;
  const double* log_phased_posteriors;
  const double* gl_diffs;
  const int32_t* n_p1;            /* [n_samples] reads tagged HP = 1 / 2 (PDP), NULL = zeros            */
  const int32_t* n_p2;
  int32_t n_reads;
  const int32_t* read_sample;     /* [n_reads] */
  const double* log_p1;
  const double* log_p2;
  const int32_t* read_bp_diff;    /* [n_reads] or NULL (ALLREADS "."), INT32_MIN = none for this read   */
  const int32_t* read_allele;     /* [n_reads] or NULL when no sample is heterozygous                   */
  int32_t n_columns;
  const int32_t* column_sample;   /* [n_columns] sample of the locus shown in the column, -1 = none     */
} ltr_vcf_locus;
int ltr_vcf_record(const ltr_vcf_locus* locus, char* out, uint32_t capacity, uint32_t* out_len);
/* The reference's output switches (Genotyper::OUTPUT_*, src/genotyper.cpp:339-346; command line --hide-allreads,
 * --hide-mallreads, --output-gls, --output-pls, --output-phased-gls, --output-filters: src/hipstr_main.cpp:178-183).
 * ltr_vcf_record_ex composes the record under any combination of them (write_vcf_record :1182-1229, :1304-1365):
 * GL / PL in the order of the record's alleles (pairs (j <= i) of the re-ordered alleles), PHASEDGL [i * K + j], FILTER =
 * PASS, and "NO_READS" behind the empty fields for a column without reads.  gls / pls / phased_gls are what
 * ltr_batch_calls delivers: per sample of the locus a slice over the KEPT alleles in candidate order (gls / pls: pair
 * (a <= b) at b(b+1)/2 + a, haploid: [a]; phased_gls: [a * K + b], haploid: [a]); a slice holds at least that many
 * entries.  extras == NULL or switches == LTR_VCF_DEFAULT: ltr_vcf_record.                                           */
#define LTR_VCF_ALLREADS 1u
#define LTR_VCF_MALLREADS 2u
#define LTR_VCF_GLS 4u
#define LTR_VCF_PLS 8u
#define LTR_VCF_PHASED_GLS 16u
#define LTR_VCF_FILTERS 32u
#define LTR_VCF_DEFAULT (LTR_VCF_ALLREADS | LTR_VCF_MALLREADS)
typedef struct ltr_vcf_extras {
  uint32_t switches;          /* LTR_VCF_*                                                  */
  const uint64_t* gl_begin;   /* [n_samples+1] slices of gls / pls (GL or PL switched on)   */
  const double* gls;
  const int32_t* pls;
  const uint64_t* pgl_begin;  /* [n_samples+1] slices of phased_gls (PHASEDGL switched on, diploid) */
  const double* phased_gls;
} ltr_vcf_extras;
int ltr_vcf_record_ex(const ltr_vcf_locus* locus, const ltr_vcf_extras* extras, char* out, uint32_t capacity, uint32_t* out_len);
int ltr_extract_cigar_bp_diff(const uint32_t* cigar_ops, uint32_t n_ops, int32_t cigar_start, int32_t region_start,
                              int32_t region_end, int32_t* bp_diff);
/* The header lines of the file (Genotyper::get_vcf_header, src/genotyper.cpp:258-336, default output switches): file format,
 * command, reference, one contig line per sequence of the FASTA, the INFO / FORMAT descriptions, the #CHROM line with the
 * sample columns.  Ends with a newline.  LTR_ERR_INVALID with *out_len set when the buffer is too small.                   */
int ltr_vcf_header(const ltr_fasta* fasta, const char* fasta_path, const char* command, const char* const* sample_names,
                   uint32_t n_samples, char* out, uint32_t capacity, uint32_t* out_len);
/* The same under the output switches `switches` (LTR_VCF_*): the FORMAT lines of the switched fields (:313-327). */
int ltr_vcf_header_ex(const ltr_fasta* fasta, const char* fasta_path, const char* command, const char* const* sample_names,
                      uint32_t n_samples, uint32_t switches, char* out, uint32_t capacity, uint32_t* out_len);

/* ---- length-based EM of the stutter model ----------------------------------------------
 * ltr_em_stutter_train  EMStutterGenotyper(...).train(...) (src/em_stutter_genotyper.{h,cpp}) for many loci at once, as
 *                       GenotyperBamProcessor::learn_stutter_model calls it (src/genotyper_bam_processor.cpp:170-225): per
 *                       read the observed size difference to the reference in bp (ExtractCigar) and its phasing terms; the
 *                       allele sizes of a locus are the distinct observed differences, the reference allele (0) first.
 *                       One warp per locus runs the whole EM loop on the device.  out_params[6 * l ..] = in-frame geometric
 *                       parameter, P(up), P(down), out-of-frame geometric parameter, P(up), P(down) of the last model;
 *                       out_trained[l] = what train() returns; out_n_iter[l] = E steps done; out_ll[l] = last total
 *                       log-likelihood; out_log_gt_priors (optional) = the allele log-frequencies, prior_stride entries per
 *                       locus (alleles beyond it are dropped).  Phasing terms must be <= 0 (the reference asserts it).  In the
 *                       reference's CLI this path is switched off by the default stutter model (hipstr_main.cpp:140, 362);
 *                       the class is the oracle (oracle/em_driver.cpp).                                                   */
typedef struct ltr_em_batch {
  uint32_t n_loci;
  const uint32_t* locus_sample_begin;  /* [n_loci+1]    samples of the locus                           */
  const uint32_t* sample_read_begin;   /* [n_samples+1] reads of the sample (reads are sample-major)   */
  const int32_t* read_bp_diff;         /* [n_reads]                                                    */
  const double* log_p1;                /* [n_reads]                                                    */
  const double* log_p2;                /* [n_reads]                                                    */
  const int32_t* locus_motif_len;      /* [n_loci]                                                     */
  const uint8_t* locus_haploid;        /* [n_loci] or NULL                                             */
} ltr_em_batch;
typedef struct ltr_em_opts {
  int32_t max_iter;          /* MAX_EM_ITER (100)       */
  double abs_ll_converge;    /* ABS_LL_CONVERGE (0.01)  */
  double frac_ll_converge;   /* FRAC_LL_CONVERGE (0.001) */
} ltr_em_opts;
void ltr_em_opts_default(ltr_em_opts* o);
int ltr_em_stutter_train(ltr_ctx* ctx, const ltr_em_batch* batch, const ltr_em_opts* opts, double* out_params,
                         int32_t* out_trained, int32_t* out_n_iter, double* out_ll, double* out_log_gt_priors,
                         uint32_t prior_stride);

/* ---- diagnostics --------------------------------------------------------------------- */
/* Sustained FP64-pipe issue rate of the device in lane-operations per second (the roofline
 * denominator of SURVEY.md section 8d): kind 0 = DADD, 1 = DSETP, 2 = the DADD,DADD,DSETP,
 * 2xFSEL pattern of one max-plus term.  ms (optional) = duration of the probe kernel.      */
int ltr_fp64_issue_rate(int device, int kind, double* lane_ops_per_s, double* ms);

/* ---- synthetic workload of BASELINE.json config 5 (configs 3 and 4: longtr_b200/synth/longtr_synth.h, a library of its own) ---- */
/* config 5 (homopolymers, --stutter-align-len path): an ltr_stutter_batch with whole pooled reads, median
 * qualities and calc_seed_base seeds, plus what is needed to replay the same loci as flat loci
 * (include/longtr_b200_locus.h): per-read reference coordinates and CIGAR, per-locus repeat coordinates. */
typedef struct ltr_synth_stutter_batch {
  ltr_stutter_batch batch;
  ltr_posterior_batch post;      /* 30 sample-reads per locus, one sample */
  const int32_t* read_start;     /* [n_reads] */
  const int32_t* read_stop;      /* [n_reads] */
  const uint32_t* cigar_off;     /* [n_reads+1] */
  const uint8_t* cigar_bytes;
  const int32_t* repeat_start;   /* [n_loci] */
  const int32_t* repeat_end;     /* [n_loci] */
  uint32_t n_alleles, n_reads, n_sreads;
} ltr_synth_stutter_batch;
int ltr_synth_stutter_generate(uint64_t base_seed, uint32_t first_locus, uint32_t n_loci, int n_threads,
                               ltr_synth_stutter_batch** out);
void ltr_synth_stutter_free(ltr_synth_stutter_batch* b);

#ifdef __cplusplus
}
#endif
#endif
