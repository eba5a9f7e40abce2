/* Compile-only stand-in for htslib/vcf.h -- see sam.h in this directory.  Only what LongTR's vcf_reader.h /
 * vcf_input.cpp name; none of it is reachable from the IO-less per-locus genotyper the oracle drives. */
#ifndef LTR_SHIM_HTSLIB_VCF_H
#define LTR_SHIM_HTSLIB_VCF_H
#include <stdint.h>
#include "tbx.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct bcf_hdr_t bcf_hdr_t;
typedef struct { int id; int n, size, type; uint8_t* p; } bcf_fmt_t;
typedef struct { int key; int type; int len; uint8_t* vptr; } bcf_info_t;
typedef struct { int m_fmt, m_info, m_id, m_als, m_allele, m_flt; int n_flt; int* flt; char* id; char* als;
                 char** allele; bcf_info_t* info; bcf_fmt_t* fmt; void* var; int n_var, var_type; } bcf_dec_t;
typedef struct { int32_t rid; int32_t pos; int32_t rlen; float qual; uint32_t n_info : 16, n_allele : 16;
                 uint32_t n_fmt : 8, n_sample : 24; kstring_t shared, indiv; bcf_dec_t d; } bcf1_t;
typedef htsFile vcfFile;
#define BCF_UN_ALL 15
#define BCF_UN_STR 1
#define BCF_HT_INT 1
#define BCF_HT_REAL 2
#define BCF_HT_STR 3
#define bcf_gt_is_missing(v) ((v) >> 1 ? 0 : 1)
#define bcf_gt_is_phased(v) ((v) & 1)
#define bcf_gt_allele(v) (((v) >> 1) - 1)
#define bcf_int32_vector_end (-2147483647 - 1 + 1)
#define bcf_int32_missing (-2147483647 - 1)
extern uint32_t bcf_float_missing;
extern uint32_t bcf_float_vector_end;
int bcf_float_is_missing(float f);
int bcf_float_is_vector_end(float f);
int bcf_hdr_nsamples_fn(const bcf_hdr_t* h);
#define bcf_hdr_nsamples(h) bcf_hdr_nsamples_fn(h)
char** bcf_hdr_samples_fn(const bcf_hdr_t* h);
int bcf_unpack(bcf1_t* b, int which);
int bcf_is_snp(bcf1_t* v);
const char* bcf_seqname(const bcf_hdr_t* hdr, const bcf1_t* rec);
bcf_fmt_t* bcf_get_fmt(const bcf_hdr_t* hdr, bcf1_t* line, const char* key);
bcf_info_t* bcf_get_info(const bcf_hdr_t* hdr, bcf1_t* line, const char* key);
int bcf_get_info_values(const bcf_hdr_t* hdr, bcf1_t* line, const char* tag, void** dst, int* ndst, int type);
int bcf_get_format_values(const bcf_hdr_t* hdr, bcf1_t* line, const char* tag, void** dst, int* ndst, int type);
int bcf_get_format_string(const bcf_hdr_t* hdr, bcf1_t* line, const char* tag, char*** dst, int* ndst);
#define bcf_get_info_int32(hdr, line, tag, dst, ndst) bcf_get_info_values(hdr, line, tag, (void**)(dst), ndst, BCF_HT_INT)
#define bcf_get_info_float(hdr, line, tag, dst, ndst) bcf_get_info_values(hdr, line, tag, (void**)(dst), ndst, BCF_HT_REAL)
#define bcf_get_info_string(hdr, line, tag, dst, ndst) bcf_get_info_values(hdr, line, tag, (void**)(dst), ndst, BCF_HT_STR)
#define bcf_get_format_int32(hdr, line, tag, dst, ndst) bcf_get_format_values(hdr, line, tag, (void**)(dst), ndst, BCF_HT_INT)
#define bcf_get_format_float(hdr, line, tag, dst, ndst) bcf_get_format_values(hdr, line, tag, (void**)(dst), ndst, BCF_HT_REAL)
#define bcf_get_genotypes(hdr, line, dst, ndst) bcf_get_format_values(hdr, line, "GT", (void**)(dst), ndst, BCF_HT_INT)
bcf_hdr_t* bcf_hdr_read(htsFile* fp);
void bcf_hdr_destroy(bcf_hdr_t* h);
bcf1_t* bcf_init(void);
void bcf_destroy(bcf1_t* v);
int bcf_read(htsFile* fp, const bcf_hdr_t* h, bcf1_t* v);
int vcf_parse(kstring_t* s, const bcf_hdr_t* h, bcf1_t* v);
int bcf_hdr_id2int(const bcf_hdr_t* hdr, int type, const char* id);
const char** bcf_hdr_seqnames(const bcf_hdr_t* h, int* nseqs);
#define bcf_init1() bcf_init()
#define bcf_destroy1(v) bcf_destroy(v)
#define BCF_DT_ID 0
#define BCF_DT_CTG 1
#define BCF_DT_SAMPLE 2
#ifdef __cplusplus
}
#endif
#endif
