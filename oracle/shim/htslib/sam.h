/* Compile-only stand-in for htslib/sam.h (htslib is NOT vendored by LongTR and is
 * not installed here).  TEST INFRASTRUCTURE ONLY: lets the reference's hot-path
 * translation units (which include bam_io.h for the CigarOp type) compile in place
 * from /root/reference for the oracle/_ref build.  Only declarations that
 * /root/reference/src/bam_io.h names are provided; none of them is ever defined
 * or linked -- the inline wrappers that use them are never emitted.            */
#ifndef LTR_SHIM_HTSLIB_SAM_H
#define LTR_SHIM_HTSLIB_SAM_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { int32_t tid; int32_t pos; uint16_t bin; uint8_t qual; uint8_t l_qname;
                 uint16_t flag; uint16_t unused; uint32_t n_cigar; int32_t l_qseq;
                 int32_t mtid; int32_t mpos; int32_t isize; } bam1_core_t;
typedef struct { bam1_core_t core; int l_data; uint32_t m_data; uint8_t* data; uint64_t id; } bam1_t;
typedef struct sam_hdr_t { int32_t n_targets; int32_t ignore_sam_err; size_t l_text;
                           uint32_t* target_len; char** target_name; char* text; void* sdict; } sam_hdr_t;
typedef sam_hdr_t bam_hdr_t;
typedef struct htsFile { uint32_t is_bin:1, is_write:1, is_be:1, is_cram:1, is_bgzf:1, dummy:27;
                         int64_t lineno; void* fp; } htsFile;
typedef htsFile samFile;
typedef struct hts_idx_t hts_idx_t;
typedef struct { uint64_t u, v; } hts_pair64_t;
typedef struct hts_itr_t { uint32_t read_rest:1, finished:1, is_cram:1, nocoor:1, multi:1, dummy:27;
                           int tid, n_off, i, n_reg; int64_t beg, end; void* reg_list;
                           int curr_tid, curr_reg, curr_intv; int64_t curr_beg, curr_end;
                           uint64_t curr_off, nocoor_off; hts_pair64_t* off; } hts_itr_t;
struct BGZF;

#define BAM_FPAIRED        1
#define BAM_FPROPER_PAIR   2
#define BAM_FUNMAP         4
#define BAM_FMUNMAP        8
#define BAM_FREVERSE      16
#define BAM_FMREVERSE     32
#define BAM_FREAD1        64
#define BAM_FREAD2       128
#define BAM_FSECONDARY   256
#define BAM_FQCFAIL      512
#define BAM_FDUP        1024
#define BAM_FSUPPLEMENTARY 2048

#define BAM_CIGAR_STR   "MIDNSHP=XB"
#define BAM_CIGAR_SHIFT 4
#define BAM_CIGAR_MASK  0xf
#define bam_cigar_op(c)    ((c)&BAM_CIGAR_MASK)
#define bam_cigar_oplen(c) ((c)>>BAM_CIGAR_SHIFT)
#define bam_cigar_opchr(c) (BAM_CIGAR_STR "??????" [bam_cigar_op(c)])
#define bam_get_qname(b) ((char*)(b)->data)
#define bam_get_cigar(b) ((uint32_t*)((b)->data + (b)->core.l_qname))
#define bam_get_seq(b)   ((b)->data + ((b)->core.n_cigar<<2) + (b)->core.l_qname)
#define bam_get_qual(b)  ((b)->data + ((b)->core.n_cigar<<2) + (b)->core.l_qname + (((b)->core.l_qseq + 1)>>1))
#define bam_seqi(s, i)   ((s)[(i)>>1] >> ((~(i)&1)<<2) & 0xf)

bam1_t* bam_init1(void);
void bam_destroy1(bam1_t* b);
bam1_t* bam_copy1(bam1_t* dst, const bam1_t* src);
int32_t bam_endpos(const bam1_t* b);
uint8_t* bam_aux_get(const bam1_t* b, const char tag[2]);
int bam_aux_del(bam1_t* b, uint8_t* s);
int bam_aux_append(bam1_t* b, const char tag[2], char type, int len, const uint8_t* data);
char bam_aux2A(const uint8_t* s);
int64_t bam_aux2i(const uint8_t* s);
double bam_aux2f(const uint8_t* s);
char* bam_aux2Z(const uint8_t* s);

samFile* sam_open(const char* fn, const char* mode);
int sam_close(samFile* fp);
sam_hdr_t* sam_hdr_read(samFile* fp);
void bam_hdr_destroy(sam_hdr_t* h);
void sam_hdr_destroy(sam_hdr_t* h);
hts_idx_t* sam_index_load(samFile* fp, const char* fn);
void hts_idx_destroy(hts_idx_t* idx);
hts_itr_t* sam_itr_querys(const hts_idx_t* idx, sam_hdr_t* hdr, const char* region);
int sam_itr_next(samFile* fp, hts_itr_t* itr, bam1_t* r);
void hts_itr_destroy(hts_itr_t* iter);
int hts_set_fai_filename(htsFile* fp, const char* fn_aux);
int bam_hdr_write(struct BGZF* fp, const sam_hdr_t* h);
int bam_write1(struct BGZF* fp, const bam1_t* b);

#ifdef __cplusplus
}
#endif
#endif
