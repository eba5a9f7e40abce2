/* Compile-only stand-in for htslib/faidx.h -- see sam.h in this directory. */
#ifndef LTR_SHIM_HTSLIB_FAIDX_H
#define LTR_SHIM_HTSLIB_FAIDX_H
#ifdef __cplusplus
extern "C" {
#endif
typedef struct faidx_t faidx_t;
faidx_t* fai_load(const char* fn);
void fai_destroy(faidx_t* fai);
char* fai_fetch(const faidx_t* fai, const char* reg, int* len);
char* faidx_fetch_seq(const faidx_t* fai, const char* c_name, int p_beg_i, int p_end_i, int* len);
int faidx_has_seq(const faidx_t* fai, const char* seq);
int faidx_seq_len(const faidx_t* fai, const char* seq);
int faidx_nseq(const faidx_t* fai);
const char* faidx_iseq(const faidx_t* fai, int i);
#ifdef __cplusplus
}
#endif
#endif
