/* Compile-only stand-in for htslib/tbx.h -- see sam.h in this directory. */
#ifndef LTR_SHIM_HTSLIB_TBX_H
#define LTR_SHIM_HTSLIB_TBX_H
#include "sam.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct tbx_t tbx_t;
typedef struct { size_t l, m; char* s; } kstring_t;
tbx_t* tbx_index_load(const char* fn);
void tbx_destroy(tbx_t* tbx);
hts_itr_t* tbx_itr_querys(tbx_t* tbx, const char* reg);
int tbx_itr_next(htsFile* fp, tbx_t* tbx, hts_itr_t* iter, void* data);
void tbx_itr_destroy(hts_itr_t* iter);
const char** tbx_seqnames(tbx_t* tbx, int* n);
int tbx_name2id(tbx_t* tbx, const char* ss);
htsFile* hts_open(const char* fn, const char* mode);
int hts_close(htsFile* fp);
int hts_getline(htsFile* fp, int delimiter, kstring_t* str);
#ifdef __cplusplus
}
#endif
#endif
