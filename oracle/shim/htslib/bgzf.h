/* Compile-only stand-in for htslib/bgzf.h -- see sam.h in this directory. */
#ifndef LTR_SHIM_HTSLIB_BGZF_H
#define LTR_SHIM_HTSLIB_BGZF_H
#include <stdint.h>
#include <sys/types.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct BGZF BGZF;
BGZF* bgzf_open(const char* path, const char* mode);
int bgzf_close(BGZF* fp);
ssize_t bgzf_write(BGZF* fp, const void* data, size_t length);
int bgzf_getc(BGZF* fp);
#ifdef __cplusplus
}
#endif
#endif
