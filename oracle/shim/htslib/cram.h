/* Compile-only stand-in for htslib/cram.h -- see sam.h in this directory. */
#ifndef LTR_SHIM_HTSLIB_CRAM_H
#define LTR_SHIM_HTSLIB_CRAM_H
#endif
