/* Compile-only stand-in for htslib/kfunc.h -- see sam.h in this directory.  kt_fisher_exact is computed by
 * LongTR (seq_stutter_genotyper.cpp:1250) but its result is never printed (output_strand_bias=false, :1168). */
#ifndef LTR_SHIM_HTSLIB_KFUNC_H
#define LTR_SHIM_HTSLIB_KFUNC_H
#ifdef __cplusplus
extern "C" {
#endif
double kt_fisher_exact(int n11, int n12, int n21, int n22, double* _left, double* _right, double* two);
#ifdef __cplusplus
}
#endif
#endif
