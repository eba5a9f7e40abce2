// flat_api.cpp -- ltr_process_reads_flat: HapAligner::process_reads on a flat locus.
#include "longtr_b200.h"

extern "C" int ltr_process_reads_flat(ltr_ctx* ctx, const ltr_flat_locus* locus, double* out_ll,
                                      int32_t* out_seeds) {
  (void)ctx; (void)locus; (void)out_ll; (void)out_seeds;
  return LTR_ERR_UNSUPPORTED;
}
