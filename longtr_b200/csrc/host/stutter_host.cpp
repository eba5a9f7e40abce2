// stutter_host.cpp -- homopolymer / --stutter-align-len path of HapAligner::process_reads
// (reference src/SeqAlignment/HapAligner.cpp:567-579, 855-975): host preparation for kernel 2.
#include "longtr_host.h"

namespace ltr {

void HapAligner::process_reads_short(const std::vector<Alignment>& alns, int init_read_index, const BaseQuality* bq,
                                     const std::vector<bool>& realign_read, double* aln_probs, int* seed_positions) {
  (void)alns; (void)init_read_index; (void)bq; (void)realign_read; (void)aln_probs; (void)seed_positions;
  status_ = LTR_ERR_UNSUPPORTED;
}

}  // namespace ltr
