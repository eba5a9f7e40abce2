// lazy_haplotype_alignment.cpp -- LongTR-side replacement of Haplotype::aln_haps_to_ref.
//
// The reference's Haplotype constructor (src/SeqAlignment/Haplotype.h:34-50) aligns EVERY candidate haplotype to the
// reference haplotype with a float affine Needleman-Wunsch and traceback (src/SeqAlignment/Haplotype.cpp:58-86 ->
// NeedlemanWunsch::Align), and the constructor runs at least four times per locus (initial haplotype, the reversed copy
// inside each HapAligner: HapAligner.h:94-120, add_and_remove_alleles: src/seq_stutter_genotyper.cpp:317-409, :941).
// Its only product, hap_aln_info_, is read by Haplotype::get_aln_info(), whose single caller is the traceback stitching
// at HapAligner.cpp:969 -- reachable only with retrace_aln, i.e. only through HapAligner::retrace, which returns NULL in
// this snapshot (HapAligner.cpp:601-810).  On the live path nothing reads it; once the DP is on the GPU
// it is what is left of the per-locus host time for long repeats (0.24-0.37 s per constructor at 1 kb).
//
// The binding therefore keeps the member's signature and leaves hap_aln_info_ with one empty string per haplotype
// (so that Haplotype::reverse, Haplotype.cpp:304-306, and the indexing in get_aln_info stay valid).  The reference's own
// definition is kept under another name (oracle/build_ref.sh renames the symbol in a COPY of the object file) and runs
// instead when LONGTR_B200_EAGER_HAP_ALIGNMENT is set -- the before/after timings of profiles/r2_nw_elision.txt come from
// that switch.  A maintainer who revives the traceback makes get_aln_info() call the original on first use.
#include <cstdlib>
#include <string>

#include "SeqAlignment/Haplotype.h"

extern "C" void ltr_orig_aln_haps_to_ref(Haplotype* self);  // the reference's Haplotype::aln_haps_to_ref, renamed

void Haplotype::aln_haps_to_ref() {
  static const bool eager = std::getenv("LONGTR_B200_EAGER_HAP_ALIGNMENT") != NULL;
  if (eager) {
    ltr_orig_aln_haps_to_ref(this);
    return;
  }
  hap_aln_info_.assign((size_t)num_combs(), std::string());
}
