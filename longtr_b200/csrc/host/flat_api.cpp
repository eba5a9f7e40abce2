// flat_api.cpp -- ltr_process_reads_flat: HapAligner::process_reads on one flat locus
// (include/longtr_b200_locus.h), through the host mirror of the reference classes.
//   blocks  : HapBlock(left flank), RepeatBlock(alleles, period, stutter model), HapBlock(right flank)
//             -- what HaplotypeGenerator::fuse_haplotype_blocks hands to Haplotype (reference
//             src/SeqAlignment/HaplotypeGenerator.cpp:580-607, Haplotype.h:34-50)
//   reads   : Alignment(start, stop, qualities, sequence) + CIGAR (AlignmentData.h:32-60)
//   call    : HapAligner(haplotype, realign_to_hap, INDEL_FLANK_LEN, SWITCH_OLD_ALIGN_LEN, params)
//             .process_reads(alns, 0, &base_quality, realign_read, out_ll, out_seeds)  (HapAligner.h:94-138)
#include <string.h>

#include <memory>

#include "longtr_host.h"

extern "C" int ltr_process_reads_flat(ltr_ctx* ctx, const ltr_flat_locus* L, double* out_ll, int32_t* out_seeds) {
  using namespace ltr;
  if (!ctx || !L || !out_ll || !out_seeds) return LTR_ERR_INVALID;
  if (!L->lflank || !L->rflank || !L->alleles || L->n_alleles < 1 || L->n_reads < 0 || !L->motif) return LTR_ERR_INVALID;
  if (L->n_reads > 0 && !L->reads) return LTR_ERR_INVALID;
  if (L->period < 1 || L->indel_flank_len < 0 || L->indel_flank_len > 35) return LTR_ERR_INVALID;
  if (L->n_aln_params != 0 && L->n_aln_params != 7) return LTR_ERR_INVALID;
  StutterModel model(L->stutter[0], L->stutter[1], L->stutter[2], L->stutter[3], L->stutter[4], L->stutter[5],
                     std::string(L->motif));
  if (!model.valid()) return LTR_ERR_INVALID;
  model.set_period(L->period);
  const std::string lflank(L->lflank), rflank(L->rflank);
  for (int a = 0; a < L->n_alleles; ++a)
    if (!L->alleles[a]) return LTR_ERR_INVALID;
  HapBlock left(L->repeat_start - (int32_t)lflank.size(), L->repeat_start, lflank);
  RepeatBlock repeat(L->repeat_start, L->repeat_end, std::string(L->alleles[0]), L->period, &model);
  for (int a = 1; a < L->n_alleles; ++a) repeat.add_alternate(std::make_pair(std::string(L->alleles[a]), false));
  HapBlock right(L->repeat_end, L->repeat_end + (int32_t)rflank.size(), rflank);
  std::vector<HapBlock*> blocks;
  blocks.push_back(&left);
  blocks.push_back(&repeat);
  blocks.push_back(&right);
  Haplotype haplotype(blocks);

  std::vector<Alignment> alns;
  alns.reserve((size_t)L->n_reads);
  for (int r = 0; r < L->n_reads; ++r) {
    const ltr_flat_read& fr = L->reads[r];
    if (!fr.seq || !fr.qual || !fr.cigar) return LTR_ERR_INVALID;
    alns.push_back(Alignment(fr.start, fr.stop, false, false, "read", std::string(fr.qual), std::string(fr.seq),
                             std::string(fr.seq)));
    if (!alns.back().set_cigar_string(fr.cigar)) return LTR_ERR_INVALID;
  }
  std::vector<bool> realign_hap((size_t)L->n_alleles, true), realign_read((size_t)L->n_reads, true);
  if (L->realign_to_hap)
    for (int a = 0; a < L->n_alleles; ++a) realign_hap[a] = L->realign_to_hap[a] != 0;
  if (L->realign_read)
    for (int r = 0; r < L->n_reads; ++r) realign_read[r] = L->realign_read[r] != 0;
  std::vector<float> params(L->aln_params, L->aln_params + L->n_aln_params);
  BaseQuality base_quality;
  HapAligner aligner(&haplotype, realign_hap, L->indel_flank_len, L->switch_old_align_len, params, ctx);
  std::vector<int> seeds((size_t)L->n_reads);
  for (int r = 0; r < L->n_reads; ++r) seeds[r] = out_seeds[r];
  aligner.process_reads(alns, 0, &base_quality, realign_read, out_ll, seeds.data());
  if (aligner.status() != LTR_OK) return aligner.status();
  for (int r = 0; r < L->n_reads; ++r) out_seeds[r] = seeds[r];
  return LTR_OK;
}
