// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product.
//
// Runs ONE locus through the UNMODIFIED reference's per-locus genotyper without any file IO:
//   SeqStutterGenotyper(...)  ->  genotype(1000, 4, 0.01, log)  ->  write_vcf_record(...)
// (reference src/seq_stutter_genotyper.h:148-207; call sequence of GenotyperBamProcessor::analyze_reads_and_phasing,
// src/genotyper_bam_processor.cpp:286-305) and returns the VCF record text the reference would write.  The
// "bgzipped" VCF stream is captured in memory through the bgzf_* stand-ins below.
// Built twice by oracle/build_ref.sh:
//   ltr_ref_full   every object is the reference's own           -> golden VCF records
//   ltr_ref_gpu    same objects, but HapAligner::process_reads and Genotyper::calc_log_sample_posteriors
//                        come from integration/reference_binding.cpp (C ABI -> GPU)   -> drop-in check
//   ltr_ref_trace  the reference's objects with a recording wrapper around calc_log_sample_posteriors
//                        (-DLTR_TRACE_POSTERIORS): LL matrix, posteriors and MAP pairs of both passes of genotype()
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#ifdef LTR_TRACE_POSTERIORS
// The trace build looks inside the reference's genotyper objects (layout is unaffected by access specifiers).
#include <iostream>
#include <map>
#include <set>
#define private public
#define protected public
#endif
#include "SeqAlignment/AlignmentData.h"
#include "mathops.h"
#include "region.h"
#include "seq_stutter_genotyper.h"
#include "stutter_model.h"
#include "vcf_writer.h"

#include "full_locus.h"

// ---- in-memory stand-ins for the htslib symbols this path links against ---------------------------------------
struct BGZF { int dummy; };
static std::string g_capture;
extern "C" {
BGZF* bgzf_open(const char*, const char*) { return new BGZF(); }
int bgzf_close(BGZF* fp) { delete fp; return 0; }
ssize_t bgzf_write(BGZF*, const void* data, size_t length) { g_capture.append((const char*)data, length); return (ssize_t)length; }
int bgzf_getc(BGZF*) { return -1; }
double kt_fisher_exact(int, int, int, int, double* l, double* r, double* two) {  // result never printed (:1168)
  if (l) *l = 1; if (r) *r = 1; if (two) *two = 1;
  return 1.0;
}
}
// Symbols that only --ref-vcf (vcf_input.cpp / vcf_reader.h) and BAM parsing (extract_indels.cpp) would reach:
// never executed by the IO-less locus run; reaching one is a driver bug -> abort loudly.
#include "bam_io.h"
#include "vcf_reader.h"
static void unreachable_full(const char* fn) {
  std::fprintf(stderr, "oracle/_ref full driver: %s reached -- not part of the IO-less locus run\n", fn);
  std::abort();
}
extern "C" {
int bcf_get_format_values(const bcf_hdr_t*, bcf1_t*, const char*, void**, int*, int) { unreachable_full("bcf_get_format_values"); return 0; }
bcf_info_t* bcf_get_info(const bcf_hdr_t*, bcf1_t*, const char*) { unreachable_full("bcf_get_info"); return 0; }
int bcf_get_info_values(const bcf_hdr_t*, bcf1_t*, const char*, void**, int*, int) { unreachable_full("bcf_get_info_values"); return 0; }
void tbx_itr_destroy(hts_itr_t*) { unreachable_full("tbx_itr_destroy"); }
hts_itr_t* tbx_itr_querys(tbx_t*, const char*) { unreachable_full("tbx_itr_querys"); return 0; }
}
void BamAlignment::ExtractSequenceFields() { unreachable_full("BamAlignment::ExtractSequenceFields"); }
bool VCF::VCFReader::get_next_variant(VCF::Variant&) { unreachable_full("VCFReader::get_next_variant"); return false; }
const std::vector<std::string>& VCF::Variant::get_samples() const { unreachable_full("Variant::get_samples"); static std::vector<std::string> v; return v; }

#ifdef LTR_TRACE_POSTERIORS
// ltr_ref_trace: oracle/build_ref.sh renames the reference's Genotyper::calc_log_sample_posteriors(std::vector<int>&)
// (src/genotyper.cpp:45-83) to ltr_orig_calc_log_sample_posteriors in a COPY of its object file; the definition below
// takes its place, records what goes in and what comes out of every call, and runs the original in between.  genotype()
// calls it twice (src/seq_stutter_genotyper.cpp:635 and, through remove_alleles, :643): before and after the uncalled
// alleles are dropped.
extern "C" double ltr_orig_calc_log_sample_posteriors(Genotyper* self, std::vector<int>& read_weights);
static std::string g_trace;
static void trace_doubles(std::ostringstream& o, const char* tag, const double* v, size_t n) {
  o << tag << " " << n;
  char buf[64];
  for (size_t i = 0; i < n; ++i) {
    std::snprintf(buf, sizeof(buf), " %a", v[i]);
    o << buf;
  }
  o << "\n";
}
double Genotyper::calc_log_sample_posteriors(std::vector<int>& read_weights) {
  std::ostringstream o;
  SeqStutterGenotyper* g = dynamic_cast<SeqStutterGenotyper*>(this);
  o << "CALL " << num_alleles_ << " " << num_reads_ << " " << num_samples_ << "\n";
  o << "ALLELES";
  if (g != NULL && g->haplotype_ != NULL && g->haplotype_->num_blocks() == 3)
    for (int a = 0; a < g->hap_blocks_[1]->num_options(); ++a) o << " " << (g->hap_blocks_[1]->get_seq(a).empty() ? "-" : g->hap_blocks_[1]->get_seq(a));
  o << "\n";
  if (g != NULL && g->haplotype_ != NULL && g->haplotype_->num_blocks() == 3)  // flank blocks and repeat coordinates
    o << "BLOCKS " << g->hap_blocks_[1]->start() << " " << g->hap_blocks_[1]->end() << " "
      << g->hap_blocks_[0]->get_seq(0) << " " << g->hap_blocks_[2]->get_seq(0) << "\n";
  o << "SEEDS";
  if (g != NULL && g->seed_positions_ != NULL)
    for (int r = 0; r < num_reads_; ++r) o << " " << g->seed_positions_[r];
  o << "\n";
  o << "LABELS";
  for (int r = 0; r < num_reads_; ++r) o << " " << sample_label_[r];
  o << "\n";
  trace_doubles(o, "LL", log_aln_probs_, (size_t)num_reads_ * num_alleles_);
  trace_doubles(o, "P1", log_p1_, (size_t)num_reads_);
  trace_doubles(o, "P2", log_p2_, (size_t)num_reads_);
  const double total = ltr_orig_calc_log_sample_posteriors(this, read_weights);
  trace_doubles(o, "POST", log_sample_posteriors_, (size_t)num_samples_ * num_alleles_ * num_alleles_);
  trace_doubles(o, "TOTALS", sample_total_LLs_, (size_t)num_samples_);
  std::vector<std::pair<int, int> > gts;
  get_optimal_haplotypes(gts);
  o << "GTS";
  for (size_t s = 0; s < gts.size(); ++s) o << " " << gts[s].first << " " << gts[s].second;
  o << "\n";
  g_trace += o.str();
  return total;
}
#endif

static void parse_cigar(const char* cigar, Alignment& aln) {
  int num = 0;
  for (const char* p = cigar; *p; ++p) {
    if (*p >= '0' && *p <= '9') num = num * 10 + (*p - '0');
    else { aln.add_cigar_element(CigarElement(*p, num)); num = 0; }
  }
}

extern "C" int32_t ltr_ref_full_locus(const ltr_full_locus* L, char* out, int32_t cap) {
  static bool logs_ready = false;
  if (!logs_ready) { precompute_integer_logs(); logs_ready = true; }
  g_capture.clear();
#ifdef LTR_TRACE_POSTERIORS
  g_trace.clear();
#endif
  const std::string chrom_seq(L->chrom_seq);
  Region region(L->chrom_name, L->region_start, L->region_stop, L->motif, L->region_name);
  RegionGroup rg(region);
  std::vector<std::string> samples;
  for (int s = 0; s < L->n_samples; ++s) samples.push_back(L->sample_names[s]);
  std::vector<Alignment> alns;
  std::vector<std::vector<double> > p1(L->n_samples), p2(L->n_samples);
  std::vector<bool> use_for_haps(1, true);
  for (int r = 0; r < L->n_reads; ++r) {
    const ltr_full_read& fr = L->reads[r];
    Alignment aln(fr.start, fr.stop, fr.rev_strand != 0, false, fr.name, fr.qual, fr.seq, fr.aln);
    parse_cigar(fr.cigar, aln);
    aln.set_hap_gen_info(use_for_haps);
    alns.push_back(aln);
    p1[fr.sample].push_back(fr.log_p1);
    p2[fr.sample].push_back(fr.log_p2);
  }
  std::vector<int> n1(L->n_p1s, L->n_p1s + L->n_samples), n2(L->n_p2s, L->n_p2s + L->n_samples);
  StutterModel sm(L->stutter[0], L->stutter[1], L->stutter[2], L->stutter[3], L->stutter[4], L->stutter[5],
                  std::string(L->stutter_motif));
  sm.set_period(L->stutter_period);
  std::vector<StutterModel*> models(1, &sm);
  std::vector<float> params(L->aln_params, L->aln_params + L->n_aln_params);
  std::ostringstream log, html;
  int32_t len = 0;
  {
    SeqStutterGenotyper g(rg, L->haploid != 0, 1, alns, p1, p2, n1, n2, samples, chrom_seq, models, NULL, log, true,
                          L->indel_flank_len, L->switch_old_align_len, params);
    if (g.genotype(1000, 4, 0.01, log)) {
      VCFWriter writer;
      writer.open("in-memory");
      g.write_vcf_record(samples, chrom_seq, false, false, html, &writer, log);
      writer.close();
      len = (int32_t)g_capture.size();
    }
  }
  if (len + 1 > cap) return -1;
  std::memcpy(out, g_capture.c_str(), (size_t)len + 1);
  return len;
}

#ifdef LTR_FULL_MAIN
// Stand-alone form (the statically linked libstdc++ of this toolchain does not survive being dlopen()ed into
// Python together with iostream use, so the full-locus runs are executables driven over stdin/stdout):
//   per case, whitespace separated:
//     chrom_name chrom_seq region_start region_stop motif region_name
//     n_samples name... n_p1... n_p2...
//     stutter[6] stutter_motif stutter_period haploid indel_flank_len switch n_params params...
//     n_reads, then per read: start stop rev sample name seq qual aln cigar log_p1 log_p2
//   answer per case: "RECORD <len>\n<text>\n"
#include <iostream>
int main() {
  std::string chrom_name, chrom_seq, motif, region_name;
  static char out[1 << 20];
  // LTR_REF_OUTPUT_SWITCHES: the reference's output switches as a bit mask in the order of the library's LTR_VCF_* flags
  // (1 ALLREADS, 2 MALLREADS, 4 GL, 8 PL, 16 PHASEDGL, 32 FILTER): what --hide-allreads / --hide-mallreads / --output-gls /
  // --output-pls / --output-phased-gls / --output-filters set (src/hipstr_main.cpp:178-183).  Unset = the defaults.
  if (const char* sw = std::getenv("LTR_REF_OUTPUT_SWITCHES")) {
    const unsigned m = (unsigned)std::strtoul(sw, NULL, 0);
    Genotyper::OUTPUT_ALLREADS = (m & 1) ? 1 : 0;
    Genotyper::OUTPUT_MALLREADS = (m & 2) ? 1 : 0;
    Genotyper::OUTPUT_GLS = (m & 4) ? 1 : 0;
    Genotyper::OUTPUT_PLS = (m & 8) ? 1 : 0;
    Genotyper::OUTPUT_PHASED_GLS = (m & 16) ? 1 : 0;
    Genotyper::OUTPUT_FILTERS = (m & 32) ? 1 : 0;
  }
  while (std::cin >> chrom_name >> chrom_seq) {
    ltr_full_locus L;
    std::memset(&L, 0, sizeof(L));
    std::cin >> L.region_start >> L.region_stop >> motif >> region_name >> L.n_samples;
    std::vector<std::string> names(L.n_samples);
    std::vector<const char*> name_ptrs(L.n_samples);
    std::vector<int32_t> n1(L.n_samples), n2(L.n_samples);
    for (int s = 0; s < L.n_samples; ++s) std::cin >> names[s];
    for (int s = 0; s < L.n_samples; ++s) std::cin >> n1[s];
    for (int s = 0; s < L.n_samples; ++s) std::cin >> n2[s];
    for (int s = 0; s < L.n_samples; ++s) name_ptrs[s] = names[s].c_str();
    std::string stutter_motif;
    for (int i = 0; i < 6; ++i) std::cin >> L.stutter[i];
    std::cin >> stutter_motif >> L.stutter_period >> L.haploid >> L.indel_flank_len >> L.switch_old_align_len >> L.n_aln_params;
    for (int i = 0; i < L.n_aln_params; ++i) std::cin >> L.aln_params[i];
    std::cin >> L.n_reads;
    std::vector<ltr_full_read> reads(L.n_reads);
    std::vector<std::string> strs((size_t)L.n_reads * 5);
    for (int r = 0; r < L.n_reads; ++r) {
      ltr_full_read& fr = reads[r];
      std::cin >> fr.start >> fr.stop >> fr.rev_strand >> fr.sample;
      for (int k = 0; k < 5; ++k) std::cin >> strs[(size_t)r * 5 + k];
      std::cin >> fr.log_p1 >> fr.log_p2;
      fr.name = strs[(size_t)r * 5].c_str(); fr.seq = strs[(size_t)r * 5 + 1].c_str(); fr.qual = strs[(size_t)r * 5 + 2].c_str();
      fr.aln = strs[(size_t)r * 5 + 3].c_str(); fr.cigar = strs[(size_t)r * 5 + 4].c_str();
    }
    if (!std::cin) return 2;
    L.chrom_name = chrom_name.c_str(); L.chrom_seq = chrom_seq.c_str(); L.motif = motif.c_str();
    L.region_name = region_name.c_str(); L.sample_names = name_ptrs.data(); L.n_p1s = n1.data(); L.n_p2s = n2.data();
    L.reads = reads.data(); L.stutter_motif = stutter_motif.c_str();
    const int32_t n = ltr_ref_full_locus(&L, out, (int32_t)sizeof(out));
    if (n < 0) return 3;
    std::printf("RECORD %d\n%s\n", n, out);
#ifdef LTR_TRACE_POSTERIORS
    std::printf("TRACE %d\n%s", (int)g_trace.size(), g_trace.c_str());
#endif
    std::fflush(stdout);
  }
  return 0;
}
#endif
