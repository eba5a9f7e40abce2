// vcf_writer.cpp -- ltr_vcf_record: the VCF record of one genotyped locus, the way SeqStutterGenotyper::write_vcf_record
// composes it (reference src/seq_stutter_genotyper.cpp:894-1402) under the reference's output switches (ltr_vcf_record: the
// defaults, ALLREADS and MALLREADS on, src/genotyper.cpp:339-346; ltr_vcf_record_ex: any combination of ALLREADS / MALLREADS /
// GL / PL / PHASEDGL / FILTER; the haplotype fields HQ / PHQ have no command-line switch) on the long-read path
// (--stutter-align-len 0: no retraced alignments, DFLANKINDEL = 0, MALLREADS = size of the allele a read is assigned to).
//   get_alleles         :688-781   alleles of the record: trimmed to the region, flanks re-attached, 1 bp pad when an
//                                   alternate allele would otherwise start differently from the reference allele
//   reorder_alleles     :667-686   alternate alleles by (length, sequence)
//   ExtractCigar        src/extract_indels.cpp:18-93   (the caller supplies its result per read: ltr_extract_cigar_bp_diff)
//   condense_read_counts  src/genotyper.h:50-64
// Inputs are the outputs of the library's own pipeline (ltr_candidate_alleles, ltr_genotyper_run with read alleles, the
// per-read bookkeeping of ltr_regions_run); ltr_vcf_header gives the header lines
// (Genotyper::get_vcf_header, src/genotyper.cpp:258-336).
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <new>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "longtr_b200.h"

namespace {

std::string fixed2(double v) {
  char buf[64];
  snprintf(buf, sizeof(buf), "%.2f", v);
  return buf;
}

std::string condense(const std::vector<int>& v) {
  if (v.empty()) return ".";
  std::map<int, int> counts;
  for (int x : v) ++counts[x];
  std::string out;
  for (auto it = counts.begin(); it != counts.end(); ++it) {
    if (it != counts.begin()) out += ";";
    out += std::to_string(it->first) + "|" + std::to_string(it->second);
  }
  return out;
}

bool by_length_and_sequence(const std::string& a, const std::string& b) {
  return a.size() != b.size() ? a.size() < b.size() : a.compare(b) < 0;
}

}  // namespace

// ExtractCigar (extract_indels.cpp:18-93): the size difference of a read to the reference between region_start and region_end
// from its CIGAR (BAM-encoded operations).  1 = *bp_diff set, 0 = the read does not tell.
extern "C" int ltr_extract_cigar_bp_diff(const uint32_t* cigar_ops, uint32_t n_ops, int32_t cigar_start, int32_t region_start,
                                         int32_t region_end, int32_t* bp_diff) {
  if (!bp_diff || (n_ops && !cigar_ops) || n_ops == 0) return 0;
  auto type = [&](size_t i) { return "MIDNSHP=X????????"[cigar_ops[i] & 15]; };
  auto len = [&](size_t i) { return (int)(cigar_ops[i] >> 4); };
  auto is_match = [](char t) { return t == 'M' || t == '=' || t == 'X'; };
  int pos = cigar_start, region_len = 0;
  for (size_t i = 0; i < n_ops; ++i)
    if (is_match(type(i)) || type(i) == 'D') region_len += len(i);
  if (region_start < cigar_start) return 0;
  if (region_end >= cigar_start + region_len) return 0;
  size_t first = 0, last_match = 0;
  while (pos < region_start && first < n_ops) {
    const char t = type(first);
    if (is_match(t) || t == 'D') pos += len(first);
    if (is_match(t)) last_match = first;
    ++first;
  }
  first = last_match;
  if (first == 0 && !is_match(type(0))) return 0;
  size_t end = n_ops - 1;
  last_match = n_ops - 1;
  pos = cigar_start + region_len;
  while (pos > region_end) {
    const char t = type(end);
    if (is_match(t) || t == 'D') pos -= len(end);
    if (is_match(t)) last_match = end;
    if (end == 0) break;
    --end;
  }
  end = last_match;
  if (end == n_ops - 1 && !is_match(type(end))) return 0;
  int d = 0;
  for (size_t i = first; i <= end; ++i) {
    if (type(i) == 'D') d -= len(i);
    else if (type(i) == 'I') d += len(i);
  }
  *bp_diff = d;
  return 1;
}

extern "C" int ltr_vcf_record(const ltr_vcf_locus* L, char* out, uint32_t capacity, uint32_t* out_len) {
  return ltr_vcf_record_ex(L, nullptr, out, capacity, out_len);
}

static int vcf_record_impl(const ltr_vcf_locus* L, const ltr_vcf_extras* X, char* out, uint32_t capacity, uint32_t* out_len);

// C ABI boundary: no exception leaves the library (exhausted memory becomes an error code)
extern "C" int ltr_vcf_record_ex(const ltr_vcf_locus* L, const ltr_vcf_extras* X, char* out, uint32_t capacity,
                                 uint32_t* out_len) {
  try {
    return vcf_record_impl(L, X, out, capacity, out_len);
  } catch (const std::bad_alloc&) {
    return LTR_ERR_OOM;
  } catch (...) {
    return LTR_ERR_INVALID;
  }
}

static int vcf_record_impl(const ltr_vcf_locus* L, const ltr_vcf_extras* X, char* out, uint32_t capacity, uint32_t* out_len) {
  const uint32_t sw = X ? X->switches : LTR_VCF_DEFAULT;
  if (sw & ~(LTR_VCF_ALLREADS | LTR_VCF_MALLREADS | LTR_VCF_GLS | LTR_VCF_PLS | LTR_VCF_PHASED_GLS | LTR_VCF_FILTERS))
    return LTR_ERR_INVALID;
  if (!L || !out_len || (capacity && !out) || !L->chrom || !L->motif || !L->chrom_seq || L->n_alleles < 1 || !L->allele_off ||
      !L->allele_bytes || L->n_samples < 0 || (L->n_samples && (!L->gts || !L->log_unphased_posteriors || !L->gl_diffs)) ||
      L->n_reads < 0 || (L->n_reads && (!L->read_sample || !L->log_p1 || !L->log_p2)) || (L->n_columns && !L->column_sample))
    return LTR_ERR_INVALID;
  const bool haploid = L->haploid != 0;
  if (!haploid && L->n_samples && !L->log_phased_posteriors) return LTR_ERR_INVALID;
  const bool show_gl = (sw & LTR_VCF_GLS) != 0, show_pl = (sw & LTR_VCF_PLS) != 0;
  const bool show_pgl = !haploid && (sw & LTR_VCF_PHASED_GLS) != 0, show_filter = (sw & LTR_VCF_FILTERS) != 0;
  if (L->n_samples && ((show_gl && (!X->gl_begin || !X->gls)) || (show_pl && (!X->gl_begin || !X->pls)) ||
                       (show_pgl && (!X->pgl_begin || !X->phased_gls))))
    return LTR_ERR_INVALID;
  auto ref_sub = [&](int64_t a, int64_t b) {  // uppercase(chrom_seq.substr(a, b - a))
    std::string s;
    for (int64_t p = a; p < b; ++p) {
      char c = (p >= L->chrom_seq_start && p < L->chrom_seq_start + L->chrom_seq_len) ? (char)L->chrom_seq[p - L->chrom_seq_start] : 'N';
      if (c >= 'a' && c <= 'z') c = (char)(c - 32);
      s.push_back(c);
    }
    return s;
  };
  // the block after the removal of uncalled alleles: the surviving candidates, in candidate order
  std::vector<int> cand_of;       // surviving allele -> candidate index
  std::vector<int> kept_of((size_t)L->n_alleles, -1);
  for (int a = 0; a < L->n_alleles; ++a)
    if (!L->kept_mask || L->kept_mask[a] || a == 0) {
      kept_of[(size_t)a] = (int)cand_of.size();
      cand_of.push_back(a);
    }
  std::vector<std::string> alleles;
  std::vector<bool> inexact;
  for (int a : cand_of) {
    alleles.push_back(std::string((const char*)L->allele_bytes + L->allele_off[a], L->allele_off[a + 1] - L->allele_off[a]));
    if (alleles.back().empty()) return LTR_ERR_UNSUPPORTED;  // "<DEL>" alleles (a read in which the repeat is deleted)
    inexact.push_back(L->allele_inexact && L->allele_inexact[a] && a != 0);
  }
  const int K = (int)alleles.size();
  // ---- get_alleles (:688-781) ----
  int32_t left_trim = 0, start = L->block_start;
  while (start + left_trim < L->region_start) {
    bool trim = true;
    for (const std::string& s : alleles)
      if ((size_t)(left_trim + 1) >= s.size() || s[(size_t)left_trim] != alleles[0][(size_t)left_trim]) {
        trim = false;
        break;
      }
    if (!trim) break;
    ++left_trim;
  }
  start += left_trim;
  for (std::string& s : alleles) s = s.substr((size_t)left_trim);
  int32_t right_trim = 0, end = L->block_end;
  while (end - right_trim > L->region_stop) {
    bool trim = true;
    const size_t ref_size = alleles[0].size();
    for (const std::string& s : alleles)
      if ((size_t)(right_trim + 1) >= s.size() || s[s.size() - (size_t)right_trim - 1] != alleles[0][ref_size - (size_t)right_trim - 1]) {
        trim = false;
        break;
      }
    if (!trim) break;
    ++right_trim;
  }
  end -= right_trim;
  for (std::string& s : alleles) s = s.substr(0, s.size() - (size_t)right_trim);
  std::string left_flank = start >= L->region_start ? ref_sub(L->region_start, start) : "";
  const std::string right_flank = end <= L->region_stop ? ref_sub(end, L->region_stop) : "";
  int32_t pos = std::min(L->region_start, start);
  if (left_flank.empty()) {
    bool pad_left = false;
    for (int i = 1; i < K; ++i)
      if (alleles[(size_t)i].empty() || alleles[(size_t)i][0] != alleles[0][0]) {
        pad_left = true;
        break;
      }
    if (pad_left) {
      pos -= 1;
      left_flank = ref_sub(pos, pos + 1);
    }
  }
  for (std::string& s : alleles) s = left_flank + s + right_flank;
  pos += 1;
  std::vector<int> bp_diffs;
  for (const std::string& s : alleles) bp_diffs.push_back((int)s.size() - (int)alleles[0].size());
  // ---- reorder_alleles (:667-686) ----
  std::vector<int> old_to_new((size_t)K, -1), new_to_old;
  {
    std::map<std::string, int> old_index;
    for (int i = 0; i < K; ++i) old_index[alleles[(size_t)i]] = i;
    std::vector<std::string> sorted = alleles;
    std::sort(sorted.begin() + 1, sorted.end(), by_length_and_sequence);
    for (int i = 0; i < K; ++i) {
      const int o = old_index[sorted[(size_t)i]];
      new_to_old.push_back(o);
      old_to_new[(size_t)o] = i;
    }
  }
  // ---- per read (:938-1042) ----
  const int S = L->n_samples;
  std::vector<int> n_aligned((size_t)S, 0), n_snps((size_t)S, 0), strand_one((size_t)S, 0), strand_two((size_t)S, 0);
  std::vector<std::vector<int> > bps((size_t)S), ml_bps((size_t)S);
  for (int r = 0; r < L->n_reads; ++r) {
    const int s = L->read_sample[r];
    if (s < 0 || s >= S) return LTR_ERR_INVALID;
    ++n_aligned[(size_t)s];
    if (fabs(L->log_p1[r] - L->log_p2[r]) > 1e-10) {  // TOLERANCE
      ++n_snps[(size_t)s];
      if (L->log_p1[r] > L->log_p2[r]) ++strand_one[(size_t)s];
      else ++strand_two[(size_t)s];
    }
    if (L->read_bp_diff && L->read_bp_diff[r] != INT_MIN) bps[(size_t)s].push_back(L->read_bp_diff[r]);
    int ga = L->gts[2 * s], gb = L->gts[2 * s + 1];
    if (ga < 0 || ga >= L->n_alleles || gb < 0 || gb >= L->n_alleles || kept_of[(size_t)ga] < 0 || kept_of[(size_t)gb] < 0)
      return LTR_ERR_INVALID;
    int best = ga;  // homozygous / haploid: the genotype's allele
    if (!haploid && ga != gb) {
      if (!L->read_allele) return LTR_ERR_INVALID;  // heterozygous calls need the per-read assignment
      best = L->read_allele[r];
      if (best != ga && best != gb) return LTR_ERR_INVALID;
    }
    ml_bps[(size_t)s].push_back(bp_diffs[(size_t)kept_of[(size_t)best]]);
  }
  // ---- allele counts over the samples with reads (:1044-1072) ----
  std::vector<int> allele_counts((size_t)K, 0);
  int allele_number = 0;
  for (int s = 0; s < S; ++s) {
    if (n_aligned[(size_t)s] == 0) continue;
    const int a = kept_of[(size_t)L->gts[2 * s]], b = kept_of[(size_t)L->gts[2 * s + 1]];
    if (haploid) {
      ++allele_counts[(size_t)a];
      ++allele_number;
    } else {
      ++allele_counts[(size_t)a];
      ++allele_counts[(size_t)b];
      allele_number += 2;
    }
  }
  // ---- the record (:1090-1330) ----
  std::string o;
  o += L->chrom;
  o += "\t" + std::to_string(pos) + "\t" + ((L->name && L->name[0]) ? std::string(L->name) : std::string("."));
  o += "\t" + alleles[(size_t)new_to_old[0]] + "\t";
  if (K == 1) o += ".";
  else
    for (int i = 1; i < K; ++i) o += alleles[(size_t)new_to_old[(size_t)i]] + (i + 1 < K ? "," : "");
  o += "\t.\t.\t";
  std::string period_str;
  {
    std::stringstream ss(L->motif);
    std::string item;
    bool first = true;
    while (std::getline(ss, item, ',')) {
      period_str += (first ? "" : ",") + std::to_string(item.size());
      first = false;
    }
  }
  std::string inexact_str;
  if (K == 1) inexact_str = ".";
  else
    for (int i = 1; i < K; ++i) inexact_str += std::string(i > 1 ? "," : "") + (inexact[(size_t)new_to_old[(size_t)i]] ? "1" : "0");
  o += "START=" + std::to_string(L->region_start + 1) + ";END=" + std::to_string(L->region_stop) + ";MOTIF=" + L->motif +
       ";PERIOD=" + period_str + ";NSKIP=0;NFILT=0;INEXACT_ALLELE=" + inexact_str + ";";
  if (K > 1) {
    o += "BPDIFFS=";
    for (int i = 1; i < K; ++i) o += std::string(i > 1 ? "," : "") + std::to_string(bp_diffs[(size_t)new_to_old[(size_t)i]]);
    o += ";";
  }
  int tot_dp = 0, tot_dsnp = 0;
  std::set<int> shown;
  for (int c = 0; c < L->n_columns; ++c) {
    const int s = L->column_sample[c];
    if (s < 0) continue;
    if (s >= S) return LTR_ERR_INVALID;
    if (!shown.insert(s).second) continue;
    tot_dp += n_aligned[(size_t)s];
    tot_dsnp += n_snps[(size_t)s];
  }
  o += "DP=" + std::to_string(tot_dp) + ";DSNP=" + std::to_string(tot_dsnp) + ";DFLANKINDEL=0;";
  o += "AN=" + std::to_string(allele_number) + ";REFAC=" + std::to_string(allele_counts[0]);
  if (K > 1) {
    o += ";AC=";
    for (int i = 1; i < K; ++i) o += std::to_string(allele_counts[(size_t)new_to_old[(size_t)i]]) + (i + 1 < K ? "," : "");
  }
  o += haploid ? "\tGT:GB:Q:DP:DFLANKINDEL:GLDIFF" : "\tGT:GB:Q:PQ:DP:DSNP:DFLANKINDEL:PDP:PSNP:GLDIFF";
  int n_fields = haploid ? 6 : 10;  // :1172-1194; FILTER is not counted
  const struct {
    bool on;
    const char* key;
  } optional[] = {{(sw & LTR_VCF_ALLREADS) != 0, ":ALLREADS"}, {(sw & LTR_VCF_MALLREADS) != 0, ":MALLREADS"}, {show_gl, ":GL"},
                  {show_pl, ":PL"}, {show_pgl, ":PHASEDGL"}};
  for (const auto& f : optional)
    if (f.on) {
      o += f.key;
      ++n_fields;
    }
  if (show_filter) o += ":FILTER";
  std::string no_reads = ".";  // a column without (realigned) reads (:1203-1215)
  if (show_filter) {
    no_reads.clear();
    for (int f = 0; f < n_fields; ++f) no_reads += ".:";
    no_reads += "NO_READS";
  }
  // GL / PL / PHASEDGL slices hold the kept alleles in candidate order; the record lists them in its own allele order
  // (:1311-1358): pairs (j <= i) of the re-ordered alleles, phased pairs [i][j]
  auto gl_list = [&](int s, auto&& value_at) {
    std::string t;
    const uint64_t g0 = X->gl_begin[s];
    if (haploid) {
      for (int i = 0; i < K; ++i) t += (i ? "," : "") + value_at(g0 + (uint64_t)new_to_old[(size_t)i]);
    } else {
      for (int i = 0; i < K; ++i)
        for (int j = 0; j <= i; ++j) {
          const int lo = std::min(new_to_old[(size_t)i], new_to_old[(size_t)j]), hi = std::max(new_to_old[(size_t)i], new_to_old[(size_t)j]);
          t += ((i || j) ? "," : "") + value_at(g0 + (uint64_t)hi * (uint64_t)(hi + 1) / 2 + (uint64_t)lo);
        }
    }
    return t;
  };
  if (show_gl || show_pl)
    for (int s = 0; s < S; ++s)
      if (X->gl_begin[s + 1] < X->gl_begin[s] || X->gl_begin[s + 1] - X->gl_begin[s] < (uint64_t)(haploid ? K : K * (K + 1) / 2))
        return LTR_ERR_INVALID;
  if (show_pgl)
    for (int s = 0; s < S; ++s)
      if (X->pgl_begin[s + 1] < X->pgl_begin[s] || X->pgl_begin[s + 1] - X->pgl_begin[s] < (uint64_t)K * (uint64_t)K)
        return LTR_ERR_INVALID;
  for (int c = 0; c < L->n_columns; ++c) {
    o += "\t";
    const int s = L->column_sample[c];
    if (s < 0 || n_aligned[(size_t)s] == 0) {
      o += no_reads;
      continue;
    }
    const int a = kept_of[(size_t)L->gts[2 * s]], b = kept_of[(size_t)L->gts[2 * s + 1]];
    const std::string gldiff = K == 1 ? std::string(".") : fixed2(L->gl_diffs[s]);
    if (!haploid) {
      o += std::to_string(old_to_new[(size_t)a]) + "|" + std::to_string(old_to_new[(size_t)b]);
      o += ":" + std::to_string(bp_diffs[(size_t)a]) + "|" + std::to_string(bp_diffs[(size_t)b]);
      o += ":" + fixed2(exp(L->log_unphased_posteriors[s])) + ":" + fixed2(exp(L->log_phased_posteriors[s]));
      o += ":" + std::to_string(n_aligned[(size_t)s]) + ":" + std::to_string(n_snps[(size_t)s]) + ":0";
      o += ":" + std::to_string(L->n_p1 ? L->n_p1[s] : 0) + "|" + std::to_string(L->n_p2 ? L->n_p2[s] : 0);
      o += ":" + std::to_string(strand_one[(size_t)s]) + "|" + std::to_string(strand_two[(size_t)s]);
      o += ":" + gldiff;
    } else {
      o += std::to_string(old_to_new[(size_t)a]) + ":" + std::to_string(bp_diffs[(size_t)a]);
      o += ":" + fixed2(exp(L->log_unphased_posteriors[s])) + ":" + std::to_string(n_aligned[(size_t)s]) + ":0:" + gldiff;
    }
    if (sw & LTR_VCF_ALLREADS) o += ":" + condense(bps[(size_t)s]);
    if (sw & LTR_VCF_MALLREADS) o += ":" + condense(ml_bps[(size_t)s]);
    if (show_gl) o += ":" + gl_list(s, [&](uint64_t k) { return fixed2(X->gls[k]); });
    if (show_pl) o += ":" + gl_list(s, [&](uint64_t k) { return std::to_string(X->pls[k]); });
    if (show_pgl) {
      o += ":";
      const uint64_t g0 = X->pgl_begin[s];
      for (int i = 0; i < K; ++i)
        for (int j = 0; j < K; ++j)
          o += std::string((i || j) ? "," : "") +
               fixed2(X->phased_gls[g0 + (uint64_t)new_to_old[(size_t)i] * (uint64_t)K + (uint64_t)new_to_old[(size_t)j]]);
    }
    if (show_filter) o += ":PASS";
  }
  *out_len = (uint32_t)o.size();
  if (o.size() + 1 > capacity) return LTR_ERR_INVALID;
  memcpy(out, o.c_str(), o.size() + 1);
  return LTR_OK;
}

// ---- header lines (Genotyper::get_vcf_header, src/genotyper.cpp:258-336; contigs: FastaReader::write_all_contigs_to_vcf,
//      src/fasta_reader.cpp:64-73) with the default output switches.  The descriptions are part of the file format LongTR's
//      users parse, so they are reproduced verbatim, as a table.
namespace {
struct FieldLine {
  const char* id;
  const char* number;
  const char* type;
  const char* description;
};
const FieldLine kInfo[] = {
    {"START", "1", "Integer", "Inclusive start coodinate for the repetitive portion of the reference allele"},
    {"END", "1", "Integer", "Inclusive end coordinate for the repetitive portion of the reference allele"},
    {"MOTIF", ".", "String", "TR motif(s)"},
    {"PERIOD", ".", "Integer", "Length of TR motif(s)"},
    {"NSKIP", "1", "Integer", "Number of samples not genotyped due to various issues"},
    {"NFILT", "1", "Integer", "Number of samples whose genotypes were filtered due to various issues"},
    {"INEXACT_ALLELE", "A", "Integer",
     "Boolean showing if each alternate allele is exact or approximated by POA, 0 for exact 1 for approximated."},
    {"BPDIFFS", "A", "Integer", "Base pair difference of each alternate allele from the reference allele"},
    {"DP", "1", "Integer", "Total number of valid reads used to genotype all samples"},
    {"DSNP", "1", "Integer", "Total number of reads with SNP phasing information"},
    {"DFLANKINDEL", "1", "Integer", "Total number of reads with an indel in the regions flanking the STR"},
    {"AN", "1", "Integer", "Total number of alleles in called genotypes"},
    {"REFAC", "1", "Integer", "Reference allele count"},
    {"AC", "A", "Integer", "Alternate allele counts"},
};
const FieldLine kFormat[] = {
    {"GT", "1", "String", "Genotype"},
    {"GB", "1", "String", "Base pair differences of genotype from reference"},
    {"Q", "1", "Float", "Posterior probability of unphased genotype"},
    {"PQ", "1", "Float", "Posterior probability of phased genotype"},
    {"DP", "1", "Integer", "Number of valid reads used for sample's genotype"},
    {"DSNP", "1", "Integer", "Number of reads with SNP phasing information"},
    {"PSNP", "1", "String", "Number of reads with SNPs supporting each haploid genotype"},
    {"PDP", "1", "String", "Fractional reads supporting each haploid genotype"},
    {"GLDIFF", "1", "Float", "Difference in likelihood between the reported and next best genotypes"},
};
struct SwitchedLine {
  uint32_t flag;
  FieldLine line;
};
const SwitchedLine kSwitched[] = {  // in the order get_vcf_header writes them (:313-327)
    {LTR_VCF_ALLREADS, {"ALLREADS", "1", "String", "Base pair difference observed in each read's Needleman-Wunsch alignment"}},
    {LTR_VCF_MALLREADS,
     {"MALLREADS", "1", "String",
      "Maximum likelihood bp diff in each read based on haplotype alignments for reads that span the repeat region by at least 5 "
      "base pairs"}},
    {LTR_VCF_GLS, {"GL", "G", "Float", "log10 genotype likelihoods"}},
    {LTR_VCF_PLS, {"PL", "G", "Integer", "Phred-scaled genotype likelihoods"}},
    {LTR_VCF_PHASED_GLS,
     {"PHASEDGL", ".", "Float",
      "log10 genotype likelihood for each phased genotype. Value for phased genotype X|Y is stored at a 0-based index of X*A + Y, "
      "where A is the number of alleles. Not applicable to haploid genotypes"}},
    {LTR_VCF_FILTERS, {"FILTER", "1", "String", "Reason for filtering the current call, or PASS if the call was not filtered"}},
};
}  // namespace

extern "C" int ltr_vcf_header(const ltr_fasta* fasta, const char* fasta_path, const char* command,
                              const char* const* sample_names, uint32_t n_samples, char* out, uint32_t capacity,
                              uint32_t* out_len) {
  return ltr_vcf_header_ex(fasta, fasta_path, command, sample_names, n_samples, LTR_VCF_DEFAULT, out, capacity, out_len);
}

extern "C" int ltr_vcf_header_ex(const ltr_fasta* fasta, const char* fasta_path, const char* command,
                                 const char* const* sample_names, uint32_t n_samples, uint32_t switches, char* out,
                                 uint32_t capacity, uint32_t* out_len) {
  if (!fasta || !fasta_path || !command || !out_len || (n_samples && !sample_names) || (capacity && !out)) return LTR_ERR_INVALID;
  std::string o = "##fileformat=VCFv4.1\n";
  o += std::string("##command=") + command + "\n##reference=" + fasta_path + "\n";
  for (int32_t i = 0; i < ltr_fasta_n_seqs(fasta); ++i) {
    const char* name = ltr_fasta_seq_name(fasta, i);
    o += std::string("##contig=<ID=") + name + ",length=" + std::to_string(ltr_fasta_seq_len(fasta, name)) + ">\n";
  }
  for (const FieldLine& f : kInfo)
    o += std::string("##INFO=<ID=") + f.id + ",Number=" + f.number + ",Type=" + f.type + ",Description=\"" + f.description + "\">\n";
  for (const FieldLine& f : kFormat)
    o += std::string("##FORMAT=<ID=") + f.id + ",Number=" + f.number + ",Type=" + f.type + ",Description=\"" + f.description + "\">\n";
  for (const SwitchedLine& w : kSwitched)
    if (switches & w.flag)
      o += std::string("##FORMAT=<ID=") + w.line.id + ",Number=" + w.line.number + ",Type=" + w.line.type + ",Description=\"" +
           w.line.description + "\">\n";
  o += "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT";
  for (uint32_t i = 0; i < n_samples; ++i) o += std::string("\t") + sample_names[i];
  o += "\n";
  *out_len = (uint32_t)o.size();
  if (o.size() + 1 > capacity) return LTR_ERR_INVALID;
  memcpy(out, o.c_str(), o.size() + 1);
  return LTR_OK;
}
