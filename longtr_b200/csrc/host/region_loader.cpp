// region_loader.cpp -- ltr_region_collect: from BAM files and a BED region to the reads of one locus, the way LongTR's
// region loop prepares them for its genotyper.  Mirrors, for single-end (long) reads:
//   BamProcessor::process_regions        src/bam_processor.cpp:584-596   the window that is fetched (+- MAX_MATE_DIST)
//   BamProcessor::read_and_filter_reads  src/bam_processor.cpp:188-487   read filters, order of reads and samples
//   SNPBamProcessor::process_phased_reads  src/snp_bam_processor.cpp:141-232  phasing terms from the HP tag (--phased-bam)
//   GenotyperBamProcessor::left_align_reads  src/genotyper_bam_processor.cpp:38-168  spanning test, cut to +-200 bp
//                                        (BamAlignment::TrimAlignment, src/bam_io.cpp:267-372), '=XID' CIGAR rebuilt
//                                        against the reference sequence, soft-clipped reads dropped
// Paired-end bookkeeping (mate lookup, NO_UNIQUE_MAPPING between mates, PCR duplicates) is not reproduced: a read whose
// PAIRED flag is set is refused with LTR_ERR_UNSUPPORTED for the region.  Output arrays have the layout of ltr_locus_batch.
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <string.h>

#include <algorithm>
#include <new>
#include <map>
#include <string>
#include <vector>

#include "longtr_b200.h"

namespace {

struct Op {
  char type;
  int32_t len;
};

struct Aln {  // BamAlignment after ExtractSequenceFields (bam_io.cpp:21-54)
  std::string name, bases, quals;
  std::vector<Op> cigar;
  int32_t pos, end_pos;  // end_pos: one past the last reference base
  uint16_t flag;
  uint8_t mapq;
  int32_t hp;
  bool has_hp, has_xa, deleted;
  uint32_t file;
};

bool aux_has(const uint8_t* raw, size_t n, char t0, char t1) {
  if (n < 32) return false;
  const uint32_t l_name = raw[8], n_cig = raw[12] | (raw[13] << 8);
  const uint32_t l_seq = (uint32_t)raw[16] | ((uint32_t)raw[17] << 8) | ((uint32_t)raw[18] << 16) | ((uint32_t)raw[19] << 24);
  size_t q = 32 + (size_t)l_name + 4 * (size_t)n_cig + (l_seq + 1) / 2 + l_seq;
  while (q + 3 <= n) {
    const bool hit = raw[q] == (uint8_t)t0 && raw[q + 1] == (uint8_t)t1;
    const char ty = (char)raw[q + 2];
    q += 3;
    if (hit) return true;
    size_t len;
    switch (ty) {
      case 'A': case 'c': case 'C': len = 1; break;
      case 's': case 'S': len = 2; break;
      case 'i': case 'I': case 'f': len = 4; break;
      case 'd': len = 8; break;
      case 'Z': case 'H': {
        size_t e = q;
        while (e < n && raw[e]) ++e;
        len = e - q + 1;
        break;
      }
      case 'B': {
        if (q + 5 > n) return false;
        const char sub = (char)raw[q];
        uint32_t cnt;
        memcpy(&cnt, raw + q + 1, 4);
        len = 5 + (size_t)cnt * ((sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4);
        break;
      }
      default: return false;
    }
    q += len;
  }
  return false;
}

// BamAlignment::TrimAlignment(min_read_start, max_read_stop) (bam_io.cpp:267-372).  The reference walks base by base; here
// every CIGAR operation is consumed in one step (as many of its bases as the reference's loop would take before its
// condition fails), which leaves the same CIGAR, sequence and positions.
void trim_alignment(Aln& a, int32_t min_read_start, int32_t max_read_stop, int32_t flank_size) {
  int64_t ltrim = 0;
  int32_t start_pos = a.pos;
  size_t first = 0;  // cigar[first..) is what is left at the front
  std::vector<Op>& c = a.cigar;
  while (start_pos < min_read_start && first < c.size()) {  // :274-300
    Op& o = c[first];
    int32_t t = o.len;  // 'I', 'S', 'H', ...: the position does not move, the whole operation goes
    switch (o.type) {
      case 'M': case '=': case 'X':
        t = std::min(o.len, min_read_start - start_pos);
        ltrim += t;
        start_pos += t;
        break;
      case 'D':
        t = std::min(o.len, min_read_start - start_pos);
        start_pos += t;
        break;
      case 'I': case 'S': ltrim += t; break;
      default: break;
    }
    if (t == o.len) ++first;
    else o.len -= t;
  }
  c.erase(c.begin(), c.begin() + (long)first);
  // :303-338 -- is the repeat deleted in this read?
  {
    int32_t repeat_pointer = start_pos;
    const int32_t repeat_start = min_read_start + flank_size, repeat_end = max_read_stop - flank_size;
    int64_t deletion_size = 0;
    if (repeat_pointer >= min_read_start)
      for (size_t k = 0; k < c.size() && repeat_pointer < repeat_end; ++k) {
        const bool match = c[k].type == 'M' || c[k].type == '=' || c[k].type == 'X';
        if (!match && c[k].type != 'D') continue;
        const int32_t t = std::min(c[k].len, repeat_end - repeat_pointer);
        if (!match && repeat_pointer + t > repeat_start) deletion_size += repeat_pointer + t - std::max(repeat_pointer, repeat_start);
        repeat_pointer += t;
      }
    if (deletion_size >= repeat_end - repeat_start) a.deleted = true;
  }
  int64_t rtrim = 0;
  int32_t end_pos = a.end_pos;
  while (end_pos > max_read_stop && !c.empty()) {  // :342-364
    Op& o = c.back();
    int32_t t = o.len;
    switch (o.type) {
      case 'M': case '=': case 'X':
        t = std::min(o.len, end_pos - max_read_stop);
        rtrim += t;
        end_pos -= t;
        break;
      case 'D':
        t = std::min(o.len, end_pos - max_read_stop);
        end_pos -= t;
        break;
      case 'I': case 'S': rtrim += t; break;
      default: break;
    }
    if (t == o.len) c.pop_back();
    else o.len -= t;
  }
  a.bases = a.bases.substr((size_t)ltrim, a.bases.size() - (size_t)ltrim - (size_t)rtrim);
  a.quals = a.quals.substr((size_t)ltrim, a.quals.size() - (size_t)ltrim - (size_t)rtrim);
  a.pos = start_pos;
  a.end_pos = end_pos;
}

struct Owner {
  ltr_region_reads pub;
  std::vector<uint32_t> sample_file, sample_read_begin, read_off, cigar_off, cigar_ops, name_off;
  std::vector<int32_t> read_start, read_stop, read_sample, read_hp;
  std::vector<uint8_t> read_bytes, qual_bytes, hap_gen_ok, deleted;
  std::vector<double> log_p1, log_p2;
  std::vector<char> names;
};

uint32_t bam_op(char t) {
  switch (t) {
    case 'M': return 0; case 'I': return 1; case 'D': return 2; case 'N': return 3; case 'S': return 4;
    case 'H': return 5; case 'P': return 6; case '=': return 7; default: return 8;
  }
}

}  // namespace

extern "C" void ltr_region_params_default(ltr_region_params* p) {
  if (!p) return;
  p->max_mate_dist = 1000;   // bam_processor.h:83
  p->min_mean_qual = 30.0;   // bam_processor.h:95 (MIN_SUM_QUAL_LOG_PROB: the MEAN Phred quality, base_quality.h:77-84)
  p->min_mapq = 20.0;        // bam_processor.h:96
  p->require_spanning = 1;   // bam_processor.h:88
  p->min_flank = 5;          // bam_processor.h:85
  p->flank_size = 200;       // bam_io.h:28
  p->phased_bam = 1;
  p->check_hard_clips = 1;   // BASE_QUAL_TRIM > ' ' (bam_processor.h:101)
}

static int ltr_region_collect_impl(const ltr_bam* const* bams, int32_t n_bams, const char* chrom, int32_t start, int32_t stop,
                                  const uint8_t* ref_seq, int64_t ref_seq_start, int64_t ref_seq_len,
                                  const ltr_region_params* params, ltr_region_reads** out) {
  if (!bams || n_bams < 1 || !chrom || !ref_seq || !params || !out || stop < start) return LTR_ERR_INVALID;
  *out = nullptr;
  Owner* O = new Owner();
  memset(&O->pub, 0, sizeof(O->pub));
  ltr_region_reads& S = O->pub;
  // ---- read_and_filter_reads (single-end path) -----------------------------------------------------------------------
  std::map<std::string, Aln> potential_strs;  // keyed by "<file number>_<name>": the reference's order of reads
  std::map<std::string, bool> potential_mates;
  int rc = LTR_OK;
  for (int32_t f = 0; f < n_bams && rc == LTR_OK; ++f) {
    const int32_t tid = ltr_bam_ref_id(bams[f], chrom);
    if (tid < 0) { rc = LTR_ERR_INVALID; break; }
    const int32_t q0 = start < params->max_mate_dist ? 0 : start - params->max_mate_dist, q1 = stop + params->max_mate_dist;
    ltr_bam_reads* R = nullptr;
    rc = ltr_bam_fetch(bams[f], tid, q0, q1, 1, &R);
    if (rc != LTR_OK) break;
    potential_mates.clear();  // :246-249 new file
    const std::string label = std::to_string(f + 1) + "_";
    for (uint32_t i = 0; i < R->n; ++i) {
      // BamCramReader::GetNextAlignment stops at pos > end + 1 (bam_io.cpp:178); the iterator starts at records that
      // overlap [q0, q1)
      if (R->pos[i] > q1 + 1) break;
      const int32_t pos = R->pos[i], end_pos = R->end[i];
      const uint16_t flag = R->flag[i];
      if (pos > stop || end_pos < start) continue;  // :208-216 (single-end: nothing a mate could add)
      const uint32_t n_cig = R->cigar_off[i + 1] - R->cigar_off[i], l_seq = R->seq_off[i + 1] - R->seq_off[i];
      if ((flag & 4) || pos == 0 || n_cig == 0 || l_seq == 0) continue;  // :224-225
      if (flag & 1) { rc = LTR_ERR_UNSUPPORTED; break; }
      const bool overlaps = pos < stop && end_pos >= start;  // :229, :255
      if (!overlaps) continue;  // single-end reads outside the region only matter as mates
      const uint32_t* cg = R->cigar_ops + R->cigar_off[i];
      if (params->check_hard_clips && ((cg[0] & 15) == 5 || (cg[n_cig - 1] & 15) == 5)) {  // :231-237
        ++S.n_overlapping;
        ++S.n_hard_clipped;
        continue;
      }
      ++S.n_overlapping;
      const uint8_t* sq = R->seq + R->seq_off[i];
      const uint8_t* ql = R->qual + R->seq_off[i];
      bool pass = false;
      if (memchr(sq, 'N', l_seq)) ++S.n_has_n;  // :265-268
      else {
        // base_quality.h:77-84 adds (quality - 33) base by base in double: integer-valued partial sums far below 2^53, so the
        // integer sum converted once is the same double
        uint64_t qsum = 0;
        uint32_t k = 0;
#if defined(__SSE2__)
        __m128i acc = _mm_setzero_si128();
        for (; k + 16 <= l_seq; k += 16) acc = _mm_add_epi64(acc, _mm_sad_epu8(_mm_loadu_si128((const __m128i*)(ql + k)), _mm_setzero_si128()));
        qsum = (uint64_t)_mm_cvtsi128_si64(acc) + (uint64_t)_mm_cvtsi128_si64(_mm_unpackhi_epi64(acc, acc));
#endif
        for (; k < l_seq; ++k) qsum += ql[k];
        const double sum = (double)((int64_t)qsum - 33 * (int64_t)l_seq);
        if (sum / (double)l_seq < params->min_mean_qual) ++S.n_low_qual;
        else if ((double)R->mapq[i] < params->min_mapq) ++S.n_low_mapq;
        else if (params->require_spanning == 1 && !(pos <= start && end_pos >= stop)) ++S.n_not_spanning;  // :175-186
        else pass = true;
      }
      const char* nm = R->names + R->name_off[i];
      std::string name(nm);
      std::string key_name = name;
      if (key_name.size() > 2 && key_name[key_name.size() - 2] == '/') key_name.resize(key_name.size() - 2);  // :163-170
      const std::string key = label + key_name;
      if (!pass) {
        potential_mates.insert(std::make_pair(key, true));  // :381-383
        continue;
      }
      Aln a;
      a.name = name;
      a.bases.assign((const char*)sq, l_seq);
      a.quals.assign((const char*)ql, l_seq);
      a.cigar.resize(n_cig);
      for (uint32_t k = 0; k < n_cig; ++k) {
        a.cigar[k].type = "MIDNSHP=X"[(cg[k] & 15) > 8 ? 8 : (cg[k] & 15)];
        a.cigar[k].len = (int32_t)(cg[k] >> 4);
      }
      {
        // a record whose CIGAR does not describe its sequence and end position (a damaged file: the reference would read past
        // its strings) refuses the region instead
        int64_t q_len = 0, r_len = 0;
        bool ok = true;
        for (const Op& c : a.cigar) {
          if (c.len < 1) ok = false;
          if (c.type == 'M' || c.type == '=' || c.type == 'X' || c.type == 'I' || c.type == 'S') q_len += c.len;
          if (c.type == 'M' || c.type == '=' || c.type == 'X' || c.type == 'D' || c.type == 'N') r_len += c.len;
        }
        if (!ok || q_len != (int64_t)l_seq || (int64_t)pos + r_len != (int64_t)end_pos) {
          rc = LTR_ERR_INVALID;
          break;
        }
      }
      a.pos = pos;
      a.end_pos = end_pos;
      a.flag = flag;
      a.mapq = R->mapq[i];
      const uint8_t* raw = R->raw + R->raw_off[i];
      const size_t raw_n = R->raw_off[i + 1] - R->raw_off[i];
      a.has_hp = aux_has(raw, raw_n, 'H', 'P');
      a.hp = R->hp[i];
      a.has_xa = aux_has(raw, raw_n, 'X', 'A');
      a.deleted = false;
      a.file = (uint32_t)f;
      // hap_gen flag (:283-317): usable for haplotype generation when the read covers region +- MIN_FLANK
      const bool hap_ok = !(params->min_flank > 0 && (pos > start - params->min_flank || end_pos < stop + params->min_flank));
      a.flag = (uint16_t)((a.flag & 0x7fff) | (hap_ok ? 0x8000 : 0));  // carried in the top bit (not a SAM flag here)
      auto mate = potential_mates.find(key);
      if (mate != potential_mates.end()) potential_mates.erase(mate);  // :324-330 (same mate number: both are "not first")
      potential_strs.emplace(key, std::move(a));  // a second alignment of the same name is ignored, as std::map::insert does
    }
    ltr_bam_reads_free(R);
    if (rc != LTR_OK) break;  // (a refusal found in this file must not be overwritten by the next file's fetch)
  }
  if (rc != LTR_OK) {
    delete O;
    return rc;
  }
  std::vector<Aln*> unpaired;  // :420-436
  for (auto it = potential_strs.begin(); it != potential_strs.end(); ++it) {
    if (it->second.has_xa) {
      ++S.n_not_unique;
      continue;
    }
    unpaired.push_back(&it->second);
  }
  S.n_passed = (uint32_t)unpaired.size();
  // :453-483: reads are taken from the BACK of the list; samples (one per file) are numbered by first appearance
  std::vector<std::vector<Aln*>> by_sample;
  std::map<uint32_t, size_t> sample_of_file;
  for (size_t k = unpaired.size(); k-- > 0;) {
    Aln* a = unpaired[k];
    auto it = sample_of_file.find(a->file);
    size_t s;
    if (it == sample_of_file.end()) {
      s = by_sample.size();
      sample_of_file[a->file] = s;
      by_sample.push_back(std::vector<Aln*>());
      O->sample_file.push_back(a->file);
    } else {
      s = it->second;
    }
    by_sample[s].push_back(a);
  }
  // ---- process_phased_reads (snp_bam_processor.cpp:141-232): the counters run on across samples, as in the reference ---
  std::vector<std::vector<double>> p1(by_sample.size()), p2(by_sample.size());
  {
    int32_t total_reads = 0, h1 = 0, h2 = 0;
    bool not_enough = false;
    for (size_t s = 0; s < by_sample.size(); ++s) {
      for (const Aln* a : by_sample[s]) {
        ++total_reads;
        const int hap = a->has_hp ? a->hp : -1;  // get_haplotype, :126-134
        if (hap == 1) ++h1;
        else if (hap == 2) ++h2;
      }
      const double unphased = (double)(total_reads - (h1 + h2)) / (double)total_reads;
      if (unphased > 0.2 || h2 <= 1 || h1 <= 1) not_enough = true;  // :190-193
      for (const Aln* a : by_sample[s]) {
        const int hap = a->has_hp ? a->hp : -1;
        if (params->phased_bam && hap != -1 && !not_enough) {
          p1[s].push_back(hap == 1 ? -0.000001 : -1000.0);  // FROM_HAP_LL / OTHER_HAP_LL, snp_bam_processor.h:16-18
          p2[s].push_back(hap == 2 ? -0.000001 : -1000.0);
        } else {
          p1[s].push_back(0.0);
          p2[s].push_back(0.0);
        }
      }
    }
  }
  // ---- left_align_reads (genotyper_bam_processor.cpp:38-168) ------------------------------------------------------------
  O->read_off.push_back(0);
  O->cigar_off.push_back(0);
  O->name_off.push_back(0);
  O->sample_read_begin.push_back(0);
  for (size_t s = 0; s < by_sample.size() && rc == LTR_OK; ++s) {
    for (size_t j = 0; j < by_sample[s].size() && rc == LTR_OK; ++j) {
      Aln& a = *by_sample[s][j];  // trimmed in place: every read of potential_strs is visited once
      if (a.pos > start || a.end_pos < stop) {  // :55-58
        ++S.n_trim_failed;
        continue;
      }
      trim_alignment(a, start > params->flank_size ? start - params->flank_size : 1, stop + params->flank_size,
                     params->flank_size);  // :60
      std::vector<Op> ops;
      bool soft = false, bad = false;
      int32_t r_start, r_stop;
      if (a.bases.empty()) {  // :61-71 the repeat (and everything around it) is deleted in this read
        r_start = start;
        r_stop = stop;
        a.deleted = true;
      } else {
        r_start = a.pos;
        r_stop = a.end_pos - 1;
        int32_t seq_index = 0;
        int64_t ref_index = a.pos;
        for (const Op& c : a.cigar) {  // :79-129
          switch (c.type) {
            case 'H': break;
            case 'S': ops.push_back(c); seq_index += c.len; soft = true; break;
            case 'I': ops.push_back(c); seq_index += c.len; break;
            case 'D': ops.push_back(c); ref_index += c.len; break;
            case 'M': case '=': case 'X': {
              char prev = '=';
              int32_t num = 0;
              for (int32_t k = 0; k < c.len; ++k, ++ref_index, ++seq_index) {
                const int64_t ri = ref_index - ref_seq_start;
                if (ri < 0 || ri >= ref_seq_len || (size_t)seq_index >= a.bases.size()) { bad = true; break; }
                int rb = ref_seq[ri];
                if (rb >= 'a' && rb <= 'z') rb -= 32;
                int qb = (unsigned char)a.bases[(size_t)seq_index];
                if (qb >= 'a' && qb <= 'z') qb -= 32;
                const char t = (qb == rb) ? '=' : 'X';
                if (t == prev) ++num;
                else {
                  if (num) ops.push_back(Op{prev, num});
                  prev = t;
                  num = 1;
                }
              }
              if (num) ops.push_back(Op{prev, num});
              break;
            }
            default: bad = true; break;  // printErrorAndDie("Invalid CIGAR option ...") in the reference
          }
          if (bad) break;
        }
      }
      if (bad) { rc = LTR_ERR_INVALID; break; }
      if (soft) {  // :131-134
        ++S.n_trim_failed;
        continue;
      }
      O->read_start.push_back(r_start);
      O->read_stop.push_back(r_stop);
      {
        const size_t at = O->read_bytes.size();
        O->read_bytes.resize(at + a.bases.size());
        uint8_t* dst = O->read_bytes.data() + at;
        for (size_t k = 0; k < a.bases.size(); ++k) {
          const char ch = a.bases[k];
          dst[k] = (uint8_t)((ch >= 'a' && ch <= 'z') ? ch - 32 : ch);
        }
      }
      O->qual_bytes.insert(O->qual_bytes.end(), a.quals.begin(), a.quals.end());
      O->read_off.push_back((uint32_t)O->read_bytes.size());
      for (const Op& c : ops) O->cigar_ops.push_back(((uint32_t)c.len << 4) | bam_op(c.type));
      O->cigar_off.push_back((uint32_t)O->cigar_ops.size());
      O->read_sample.push_back((int32_t)s);
      O->log_p1.push_back(p1[s][j]);
      O->log_p2.push_back(p2[s][j]);
      O->hap_gen_ok.push_back(a.deleted && a.bases.empty() ? 1 : ((a.flag & 0x8000) ? 1 : 0));
      O->deleted.push_back(a.deleted ? 1 : 0);
      O->read_hp.push_back(a.has_hp ? a.hp : 0);  // left_align_reads counts these per sample (:146-151): PDP of the record
      O->names.insert(O->names.end(), a.name.begin(), a.name.end());
      O->names.push_back('\0');
      O->name_off.push_back((uint32_t)O->names.size());
    }
    O->sample_read_begin.push_back((uint32_t)O->read_start.size());
  }
  if (rc != LTR_OK) {
    delete O;
    return rc;
  }
  S.n_samples = (uint32_t)by_sample.size();
  S.sample_file = O->sample_file.data();
  S.sample_read_begin = O->sample_read_begin.data();
  S.n_reads = (uint32_t)O->read_start.size();
  S.read_start = O->read_start.data();
  S.read_stop = O->read_stop.data();
  S.read_off = O->read_off.data();
  S.read_bytes = O->read_bytes.data();
  S.qual_bytes = O->qual_bytes.data();
  S.cigar_off = O->cigar_off.data();
  S.cigar_ops = O->cigar_ops.data();
  S.read_sample = O->read_sample.data();
  S.log_p1 = O->log_p1.data();
  S.log_p2 = O->log_p2.data();
  S.hap_gen_ok = O->hap_gen_ok.data();
  S.deleted = O->deleted.data();
  S.read_hp = O->read_hp.data();
  S.name_off = O->name_off.data();
  S.names = O->names.data();
  S.owner = O;
  *out = &O->pub;
  return LTR_OK;
}

// C ABI boundary: no exception leaves the library (malformed input and exhausted memory become error codes)
extern "C" int ltr_region_collect(const ltr_bam* const* bams, int32_t n_bams, const char* chrom, int32_t start, int32_t stop,
                                  const uint8_t* ref_seq, int64_t ref_seq_start, int64_t ref_seq_len,
                                  const ltr_region_params* params, ltr_region_reads** out) {
  try {
    return ltr_region_collect_impl(bams, n_bams, chrom, start, stop, ref_seq, ref_seq_start, ref_seq_len, params, out);
  } catch (const std::bad_alloc&) {
    return LTR_ERR_OOM;
  } catch (...) {
    return LTR_ERR_INVALID;
  }
}

extern "C" void ltr_region_reads_free(ltr_region_reads* r) {
  if (r) delete static_cast<Owner*>(r->owner);
}
