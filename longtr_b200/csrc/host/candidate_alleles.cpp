// candidate_alleles.cpp -- ltr_candidate_alleles: the candidate haplotype block of one region from its reads, the way
// SeqStutterGenotyper::build_haplotype obtains it (reference src/seq_stutter_genotyper.cpp:416-476) through
//   HaplotypeGenerator::add_haplotype_block     src/SeqAlignment/HaplotypeGenerator.cpp:521-570
//   HaplotypeGenerator::gen_candidate_seqs      :296-481  (allele support counts, thresholds, ordering)
//   HaplotypeGenerator::extract_sequence        :98-164   (the region's bases of one alignment)
//   HaplotypeGenerator::trim                    :14-96    (identical allele ends are clipped)
//   HaplotypeGenerator::fuse_haplotype_blocks   :572-607  (reference flanks of up to 35 bp)
//   HaplotypeGenerator::greedy_clustering / merge_clusters / poa   :167-292  (the assembly branch of gen_candidate_seqs)
// SURVEY.md section 8f, N2.  When a sample leaves more than a quarter of its reads without a candidate the reference
// clusters those reads (greedy_clustering, thresholds 20 ... 700), replaces every cluster by a partial-order consensus (spoa),
// merges clusters whose consensus sequences are close, and adds the consensus of every well-supported cluster as an "inexact"
// allele (:397-471).  That branch runs here on the host thread that prepares the region (north star: candidate-haplotype
// construction stays on host threads), with poa.cpp standing in for the un-vendored spoa (parity of the consensus itself
// unpinned, see there).  With LTR_CAND_FLAG_NO_ASSEMBLY the region is answered with LTR_CAND_NEEDS_ASSEMBLY instead and the
// sets that would be clustered are listed (for ltr_cluster_greedy on the device).
#include <limits.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "longtr_b200.h"
#include "poa.h"

namespace {

struct ReadView {
  int32_t start, stop;
  bool deleted;
  const uint32_t* cigar;
  uint32_t n_cigar;
  const uint8_t* bases;
  uint32_t n_bases;
};

std::string upper(std::string s) {
  for (char& c : s)
    if (c >= 'a' && c <= 'z') c = (char)(c - 32);
  return s;
}

// :98-164.  0 = read does not cover the region, 1 = seq holds the region's bases, -1 = malformed CIGAR
int extract_sequence(const ReadView& a, int32_t region_start, int32_t region_end, std::string& seq) {
  if (a.deleted) {
    seq.clear();
    return 1;
  }
  if (a.start >= region_start) return 0;
  if (a.stop <= region_end) return 0;
  uint32_t ci = 0;
  int32_t char_index = 0;
  int32_t pos = a.start;
  uint32_t si = 0;  // index into the read's bases (the reference indexes the gapped alignment string)
  std::string reg;
  auto take = [&](int32_t n) {
    if (si + (uint32_t)n > a.n_bases) return false;
    reg.append((const char*)a.bases + si, (size_t)n);
    return true;
  };
  while (ci < a.n_cigar) {
    const int32_t num = (int32_t)(a.cigar[ci] >> 4);
    const uint32_t op = a.cigar[ci] & 15;
    const bool ins = op == 1, del = op == 2, mat = (op == 0 || op == 7 || op == 8);
    if (!ins && !del && !mat) return -1;
    if (char_index == num) {
      ++ci;
      char_index = 0;
    } else if (pos > region_end) {
      seq = upper(reg);
      return 1;
    } else if (pos == region_end) {
      if (ins) {
        if (!take(num)) return -1;
        si += (uint32_t)num;
        char_index = 0;
        ++ci;
      } else {
        seq = upper(reg);
        return 1;
      }
    } else if (pos >= region_start) {
      int32_t n = std::min(region_end - pos, num - char_index);
      if (ins) {
        n = num;
        if (!take(n)) return -1;
        si += (uint32_t)n;
      } else if (mat) {
        if (!take(n)) return -1;
        si += (uint32_t)n;
        pos += n;
      } else {
        pos += n;
      }
      char_index += n;
    } else {
      int32_t n;
      if (ins) {
        n = num - char_index;
        si += (uint32_t)n;
      } else {
        n = std::min(region_start - pos, num - char_index);
        pos += n;
        if (mat) si += (uint32_t)n;
      }
      char_index += n;
    }
  }
  return -1;  // "Logical error in extract_sequence"
}

typedef std::pair<std::string, bool> Seq;

bool by_length_and_sequence(const Seq& a, const Seq& b) {  // stringops.cpp:41-45
  if (a.first.size() != b.first.size()) return a.first.size() < b.first.size();
  return a.first.compare(b.first) < 0;
}

// :14-96
void trim(int ideal_min_length, int left_pad, int right_pad, int32_t& region_start, int32_t& region_end, std::vector<Seq>& seqs) {
  int min_len = INT_MAX;
  for (const Seq& s : seqs) min_len = std::min(min_len, (int)s.first.size());
  if (min_len <= ideal_min_length) return;
  int max_left = 0, max_right = 0;
  while (max_left < min_len - ideal_min_length) {
    size_t j = 1;
    while (j < seqs.size() && seqs[j].first[(size_t)max_left] == seqs[j - 1].first[(size_t)max_left]) ++j;
    if (j != seqs.size()) break;
    ++max_left;
  }
  while (max_right < min_len - ideal_min_length) {
    const char c = seqs[0].first[seqs[0].first.size() - 1 - (size_t)max_right];
    size_t j = 1;
    while (j < seqs.size() && seqs[j].first[seqs[j].first.size() - 1 - (size_t)max_right] == c) ++j;
    if (j != seqs.size()) break;
    ++max_right;
  }
  max_left = std::min(left_pad, max_left);
  max_right = std::min(right_pad, max_right);
  max_left = std::max(0, std::min(min_len - right_pad, max_left));
  max_right = std::max(0, std::min(min_len - left_pad, max_right));
  int lt, rt;
  if (min_len - 2 * std::min(max_left, max_right) <= ideal_min_length) {
    lt = rt = std::min(max_left, max_right);
    while (min_len - lt - rt < ideal_min_length) {
      if (lt > rt) --lt;
      else --rt;
    }
  } else if (max_left > max_right) {
    rt = max_right;
    lt = std::min(max_left, min_len - ideal_min_length - max_right);
  } else {
    lt = max_left;
    rt = std::min(max_right, min_len - ideal_min_length - max_left);
  }
  for (Seq& s : seqs) s.first = s.first.substr((size_t)lt, s.first.size() - (size_t)lt - (size_t)rt);
  region_start += lt;
  region_end -= rt;
}

typedef std::map<std::string, std::vector<std::string> > Clusters;  // centroid -> members; iteration in key order matters

// The ladder of thresholds and the consensus / merge rounds compare the same pairs of sequences again and again, and only the
// threshold changes.  Per region a pair's distance is remembered -- exactly when it was found below the cut-off it was computed
// with (ltr::bounded_edit_distance evaluates only the diagonals within the cut-off), as a lower bound otherwise; a larger
// threshold later recomputes with a doubled cut-off.  The callers only test `score < T` and compare scores below T.
struct DistanceMemo {
  struct Entry {
    int value;    // the distance if exact, else a value the distance is known to exceed or equal
    bool exact;
  };
  std::map<std::pair<std::string, std::string>, Entry> known;
  int operator()(const std::string& cent_seq, const std::string& read_seq, int T) {
    const int n = (int)cent_seq.size(), m = (int)read_seq.size();
    if (std::abs(n - m) > T || n == 0 || m == 0) return ltr::thresholded_from_distance(n, m, 0, T);
    Entry& e = known.insert(std::make_pair(std::make_pair(cent_seq, read_seq), Entry{0, false})).first->second;
    if (!e.exact && e.value < T) {
      int k = std::max(T, 2 * e.value);
      k = std::min(std::max(n, m), (k + 63) / 64 * 64);
      const int d = ltr::bounded_edit_distance(cent_seq, read_seq, k);
      e.exact = d <= k;
      e.value = d;  // k + 1 when not exact: the distance is at least that
    }
    if (e.exact) return ltr::thresholded_from_distance(n, m, e.value, T);
    return T + 1;  // distance >= e.value >= T
  }
};

// :238-271.  false: more than 15 centroids at this threshold.
bool greedy_clustering(const std::vector<std::string>& seqs, Clusters& clusters, int T, DistanceMemo& dist) {
  std::vector<const std::string*> centroids(1, &seqs[0]);
  clusters[seqs[0]].push_back(seqs[0]);
  for (size_t i = 1; i < seqs.size(); ++i) {
    int min_score = INT_MAX, min_cntr = -1;
    for (size_t j = 0; j < centroids.size(); ++j) {
      const int score = dist(seqs[i], *centroids[j], T);
      if (score < T && score < min_score) {
        min_cntr = (int)j;
        min_score = score;
      }
    }
    if (min_cntr != -1) {
      clusters[*centroids[(size_t)min_cntr]].push_back(seqs[i]);
    } else {
      centroids.push_back(&seqs[i]);
      if (centroids.size() > 15) return false;
      clusters[seqs[i]].push_back(seqs[i]);
    }
  }
  return true;
}

// :274-292 (the inner index starts at 1 there as well)
bool merge_clusters(const std::vector<std::string>& cent, Clusters& clusters, int T, DistanceMemo& dist) {
  bool updated = false;
  for (size_t i = 0; i < cent.size(); ++i)
    for (size_t j = 1; j < cent.size(); ++j) {
      if (i == j || clusters.find(cent[i]) == clusters.end() || clusters.find(cent[j]) == clusters.end()) continue;
      if (dist(cent[i], cent[j], T) < T) {
        updated = true;
        const std::vector<std::string> moved = clusters[cent[j]];
        std::vector<std::string>& into = clusters[cent[i]];
        into.insert(into.end(), moved.begin(), moved.end());
        clusters.erase(cent[j]);
      }
    }
  return updated;
}

// :167-199.  Fewer than 30 sequences: all of them in order.  Otherwise the reference draws 30 distinct indices from
// std::random_device (not reproducible by construction); here the same rejection loop runs on a generator with a fixed seed,
// so that a region always gets the same alleles.
struct Lcg {  // minimal standard generator: the draw must not depend on the C++ library's distribution code
  uint64_t x;
  uint32_t next(uint32_t n) {
    x = x * 6364136223846793005ull + 1442695040888963407ull;
    return (uint32_t)((x >> 33) % n);
  }
};
// The consensus is a function of the list of sequences alone, and the ladder of thresholds and the consensus / merge rounds
// keep asking for the same lists: answers are remembered per region (key: the sequences with their lengths).
typedef std::map<std::string, std::string> PoaMemo;
void poa(ltr::PoaGraph& graph, PoaMemo& memo, const std::vector<std::string>& seqs, std::string& consensus, uint32_t& n_poa) {
  const size_t kLimit = 30;
  std::string key;
  for (const std::string& s : seqs) {
    const uint32_t n = (uint32_t)s.size();
    key.append((const char*)&n, sizeof(n));
    key.append(s);
  }
  auto hit = memo.find(key);
  if (hit != memo.end()) {
    consensus = hit->second;
    return;
  }
  graph.clear();
  ++n_poa;
  if (seqs.size() < kLimit) {
    for (const std::string& s : seqs) graph.add((const uint8_t*)s.data(), (uint32_t)s.size());
  } else {
    std::vector<uint32_t> idx;
    Lcg gen = {0x4c6f6e675452ull + seqs.size()};
    while (idx.size() < kLimit) {
      const uint32_t r = gen.next((uint32_t)seqs.size());
      if (std::find(idx.begin(), idx.end(), r) == idx.end()) idx.push_back(r);
    }
    for (uint32_t k : idx) graph.add((const uint8_t*)seqs[k].data(), (uint32_t)seqs[k].size());
  }
  graph.consensus(consensus);
  memo[key] = consensus;
}

// :397-471 for one sample: ladder of thresholds, clustering, consensus / merge until nothing merges, support tests.
void assemble_sample(const std::map<std::string, int>& not_added, int n_ignored, std::vector<Seq>& seqs, uint32_t& n_poa,
                     int32_t& threshold_used) {
  std::vector<std::string> uniq;
  for (auto it = not_added.begin(); it != not_added.end(); ++it) uniq.push_back(it->first);
  std::sort(uniq.begin() + 1, uniq.end(), [](const std::string& a, const std::string& b) {
    return a.size() != b.size() ? a.size() < b.size() : a.compare(b) < 0;
  });
  static const int kThresholds[] = {20, 50, 80, 100, 150, 200, 300, 400, 500, 600, 700};
  ltr::PoaGraph graph;
  PoaMemo memo;
  DistanceMemo dist;
  for (int t : kThresholds) {
    Clusters clusters;
    if (!greedy_clustering(uniq, clusters, t, dist)) continue;
    bool not_converged = true;
    while (not_converged) {
      Clusters updated;
      std::vector<std::string> cent;
      for (auto it = clusters.begin(); it != clusters.end(); ++it) {
        std::string consensus;
        poa(graph, memo, it->second, consensus, n_poa);
        if (std::find(cent.begin(), cent.end(), consensus) == cent.end()) {
          cent.push_back(consensus);
          updated[consensus] = it->second;
        } else {
          std::vector<std::string>& into = updated[consensus];
          into.insert(into.end(), it->second.begin(), it->second.end());
        }
      }
      std::sort(cent.begin() + 1, cent.end(), [](const std::string& a, const std::string& b) {
        return a.size() != b.size() ? a.size() < b.size() : a.compare(b) < 0;
      });
      not_converged = merge_clusters(cent, updated, t, dist);
      clusters.swap(updated);
    }
    int covered = 0;
    std::vector<Seq> potential;
    for (auto it = clusters.begin(); it != clusters.end(); ++it) {
      int sum = 0;
      for (const std::string& s : it->second) {
        auto f = not_added.find(s);
        if (f != not_added.end()) sum += f->second;  // (the reference's operator[] would insert a zero)
      }
      if (sum > std::min((int)(n_ignored * 0.10), 10)) {
        covered += sum;
        if (std::find(seqs.begin(), seqs.end(), Seq(it->first, false)) == seqs.end() &&
            std::find(seqs.begin(), seqs.end(), Seq(it->first, true)) == seqs.end())
          potential.push_back(Seq(it->first, true));
      }
    }
    if (covered >= (int)(0.80 * n_ignored)) {
      for (const Seq& s : potential) seqs.push_back(s);
      threshold_used = t;
      return;
    }
  }
}

struct Owner {
  ltr_candidates pub;
  std::vector<uint8_t> allele_inexact;
  std::vector<uint32_t> allele_off;
  std::vector<uint8_t> allele_bytes;
  std::string lflank, rflank;
  std::vector<uint32_t> cluster_sample_begin, cluster_off;
  std::vector<uint8_t> cluster_bytes;
  std::vector<int32_t> cluster_count;
};

}  // namespace

static int ltr_candidate_alleles_flags_impl(const ltr_region_reads* reads, int32_t region_start, int32_t region_stop,
                                           int32_t period, const uint8_t* ref_seq, int64_t ref_seq_start, int64_t ref_seq_len,
                                           int32_t indel_flank_len, uint32_t flags, ltr_candidates** out) {
  if (!reads || !ref_seq || !out || region_stop < region_start || period < 1 || indel_flank_len < 0) return LTR_ERR_INVALID;
  *out = nullptr;
  Owner* O = new Owner();
  memset(&O->pub, 0, sizeof(O->pub));
  ltr_candidates& C = O->pub;
  C.owner = O;
  *out = &O->pub;
  const double MIN_FRAC_READS = 0.05, MIN_FRAC_SAMPLES = 0.05, MIN_FRAC_STRONG_SAMPLE = 0.2, MIN_READS_STRONG_SAMPLE = 2,
               MIN_STRONG_SAMPLES = 1;  // HaplotypeGenerator.h:60-66
  const int LEFT_PAD = indel_flank_len, RIGHT_PAD = indel_flank_len;
  const int32_t REF_FLANK_LEN = 35;
  const int64_t chrom_end = ref_seq_start + ref_seq_len;  // the slice stands in for the chromosome
  auto ref_sub = [&](int64_t a, int64_t b) {  // uppercase(chrom_seq.substr(a, b - a))
    std::string s;
    for (int64_t p = a; p < b; ++p) {
      char c = (p >= ref_seq_start && p < chrom_end) ? (char)ref_seq[p - ref_seq_start] : 'N';
      if (c >= 'a' && c <= 'z') c = (char)(c - 32);
      s.push_back(c);
    }
    return s;
  };
  // :527-530
  if (region_start < REF_FLANK_LEN + LEFT_PAD || (int64_t)region_stop + REF_FLANK_LEN + RIGHT_PAD > chrom_end ||
      region_start - LEFT_PAD - REF_FLANK_LEN < ref_seq_start) {
    C.status = LTR_CAND_NEAR_CHROM_END;
    return LTR_OK;
  }
  // reads: all of them bound the flanks (build_haplotype :423-427), those marked hap_gen_ok generate alleles (:433-436)
  int32_t all_min = INT_MAX, all_max = INT_MIN, min_aln_start = INT_MAX, max_aln_stop = INT_MIN;
  std::vector<std::vector<ReadView>> by_sample(reads->n_samples);
  for (uint32_t i = 0; i < reads->n_reads; ++i) {
    all_min = std::min(all_min, reads->read_start[i]);
    all_max = std::max(all_max, reads->read_stop[i]);
    if (!reads->hap_gen_ok[i]) continue;
    const int32_t s = reads->read_sample[i];
    if (s < 0 || (uint32_t)s >= reads->n_samples) return LTR_ERR_INVALID;
    ReadView v;
    v.start = reads->read_start[i];
    v.stop = reads->read_stop[i];
    v.deleted = reads->deleted[i] != 0;
    v.cigar = reads->cigar_ops + reads->cigar_off[i];
    v.n_cigar = reads->cigar_off[i + 1] - reads->cigar_off[i];
    v.bases = reads->read_bytes + reads->read_off[i];
    v.n_bases = reads->read_off[i + 1] - reads->read_off[i];
    by_sample[(size_t)s].push_back(v);
    min_aln_start = std::min(min_aln_start, v.start);  // get_aln_bounds :483-494
    max_aln_stop = std::max(max_aln_stop, v.stop);
  }
  int32_t rs = region_start - LEFT_PAD, re = region_stop + RIGHT_PAD;
  const std::string ref_allele = ref_sub(rs, re);
  if ((int64_t)min_aln_start + 5 >= rs || (int64_t)max_aln_stop - 5 <= re) {  // :540-543
    C.status = LTR_CAND_NO_SPANNING;
    return LTR_OK;
  }
  // ---- gen_candidate_seqs (:296-481) -----------------------------------------------------------------------------------
  std::map<std::string, double> sample_counts;
  std::map<std::string, int> read_counts, must_inc;
  int tot_reads = 0, tot_samples = 0;
  std::string sub;
  for (size_t s = 0; s < by_sample.size(); ++s) {
    int samp_reads = 0;
    std::map<std::string, int> counts;
    for (const ReadView& v : by_sample[s]) {
      const int e = extract_sequence(v, rs, re, sub);
      if (e < 0) return LTR_ERR_INVALID;
      if (e) {
        read_counts[sub] += 1;
        counts[sub] += 1;
        ++tot_reads;
        ++samp_reads;
      }
    }
    for (auto it = counts.begin(); it != counts.end(); ++it) {
      if (it->second >= MIN_READS_STRONG_SAMPLE && it->second >= MIN_FRAC_STRONG_SAMPLE * samp_reads) must_inc[it->first] += 1;
      sample_counts[it->first] += it->second * 1.0 / samp_reads;
    }
    if (samp_reads > 0) ++tot_samples;
  }
  std::vector<Seq> seqs;
  int ref_index = -1;
  for (auto it = must_inc.begin(); it != must_inc.end(); ++it) {  // :346-356
    if (it->second >= MIN_STRONG_SAMPLES) {
      sample_counts.erase(it->first);
      read_counts.erase(it->first);
      seqs.push_back(Seq(it->first, false));
      if (it->first.compare(ref_allele) == 0) ref_index = (int)seqs.size() - 1;
    }
  }
  for (auto it = sample_counts.begin(); it != sample_counts.end(); ++it) {  // :359-365
    if (it->second > MIN_FRAC_SAMPLES * tot_samples * 2 || read_counts[it->first] > MIN_FRAC_READS * tot_reads * 2) {
      seqs.push_back(Seq(it->first, false));
      if (ref_index == -1 && it->first.compare(ref_allele) == 0) ref_index = (int)seqs.size() - 1;
    }
  }
  if (ref_index == -1) seqs.insert(seqs.begin(), Seq(ref_allele, false));  // :368-373
  else {
    seqs[(size_t)ref_index] = seqs[0];
    seqs[0] = Seq(ref_allele, false);
  }
  // :377-395 -- reads without a candidate, per sample; a sample with more than a quarter of them triggers the assembly
  O->cluster_sample_begin.push_back(0);
  O->cluster_off.push_back(0);
  std::vector<std::pair<std::map<std::string, int>, int> > pending;
  for (size_t s = 0; s < by_sample.size(); ++s) {
    std::map<std::string, int> not_added;
    int samp_reads = 0, samp_ignored = 0;
    for (const ReadView& v : by_sample[s]) {
      if (extract_sequence(v, rs, re, sub) == 1) {
        ++samp_reads;
        if (std::find(seqs.begin(), seqs.end(), Seq(sub, false)) == seqs.end()) {
          not_added[sub] += 1;
          ++samp_ignored;
        }
      }
    }
    if (samp_ignored > samp_reads * 0.25) {
      // the sequences greedy_clustering sees, in the reference's order (:398-403), with their read counts
      std::vector<std::string> uniq;
      for (auto it = not_added.begin(); it != not_added.end(); ++it) uniq.push_back(it->first);
      if (uniq.size() > 1)
        std::sort(uniq.begin() + 1, uniq.end(), [](const std::string& a, const std::string& b) {
          return a.size() != b.size() ? a.size() < b.size() : a.compare(b) < 0;
        });
      for (const std::string& u : uniq) {
        O->cluster_bytes.insert(O->cluster_bytes.end(), u.begin(), u.end());
        O->cluster_off.push_back((uint32_t)O->cluster_bytes.size());
        O->cluster_count.push_back(not_added[u]);
      }
      O->cluster_sample_begin.push_back((uint32_t)O->cluster_count.size());
      pending.push_back(std::make_pair(not_added, samp_ignored));
    }
  }
  if (!pending.empty()) {
    if (flags & LTR_CAND_FLAG_NO_ASSEMBLY) {
      C.status = LTR_CAND_NEEDS_ASSEMBLY;
    } else {
      for (const auto& p : pending) {  // :397-471, sample after sample on the growing allele list
        int32_t t_used = 0;
        assemble_sample(p.first, p.second, seqs, C.n_consensus, t_used);
        C.assembly_threshold = std::max(C.assembly_threshold, t_used);
      }
    }
  }
  std::sort(seqs.begin() + 1, seqs.end(), by_length_and_sequence);  // :475
  trim(3 * period, LEFT_PAD, RIGHT_PAD, rs, re, seqs);               // :480, ideal_min_length :556
  // ---- fuse_haplotype_blocks (:572-607) ----------------------------------------------------------------------------------
  // min_aln_start_ / max_aln_stop_ of the generator: over ALL reads of the locus (build_haplotype :423-429)
  const int32_t min_start = std::min(rs - 10, std::max(rs - REF_FLANK_LEN, all_min));
  const int32_t max_stop = std::max(re + 10, std::min(re + REF_FLANK_LEN, all_max));
  O->lflank = ref_sub(min_start, rs);
  O->rflank = ref_sub(re, max_stop);
  O->allele_off.push_back(0);
  for (const Seq& s : seqs) {
    O->allele_bytes.insert(O->allele_bytes.end(), s.first.begin(), s.first.end());
    O->allele_off.push_back((uint32_t)O->allele_bytes.size());
    O->allele_inexact.push_back(s.second ? 1 : 0);
  }
  C.block_start = rs;
  C.block_end = re;
  C.n_alleles = (int32_t)seqs.size();
  C.allele_off = O->allele_off.data();
  C.allele_bytes = O->allele_bytes.data();
  C.allele_inexact = O->allele_inexact.data();
  C.lflank_start = min_start;
  C.lflank = O->lflank.c_str();
  C.rflank = O->rflank.c_str();
  C.n_cluster_samples = (uint32_t)O->cluster_sample_begin.size() - 1;
  C.cluster_sample_begin = O->cluster_sample_begin.data();
  C.cluster_off = O->cluster_off.data();
  C.cluster_bytes = O->cluster_bytes.data();
  C.cluster_count = O->cluster_count.data();
  return LTR_OK;
}

// C ABI boundary: no exception leaves the library (malformed input and exhausted memory become error codes)
extern "C" int ltr_candidate_alleles_flags(const ltr_region_reads* reads, int32_t region_start, int32_t region_stop,
                                           int32_t period, const uint8_t* ref_seq, int64_t ref_seq_start, int64_t ref_seq_len,
                                           int32_t indel_flank_len, uint32_t flags, ltr_candidates** out) {
  try {
    return ltr_candidate_alleles_flags_impl(reads, region_start, region_stop, period, ref_seq, ref_seq_start, ref_seq_len, indel_flank_len, flags, out);
  } catch (const std::bad_alloc&) {
    return LTR_ERR_OOM;
  } catch (...) {
    return LTR_ERR_INVALID;
  }
}

extern "C" int ltr_candidate_alleles(const ltr_region_reads* reads, int32_t region_start, int32_t region_stop, int32_t period,
                                     const uint8_t* ref_seq, int64_t ref_seq_start, int64_t ref_seq_len,
                                     int32_t indel_flank_len, ltr_candidates** out) {
  return ltr_candidate_alleles_flags(reads, region_start, region_stop, period, ref_seq, ref_seq_start, ref_seq_len,
                                     indel_flank_len, 0u, out);
}

// The consensus alone (diagnostics, tests, hosts that cluster elsewhere): sequences in the order they are to be added.
extern "C" int ltr_poa_consensus(const uint8_t* seq_bytes, const uint32_t* seq_off, uint32_t n_seqs, uint8_t* out,
                                 uint32_t out_capacity, uint32_t* out_len) {
  if ((n_seqs && (!seq_bytes || !seq_off)) || !out_len || (out_capacity && !out)) return LTR_ERR_INVALID;
  for (uint32_t i = 0; i < n_seqs; ++i)
    if (seq_off[i + 1] < seq_off[i]) return LTR_ERR_INVALID;
  ltr::PoaGraph graph;
  for (uint32_t i = 0; i < n_seqs; ++i) graph.add(seq_bytes + seq_off[i], seq_off[i + 1] - seq_off[i]);
  std::string c;
  graph.consensus(c);
  *out_len = (uint32_t)c.size();
  if (c.size() > out_capacity) return LTR_ERR_INVALID;
  if (!c.empty()) memcpy(out, c.data(), c.size());
  return LTR_OK;
}

extern "C" void ltr_candidates_free(ltr_candidates* c) {
  if (c) delete static_cast<Owner*>(c->owner);
}
