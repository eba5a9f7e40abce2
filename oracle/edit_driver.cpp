// TEST INFRASTRUCTURE ONLY -- C entry points around the reference's own HaplotypeGenerator::needleman_wunsch and
// greedy_clustering (src/SeqAlignment/HaplotypeGenerator.cpp:201-271), compiled IN PLACE from /root/reference by
// oracle/build_ref.sh and linked into oracle/_ref/libltr_ref.so.  The two members are private; this translation unit is
// compiled with -fno-access-control (the reference's sources are not touched).
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "SeqAlignment/HaplotypeGenerator.h"

extern "C" int32_t ltr_ref_edit_score(const char* cent, int32_t n, const char* read, int32_t m, int32_t T) {
  HaplotypeGenerator gen(0, 0, 5);
  int score = -1;
  gen.needleman_wunsch(std::string(cent, (size_t)n), std::string(read, (size_t)m), score, T);
  return score;
}

// The sequences of one set must be distinct (the reference keys its clusters by the centroid STRING, and its callers
// pass the keys of a map).  centroid_of[i] = index of the centroid whose cluster holds seqs[i].
extern "C" int32_t ltr_ref_greedy_cluster(const uint8_t* seq_bytes, const uint32_t* seq_off, const uint32_t* items,
                                          int32_t n_items, int32_t threshold, int32_t* centroid_of,
                                          int32_t* n_centroids_out) {
  HaplotypeGenerator gen(0, 0, 5);
  std::vector<std::string> seqs;
  std::map<std::string, int32_t> index_of;
  for (int32_t i = 0; i < n_items; ++i) {
    const uint32_t s = items[i];
    seqs.push_back(std::string((const char*)seq_bytes + seq_off[s], (size_t)(seq_off[s + 1] - seq_off[s])));
    if (index_of.count(seqs.back())) return -1;
    index_of[seqs.back()] = i;
  }
  std::map<std::string, std::vector<std::string> > clusters;
  const bool ok = gen.greedy_clustering(seqs, clusters, threshold);
  for (int32_t i = 0; i < n_items; ++i) centroid_of[i] = -1;
  for (std::map<std::string, std::vector<std::string> >::const_iterator it = clusters.begin(); it != clusters.end(); ++it)
    for (size_t k = 0; k < it->second.size(); ++k) centroid_of[index_of[it->second[k]]] = index_of[it->first];
  *n_centroids_out = (int32_t)clusters.size();
  return ok ? 1 : 0;
}
