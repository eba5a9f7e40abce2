// poa.h -- partial-order consensus and thresholded edit distance for the candidate-allele assembly (host side of SURVEY.md
// section 8f, N2; reference src/SeqAlignment/HaplotypeGenerator.cpp:167-292).
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

namespace ltr {

// Consensus of the sequences, added in the given order, under the configuration of the reference's only call site
// (HaplotypeGenerator.cpp:167-178): global alignment, match +1, mismatch -1, linear gap -1, weight 1 per base, heaviest
// bundle.  Restates spoa's published algorithm (spoa itself is un-vendored and unpinned: parity UNPINNED, see poa.cpp).
class PoaGraph {
 public:
  void clear();
  void add(const uint8_t* seq, uint32_t len);
  void consensus(std::string& out);
  uint32_t n_nodes() const { return (uint32_t)code_.size(); }
  void force_wide_cells(bool on) { force_wide_ = on; }  // tests: 32-bit matrix cells even where 16 bits are enough

 private:
  uint32_t new_node(uint8_t code);
  void link(uint32_t tail, uint32_t head, uint32_t weight);
  void sort_nodes();
  void align(const uint8_t* seq, uint32_t len, std::vector<int32_t>& aln_node, std::vector<int32_t>& aln_pos);
  template <typename T>
  void align_cells(const uint8_t* seq, uint32_t len, std::vector<int32_t>& aln_node, std::vector<int32_t>& aln_pos);
  uint32_t complete_branch(uint32_t rank, std::vector<int64_t>& score, std::vector<int32_t>& pred) const;

  // nodes
  std::vector<uint8_t> code_;                    // base byte of the node
  std::vector<std::vector<uint32_t> > in_, out_; // edge ids in order of creation
  std::vector<std::vector<uint32_t> > peers_;    // nodes aligned to this one (other bases of the same column)
  // edges
  std::vector<uint32_t> tail_, head_;
  std::vector<int64_t> weight_;
  // order
  std::vector<uint32_t> order_, rank_;           // topological order that keeps aligned nodes adjacent; its inverse
  // scratch
  std::vector<int32_t> H_;
  std::vector<int32_t> aln_node_, aln_pos_;
  bool force_wide_ = false;
};

// HaplotypeGenerator::needleman_wunsch (:201-235) as far as its callers (greedy_clustering :238-271, merge_clusters
// :274-292) look at it: they only test `score < T` and compare scores below T.  Returns the unit-cost edit distance when it
// is below T and some value >= T otherwise (bit-vector recurrence on 64-bit words).  Empty strings as in the reference:
// an empty cent_seq gives |read_seq|, an empty read_seq with a non-empty cent_seq gives T + 1.
int thresholded_edit_distance(const std::string& cent_seq, const std::string& read_seq, int T);
// The two halves of it: the unit-cost edit distance itself (independent of T, so a caller that walks the ladder of
// thresholds can remember it), and the reference's answer for (|cent_seq|, |read_seq|, distance, T).
int edit_distance(const std::string& a, const std::string& b);
// the distance when it is <= k, k + 1 otherwise: only the blocks within k diagonals of the main one are evaluated
int bounded_edit_distance(const std::string& a, const std::string& b, int k);
int thresholded_from_distance(int n, int m, int distance, int T);

}  // namespace ltr
