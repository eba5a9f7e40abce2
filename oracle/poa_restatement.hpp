// TEST INFRASTRUCTURE ONLY -- CPU restatement of the partial-order alignment the reference obtains from spoa.
//
// PARITY UNPINNED.  spoa (github.com/rvaser/spoa) is a third-party dependency the reference clones at build time without
// a version pin (reference Makefile:96-103, HEAD of the default branch) and that is absent from /root/reference and from
// this image; nothing of it can be compiled or run here and the reference has no test or golden vector for it.  What follows
// restates the algorithm spoa 4.x publishes (Lee, Grasso & Sharlow 2002 partial-order alignment; Vaser et al. 2017) for the
// one configuration the reference's only call site uses (src/SeqAlignment/HaplotypeGenerator.cpp:167-199):
//   AlignmentEngine::Create(AlignmentType::kNW, m = 1, n = -1, g = -1)    global alignment, linear gaps
//   for every sequence:  alignment = engine->Align(seq, graph);  graph.AddAlignment(alignment, seq)   (weight 1 per base)
//   consensus = graph.GenerateConsensus()                                  heaviest bundle + branch completion
// in spoa's own structure (nodes / edges / aligned-node rings, topological sort that keeps aligned nodes adjacent, row per
// node in rank order, predecessor order = order of edge creation, back-track preference match > graph gap > sequence gap,
// first maximum among the sink rows).  It is the checker for the product's own implementation (longtr_b200/csrc/host/poa.cpp,
// flat arrays, written independently) and -- through oracle/shim/spoa/spoa.hpp with -DLTR_SPOA_RESTATEMENT -- what the
// reference's HaplotypeGenerator, compiled in place, calls when a region needs the assembly.
#ifndef LTR_ORACLE_POA_RESTATEMENT_HPP
#define LTR_ORACLE_POA_RESTATEMENT_HPP
#include <algorithm>
#include <cstdint>
#include <limits>
#include <memory>
#include <stack>
#include <string>
#include <utility>
#include <vector>

namespace ltr_poa_oracle {

typedef std::vector<std::pair<std::int32_t, std::int32_t> > Alignment;  // (node id | -1, sequence position | -1)

class Graph {
 public:
  struct Edge;
  struct Node {
    std::uint32_t id, code;
    std::vector<Edge*> inedges, outedges;
    std::vector<Node*> aligned_nodes;
  };
  struct Edge {
    Node *tail, *head;
    std::int64_t weight;
  };

  Graph() : num_codes_(0), coder_(256, -1), decoder_(256, -1), n_sequences_(0) {}

  const std::vector<Node*>& rank_to_node() const { return rank_to_node_; }
  std::size_t n_nodes() const { return nodes_.size(); }
  std::int32_t coder(char c) const { return coder_[(unsigned char)c]; }
  std::int32_t decoder(std::uint32_t code) const { return decoder_[code]; }
  std::uint32_t num_codes() const { return num_codes_; }

  void AddAlignment(const Alignment& alignment, const std::string& sequence, std::uint32_t weight = 1) {
    const std::uint32_t len = (std::uint32_t)sequence.size();
    if (len == 0) return;
    for (std::uint32_t i = 0; i < len; ++i) {
      const unsigned char c = (unsigned char)sequence[i];
      if (coder_[c] == -1) {
        coder_[c] = (std::int32_t)num_codes_;
        decoder_[num_codes_++] = c;
      }
    }
    const std::vector<std::uint32_t> weights(len, weight);
    if (alignment.empty()) {
      AddSequence(sequence, weights, 0, len);
      ++n_sequences_;
      TopologicalSort();
      return;
    }
    std::vector<std::uint32_t> valid;
    for (std::size_t k = 0; k < alignment.size(); ++k)
      if (alignment[k].second != -1) valid.push_back((std::uint32_t)alignment[k].second);
    Node* begin = AddSequence(sequence, weights, 0, valid.front());
    Node* prev = begin ? nodes_.back().get() : nullptr;
    Node* last = AddSequence(sequence, weights, valid.back() + 1, len);
    for (std::size_t k = 0; k < alignment.size(); ++k) {
      const std::int32_t nid = alignment[k].first, pos = alignment[k].second;
      if (pos == -1) continue;
      const std::uint32_t code = (std::uint32_t)coder_[(unsigned char)sequence[(std::size_t)pos]];
      Node* curr = nullptr;
      if (nid == -1) {
        curr = AddNode(code);
      } else {
        Node* jt = nodes_[(std::size_t)nid].get();
        if (jt->code == code) {
          curr = jt;
        } else {
          for (Node* kt : jt->aligned_nodes)
            if (kt->code == code) {
              curr = kt;
              break;
            }
          if (!curr) {
            curr = AddNode(code);
            for (Node* kt : jt->aligned_nodes) {
              kt->aligned_nodes.push_back(curr);
              curr->aligned_nodes.push_back(kt);
            }
            jt->aligned_nodes.push_back(curr);
            curr->aligned_nodes.push_back(jt);
          }
        }
      }
      if (!begin) begin = curr;
      if (prev) AddEdge(prev, curr, weights[(std::size_t)pos - 1] + weights[(std::size_t)pos]);
      prev = curr;
    }
    if (last) AddEdge(prev, last, weights[valid.back()] + weights[valid.back() + 1]);
    ++n_sequences_;
    TopologicalSort();
  }

  std::string GenerateConsensus() {
    std::string dst;
    if (rank_to_node_.empty()) return dst;
    std::vector<Node*> pred(nodes_.size(), nullptr);
    std::vector<std::int64_t> score(nodes_.size(), -1);
    Node* max = nullptr;
    for (Node* it : rank_to_node_) {
      for (Edge* jt : it->inedges) {
        if (score[it->id] < jt->weight ||
            (score[it->id] == jt->weight && score[pred[it->id]->id] <= score[jt->tail->id])) {
          score[it->id] = jt->weight;
          pred[it->id] = jt->tail;
        }
      }
      if (pred[it->id]) score[it->id] += score[pred[it->id]->id];
      if (!max || score[max->id] < score[it->id]) max = it;
    }
    if (!max->outedges.empty()) {
      std::vector<std::uint32_t> rank(nodes_.size(), 0);
      for (std::uint32_t i = 0; i < rank_to_node_.size(); ++i) rank[rank_to_node_[i]->id] = i;
      while (!max->outedges.empty()) max = BranchCompletion(rank[max->id], score, pred);
    }
    std::vector<Node*> path;
    while (pred[max->id]) {
      path.push_back(max);
      max = pred[max->id];
    }
    path.push_back(max);
    std::reverse(path.begin(), path.end());
    for (Node* n : path) dst += (char)decoder_[n->code];
    return dst;
  }

 private:
  Node* AddNode(std::uint32_t code) {
    nodes_.emplace_back(new Node());
    nodes_.back()->id = (std::uint32_t)nodes_.size() - 1;
    nodes_.back()->code = code;
    return nodes_.back().get();
  }
  void AddEdge(Node* tail, Node* head, std::uint32_t weight) {
    for (Edge* e : tail->outedges)
      if (e->head == head) {
        e->weight += weight;
        return;
      }
    edges_.emplace_back(new Edge());
    Edge* e = edges_.back().get();
    e->tail = tail;
    e->head = head;
    e->weight = weight;
    tail->outedges.push_back(e);
    head->inedges.push_back(e);
  }
  Node* AddSequence(const std::string& s, const std::vector<std::uint32_t>& w, std::uint32_t begin, std::uint32_t end) {
    if (begin == end) return nullptr;
    Node* prev = nullptr;
    for (std::uint32_t i = begin; i < end; ++i) {
      Node* curr = AddNode((std::uint32_t)coder_[(unsigned char)s[i]]);
      if (prev) AddEdge(prev, curr, w[i - 1] + w[i]);
      prev = curr;
    }
    return nodes_[nodes_.size() - (end - begin)].get();
  }
  void TopologicalSort() {
    rank_to_node_.clear();
    std::vector<std::uint8_t> marks(nodes_.size(), 0);
    std::vector<bool> ignored(nodes_.size(), false);
    std::stack<Node*> stack;
    for (const std::unique_ptr<Node>& it : nodes_) {
      if (marks[it->id] != 0) continue;
      stack.push(it.get());
      while (!stack.empty()) {
        Node* curr = stack.top();
        bool is_valid = true;
        if (marks[curr->id] != 2) {
          for (Edge* jt : curr->inedges)
            if (marks[jt->tail->id] != 2) {
              stack.push(jt->tail);
              is_valid = false;
            }
          if (!ignored[curr->id])
            for (Node* jt : curr->aligned_nodes)
              if (marks[jt->id] != 2) {
                stack.push(jt);
                ignored[jt->id] = true;
                is_valid = false;
              }
          if (is_valid) {
            marks[curr->id] = 2;
            if (!ignored[curr->id]) {
              rank_to_node_.push_back(curr);
              for (Node* jt : curr->aligned_nodes) rank_to_node_.push_back(jt);
            }
          } else {
            marks[curr->id] = 1;
          }
        }
        if (is_valid) stack.pop();
      }
    }
  }
  Node* BranchCompletion(std::uint32_t rank, std::vector<std::int64_t>& score, std::vector<Node*>& pred) {
    Node* start = rank_to_node_[rank];
    for (Edge* it : start->outedges)
      for (Edge* jt : it->head->inedges)
        if (jt->tail != start) score[jt->tail->id] = -1;
    Node* max = nullptr;
    for (std::uint32_t i = rank + 1; i < rank_to_node_.size(); ++i) {
      Node* it = rank_to_node_[i];
      score[it->id] = -1;
      pred[it->id] = nullptr;
      for (Edge* jt : it->inedges) {
        if (score[jt->tail->id] == -1) continue;
        if (score[it->id] < jt->weight ||
            (score[it->id] == jt->weight && score[pred[it->id]->id] <= score[jt->tail->id])) {
          score[it->id] = jt->weight;
          pred[it->id] = jt->tail;
        }
      }
      if (pred[it->id]) score[it->id] += score[pred[it->id]->id];
      if (!max || score[max->id] < score[it->id]) max = it;
    }
    return max;
  }

  std::uint32_t num_codes_;
  std::vector<std::int32_t> coder_, decoder_;
  std::uint32_t n_sequences_;
  std::vector<std::unique_ptr<Node> > nodes_;
  std::vector<std::unique_ptr<Edge> > edges_;
  std::vector<Node*> rank_to_node_;
};

// Global alignment of a sequence to the graph with linear gap cost g (g == e), match m, mismatch n.
inline Alignment AlignNW(const std::string& sequence, const Graph& graph, std::int32_t m, std::int32_t n, std::int32_t g) {
  const std::uint32_t len = (std::uint32_t)sequence.size();
  const std::vector<Graph::Node*>& r2n = graph.rank_to_node();
  if (r2n.empty() || len == 0) return Alignment();
  const std::uint64_t W = (std::uint64_t)len + 1, Hh = r2n.size() + 1;
  const std::int32_t kNeg = std::numeric_limits<std::int32_t>::min() + 1024;
  std::vector<std::int32_t> profile((std::size_t)graph.num_codes() * W);
  for (std::uint32_t c = 0; c < graph.num_codes(); ++c) {
    profile[c * W] = 0;
    for (std::uint64_t j = 1; j < W; ++j) profile[c * W + j] = (graph.decoder(c) == (unsigned char)sequence[j - 1]) ? m : n;
  }
  std::vector<std::uint32_t> rank(graph.n_nodes(), 0);
  for (std::uint32_t i = 0; i < r2n.size(); ++i) rank[r2n[i]->id] = i;
  std::vector<std::int32_t> H(Hh * W, 0);
  for (std::uint64_t j = 1; j < W; ++j) H[j] = (std::int32_t)j * g;
  for (std::uint64_t i = 1; i < Hh; ++i) {
    const Graph::Node* it = r2n[i - 1];
    if (it->inedges.empty()) {
      H[i * W] = g;
    } else {
      std::int32_t pen = kNeg;
      for (const Graph::Edge* e : it->inedges) pen = std::max(pen, H[(std::uint64_t)(rank[e->tail->id] + 1) * W]);
      H[i * W] = pen + g;
    }
  }
  std::int32_t max_score = kNeg;
  std::uint32_t max_i = 0, max_j = 0;
  for (const Graph::Node* it : r2n) {
    const std::int32_t* prof = &profile[(std::size_t)it->code * W];
    const std::uint32_t i = rank[it->id] + 1;
    std::uint32_t pred_i = it->inedges.empty() ? 0 : rank[it->inedges[0]->tail->id] + 1;
    std::int32_t* row = &H[(std::uint64_t)i * W];
    const std::int32_t* prow = &H[(std::uint64_t)pred_i * W];
    for (std::uint64_t j = 1; j < W; ++j) row[j] = std::max(prow[j - 1] + prof[j], prow[j] + g);
    for (std::size_t p = 1; p < it->inedges.size(); ++p) {
      pred_i = rank[it->inedges[p]->tail->id] + 1;
      prow = &H[(std::uint64_t)pred_i * W];
      for (std::uint64_t j = 1; j < W; ++j) row[j] = std::max(prow[j - 1] + prof[j], std::max(row[j], prow[j] + g));
    }
    for (std::uint64_t j = 1; j < W; ++j) {
      row[j] = std::max(row[j - 1] + g, row[j]);
      if (j == W - 1 && it->outedges.empty() && max_score < row[j]) {
        max_score = row[j];
        max_i = i;
        max_j = (std::uint32_t)j;
      }
    }
  }
  if (max_i == 0 && max_j == 0) return Alignment();
  Alignment aln;
  std::uint32_t i = max_i, j = max_j, prev_i = 0, prev_j = 0;
  while (!(i == 0 && j == 0)) {
    const std::int32_t Hij = H[(std::uint64_t)i * W + j];
    bool found = false;
    if (i != 0 && j != 0) {
      const Graph::Node* it = r2n[i - 1];
      const std::int32_t mc = profile[(std::size_t)it->code * W + j];
      std::uint32_t pred_i = it->inedges.empty() ? 0 : rank[it->inedges[0]->tail->id] + 1;
      if (Hij == H[(std::uint64_t)pred_i * W + (j - 1)] + mc) {
        prev_i = pred_i;
        prev_j = j - 1;
        found = true;
      } else {
        for (std::size_t p = 1; p < it->inedges.size(); ++p) {
          pred_i = rank[it->inedges[p]->tail->id] + 1;
          if (Hij == H[(std::uint64_t)pred_i * W + (j - 1)] + mc) {
            prev_i = pred_i;
            prev_j = j - 1;
            found = true;
            break;
          }
        }
      }
    }
    if (!found && i != 0) {
      const Graph::Node* it = r2n[i - 1];
      std::uint32_t pred_i = it->inedges.empty() ? 0 : rank[it->inedges[0]->tail->id] + 1;
      if (Hij == H[(std::uint64_t)pred_i * W + j] + g) {
        prev_i = pred_i;
        prev_j = j;
        found = true;
      } else {
        for (std::size_t p = 1; p < it->inedges.size(); ++p) {
          pred_i = rank[it->inedges[p]->tail->id] + 1;
          if (Hij == H[(std::uint64_t)pred_i * W + j] + g) {
            prev_i = pred_i;
            prev_j = j;
            found = true;
            break;
          }
        }
      }
    }
    if (!found && j != 0 && Hij == H[(std::uint64_t)i * W + j - 1] + g) {
      prev_i = i;
      prev_j = j - 1;
      found = true;
    }
    aln.push_back(std::make_pair(i == prev_i ? -1 : (std::int32_t)r2n[i - 1]->id, j == prev_j ? -1 : (std::int32_t)j - 1));
    i = prev_i;
    j = prev_j;
  }
  std::reverse(aln.begin(), aln.end());
  return aln;
}

inline std::string consensus(const std::vector<std::string>& seqs) {
  Graph graph;
  for (const std::string& s : seqs) graph.AddAlignment(AlignNW(s, graph, 1, -1, -1), s);
  return graph.GenerateConsensus();
}

}  // namespace ltr_poa_oracle
#endif
