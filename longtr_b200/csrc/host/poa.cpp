// poa.cpp -- partial-order consensus for the candidate-allele assembly, and the thresholded edit distance of its clustering.
//
// The reference replaces every cluster of reads that have no candidate allele by a consensus it obtains from spoa
// (src/SeqAlignment/HaplotypeGenerator.cpp:167-199: AlignmentEngine::Create(kNW, 1, -1, -1), Align + AddAlignment per
// sequence, GenerateConsensus).  spoa (github.com/rvaser/spoa) is cloned at build time without a version pin (reference
// Makefile:96-103) and exists neither in /root/reference nor in this image, so this file restates the algorithm spoa 4.x
// publishes -- PARITY UNPINNED: no output of the real library exists to hold it to.  What IS held: this implementation (flat
// index arrays, bytes instead of code tables, one matrix reused across calls) against an independent restatement in spoa's
// own object structure (oracle/poa_restatement.hpp) on seeded clusters, and the reference's whole assembly branch compiled in
// place on top of that restatement (tests/test_assembly.py).
//
// The algorithm, in the terms used below:
//   graph     one node per (alignment column, base); nodes of one column are each other's `peers`; an edge tail -> head
//             carries the number of sequence steps through it, twice (weight 1 per base, both ends added)
//   order     depth-first topological order in which a node is emitted together with its peers, so that the rows of one
//             column are adjacent; rebuilt after every sequence
//   align     global alignment of the sequence to the graph, one matrix row per node in that order, a row's predecessors are
//             the rows of its in-edges' tails (row 0 for a node without in-edges); linear gap -1, match +1, mismatch -1; the
//             end row is the first sink row (in order) with the highest score in the last column; back-track prefers a
//             diagonal step, then a step along the graph, then a step along the sequence, predecessors in edge order
//   merge     an aligned base joins the node it is aligned to, or that node's peer with the same base, or becomes a new
//             peer; unaligned bases become new nodes; consecutive bases are linked
//   consensus heaviest bundle: every node picks the in-edge of highest weight (ties: the tail with the higher score, later
//             edge wins on equality), score = weight + score of that tail (-1 for a node without in-edges); the best node
//             ends the path, and if it is not a sink the path is completed branch by branch (scores of competing tails are
//             voided and the ranks behind it re-scored)
#include "poa.h"

#include <algorithm>
#include <limits>

namespace ltr {

void PoaGraph::clear() {
  code_.clear();
  in_.clear();
  out_.clear();
  peers_.clear();
  tail_.clear();
  head_.clear();
  weight_.clear();
  order_.clear();
  rank_.clear();
}

uint32_t PoaGraph::new_node(uint8_t code) {
  code_.push_back(code);
  in_.emplace_back();
  out_.emplace_back();
  peers_.emplace_back();
  return (uint32_t)code_.size() - 1;
}

void PoaGraph::link(uint32_t tail, uint32_t head, uint32_t weight) {
  for (uint32_t e : out_[tail])
    if (head_[e] == head) {
      weight_[e] += weight;
      return;
    }
  const uint32_t e = (uint32_t)tail_.size();
  tail_.push_back(tail);
  head_.push_back(head);
  weight_.push_back(weight);
  out_[tail].push_back(e);
  in_[head].push_back(e);
}

// Nodes are tried in index order; a node is emitted once all tails of its in-edges and -- unless it was itself reached as a
// peer -- all of its peers are done, and a node emitted on its own account takes its peers along.
void PoaGraph::sort_nodes() {
  const uint32_t N = n_nodes();
  order_.clear();
  order_.reserve(N);
  enum { kNew = 0, kOpen = 1, kDone = 2 };
  std::vector<uint8_t> state(N, kNew), as_peer(N, 0);
  std::vector<uint32_t> todo;
  for (uint32_t root = 0; root < N; ++root) {
    if (state[root] != kNew) continue;
    todo.push_back(root);
    while (!todo.empty()) {
      const uint32_t v = todo.back();
      bool ready = true;
      if (state[v] != kDone) {
        for (uint32_t e : in_[v])
          if (state[tail_[e]] != kDone) {
            todo.push_back(tail_[e]);
            ready = false;
          }
        if (!as_peer[v])
          for (uint32_t p : peers_[v])
            if (state[p] != kDone) {
              todo.push_back(p);
              as_peer[p] = 1;
              ready = false;
            }
        if (ready) {
          state[v] = kDone;
          if (!as_peer[v]) {
            order_.push_back(v);
            for (uint32_t p : peers_[v]) order_.push_back(p);
          }
        } else {
          state[v] = kOpen;
        }
      }
      if (ready) todo.pop_back();
    }
  }
  rank_.assign(N, 0);
  for (uint32_t i = 0; i < N; ++i) rank_[order_[i]] = i;
}

// One matrix row: columns 1 .. len of `row` from the predecessor rows (`preds`: n_pred row pointers, the first one first),
//   row[j] = max over predecessors p of max(p[j-1] + (seq[j-1] == c ? +1 : -1), p[j] - 1), then row[j] = max(row[j], row[j-1] - 1).
// The second pass is a running maximum of row[j] + j.  A plain version for either cell type and two AVX2 versions (runtime
// dispatch): 8 int32 columns per step, or 16 int16 columns per step when every score of the matrix fits (see align); the
// running maximum is an in-register scan plus a carried lane.  Same integers either way.
template <typename T>
static void poa_row_plain(T* row, const T* const* preds, size_t n_pred, const uint8_t* seq, size_t len, uint8_t c) {
  const T* p0 = preds[0];
  for (size_t j = 1; j <= len; ++j) row[j] = (T)std::max(p0[j - 1] + (seq[j - 1] == c ? 1 : -1), p0[j] - 1);
  for (size_t k = 1; k < n_pred; ++k) {
    const T* pk = preds[k];
    for (size_t j = 1; j <= len; ++j)
      row[j] = (T)std::max(pk[j - 1] + (seq[j - 1] == c ? 1 : -1), std::max((int)row[j], pk[j] - 1));
  }
  for (size_t j = 1; j <= len; ++j) row[j] = (T)std::max(row[j - 1] - 1, (int)row[j]);
}

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define LTR_POA_AVX2 1
// Rows are padded: columns up to the next multiple of 16 past len exist in every row and in seq (their values are never used).
__attribute__((target("avx2"))) static void poa_row_avx2(int32_t* row, const int32_t* const* preds, size_t n_pred,
                                                          const uint8_t* seq, size_t len, uint8_t c, int32_t) {
  const __m256i vc = _mm256_set1_epi32((int)c), one = _mm256_set1_epi32(1), two = _mm256_set1_epi32(2);
  const __m256i neg = _mm256_set1_epi32(INT32_MIN / 2);
  const __m256i sh1 = _mm256_setr_epi32(0, 0, 1, 2, 3, 4, 5, 6), sh2 = _mm256_setr_epi32(0, 0, 0, 1, 2, 3, 4, 5),
                sh4 = _mm256_setr_epi32(0, 0, 0, 0, 0, 1, 2, 3), last = _mm256_set1_epi32(7);
  __m256i jv = _mm256_setr_epi32(1, 2, 3, 4, 5, 6, 7, 8);
  __m256i carry = _mm256_set1_epi32(row[0]);  // running maximum of row[j] + j, starting from column 0
  const __m256i eight = _mm256_set1_epi32(8);
  for (size_t j = 1; j <= len; j += 8) {
    const __m256i ch = _mm256_cvtepu8_epi32(_mm_loadl_epi64((const __m128i*)(seq + j - 1)));
    const __m256i s = _mm256_sub_epi32(_mm256_and_si256(_mm256_cmpeq_epi32(ch, vc), two), one);  // +1 / -1
    __m256i r = _mm256_max_epi32(_mm256_add_epi32(_mm256_loadu_si256((const __m256i*)(preds[0] + j - 1)), s),
                                 _mm256_sub_epi32(_mm256_loadu_si256((const __m256i*)(preds[0] + j)), one));
    for (size_t k = 1; k < n_pred; ++k) {
      const __m256i d = _mm256_add_epi32(_mm256_loadu_si256((const __m256i*)(preds[k] + j - 1)), s);
      const __m256i u = _mm256_sub_epi32(_mm256_loadu_si256((const __m256i*)(preds[k] + j)), one);
      r = _mm256_max_epi32(d, _mm256_max_epi32(r, u));
    }
    __m256i x = _mm256_add_epi32(r, jv);
    x = _mm256_max_epi32(x, _mm256_blend_epi32(_mm256_permutevar8x32_epi32(x, sh1), neg, 0x01));
    x = _mm256_max_epi32(x, _mm256_blend_epi32(_mm256_permutevar8x32_epi32(x, sh2), neg, 0x03));
    x = _mm256_max_epi32(x, _mm256_blend_epi32(_mm256_permutevar8x32_epi32(x, sh4), neg, 0x0f));
    x = _mm256_max_epi32(x, carry);
    carry = _mm256_permutevar8x32_epi32(x, last);
    _mm256_storeu_si256((__m256i*)(row + j), _mm256_sub_epi32(x, jv));
    jv = _mm256_add_epi32(jv, eight);
  }
}
// 16 columns per step.  The scan runs on row[j] + j + bias with bias > |lowest score|, so that the zeros shifted in by the
// lane shifts act as minus infinity; the caller guarantees bias + 2 len < 2^15.
__attribute__((target("avx2"))) static void poa_row_avx2(int16_t* row, const int16_t* const* preds, size_t n_pred,
                                                          const uint8_t* seq, size_t len, uint8_t c, int32_t bias) {
  const __m128i vc = _mm_set1_epi8((char)c);
  const __m256i one = _mm256_set1_epi16(1), minus1 = _mm256_set1_epi16(-1), sixteen = _mm256_set1_epi16(16);
  const __m256i vbias = _mm256_set1_epi16((short)bias);
  const __m256i top = _mm256_setr_epi8(14, 15, 14, 15, 14, 15, 14, 15, 14, 15, 14, 15, 14, 15, 14, 15,
                                       14, 15, 14, 15, 14, 15, 14, 15, 14, 15, 14, 15, 14, 15, 14, 15);
  __m256i jv = _mm256_setr_epi16(1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16);
  __m256i carry = _mm256_set1_epi16((short)(row[0] + bias));
  for (size_t j = 1; j <= len; j += 16) {
    const __m128i eq = _mm_cmpeq_epi8(_mm_loadu_si128((const __m128i*)(seq + j - 1)), vc);  // 0xff where the base matches
    const __m256i m = _mm256_cvtepi8_epi16(eq);                                               // -1 / 0
    const __m256i s = _mm256_sub_epi16(minus1, _mm256_add_epi16(m, m));                       // +1 / -1
    __m256i r = _mm256_max_epi16(_mm256_add_epi16(_mm256_loadu_si256((const __m256i*)(preds[0] + j - 1)), s),
                                 _mm256_sub_epi16(_mm256_loadu_si256((const __m256i*)(preds[0] + j)), one));
    for (size_t k = 1; k < n_pred; ++k) {
      const __m256i d = _mm256_add_epi16(_mm256_loadu_si256((const __m256i*)(preds[k] + j - 1)), s);
      const __m256i u = _mm256_sub_epi16(_mm256_loadu_si256((const __m256i*)(preds[k] + j)), one);
      r = _mm256_max_epi16(d, _mm256_max_epi16(r, u));
    }
    __m256i x = _mm256_add_epi16(_mm256_add_epi16(r, jv), vbias);  // > 0
    __m256i lo = _mm256_permute2x128_si256(x, x, 0x08);            // [0, x.low]
    x = _mm256_max_epi16(x, _mm256_alignr_epi8(x, lo, 14));        // shifted by 1 lane, zeros in
    lo = _mm256_permute2x128_si256(x, x, 0x08);
    x = _mm256_max_epi16(x, _mm256_alignr_epi8(x, lo, 12));        // by 2
    lo = _mm256_permute2x128_si256(x, x, 0x08);
    x = _mm256_max_epi16(x, _mm256_alignr_epi8(x, lo, 8));         // by 4
    lo = _mm256_permute2x128_si256(x, x, 0x08);
    x = _mm256_max_epi16(x, lo);                                   // by 8
    x = _mm256_max_epi16(x, carry);
    carry = _mm256_shuffle_epi8(_mm256_permute2x128_si256(x, x, 0x11), top);  // lane 15 everywhere
    _mm256_storeu_si256((__m256i*)(row + j), _mm256_sub_epi16(_mm256_sub_epi16(x, jv), vbias));
    jv = _mm256_add_epi16(jv, sixteen);
  }
}
#endif

template <typename T>
void PoaGraph::align_cells(const uint8_t* seq_in, uint32_t len, std::vector<int32_t>& aln_node, std::vector<int32_t>& aln_pos) {
  const uint32_t N = n_nodes();
  const size_t W = ((size_t)len + 16) / 16 * 16 + 16;  // row stride: column 0, len columns, padding to whole steps of 16
  const int32_t gap = -1, hit = 1, miss = -1;
  const size_t need = ((size_t)(N + 1) * W * sizeof(T) + sizeof(int32_t) - 1) / sizeof(int32_t);
  if (H_.size() < need) H_.resize(need);
  T* H = reinterpret_cast<T*>(H_.data());
  std::vector<uint8_t> padded(W + 16, 0);
  std::copy(seq_in, seq_in + len, padded.begin());
  const uint8_t* seq = padded.data();
  const int32_t bias = (int32_t)N + (int32_t)len + 2;  // every score is >= -(N + len)
#ifdef LTR_POA_AVX2
  static const bool use_avx2 = __builtin_cpu_supports("avx2");
#endif
  for (size_t j = 0; j < W; ++j) H[j] = (T)(-(int32_t)j);
  auto row_of_tail = [&](uint32_t e) { return (size_t)rank_[tail_[e]] + 1; };
  int32_t best = std::numeric_limits<int32_t>::min();
  uint32_t best_row = 0;
  std::vector<const T*> preds;
  for (uint32_t r = 0; r < N; ++r) {
    const uint32_t v = order_[r];
    T* row = H + (size_t)(r + 1) * W;
    const std::vector<uint32_t>& in = in_[v];
    // column 0: one graph step below the best predecessor
    preds.clear();
    if (in.empty()) {
      row[0] = (T)gap;
      preds.push_back(H);
    } else {
      int32_t top = std::numeric_limits<int32_t>::min() + 1024;
      for (uint32_t e : in) {
        preds.push_back(H + row_of_tail(e) * W);
        top = std::max(top, (int32_t)preds.back()[0]);
      }
      row[0] = (T)(top + gap);
    }
#ifdef LTR_POA_AVX2
    if (use_avx2) poa_row_avx2(row, preds.data(), preds.size(), seq, len, code_[v], bias);
    else
#endif
      poa_row_plain<T>(row, preds.data(), preds.size(), seq, len, code_[v]);
    if (out_[v].empty() && best < (int32_t)row[len]) {
      best = row[len];
      best_row = r + 1;
    }
  }
  // back-track from (best_row, len) to (0, 0)
  size_t i = best_row, j = len;
  while (i != 0 || j != 0) {
    const int32_t here = H[i * W + j];
    size_t pi = i, pj = j;
    bool found = false;
    if (i != 0) {
      const uint32_t v = order_[i - 1];
      const std::vector<uint32_t>& in = in_[v];
      const size_t n_pred = in.empty() ? 1 : in.size();
      if (j != 0) {
        const int32_t s = seq[j - 1] == code_[v] ? hit : miss;
        for (size_t k = 0; k < n_pred && !found; ++k) {
          const size_t p = in.empty() ? 0 : row_of_tail(in[k]);
          if (here == (int32_t)H[p * W + (j - 1)] + s) {
            pi = p;
            pj = j - 1;
            found = true;
          }
        }
      }
      for (size_t k = 0; k < n_pred && !found; ++k) {
        const size_t p = in.empty() ? 0 : row_of_tail(in[k]);
        if (here == (int32_t)H[p * W + j] + gap) {
          pi = p;
          pj = j;
          found = true;
        }
      }
    }
    if (!found && j != 0 && here == (int32_t)H[i * W + j - 1] + gap) {
      pj = j - 1;
      found = true;
    }
    if (!found) {  // cannot happen: every cell is the maximum of the three moves tested above
      aln_node.clear();
      aln_pos.clear();
      return;
    }
    aln_node.push_back(i == pi ? -1 : (int32_t)order_[i - 1]);
    aln_pos.push_back(j == pj ? -1 : (int32_t)j - 1);
    i = pi;
    j = pj;
  }
  std::reverse(aln_node.begin(), aln_node.end());
  std::reverse(aln_pos.begin(), aln_pos.end());
}

void PoaGraph::align(const uint8_t* seq, uint32_t len, std::vector<int32_t>& aln_node, std::vector<int32_t>& aln_pos) {
  aln_node.clear();
  aln_pos.clear();
  const uint32_t N = n_nodes();
  if (N == 0 || len == 0) return;
  // scores lie in [-(N + len), len]; the 16-bit row kernel also forms score + column + bias with bias = N + len + 2
  if ((uint64_t)N + 3ull * len + 64 < 32000ull && !force_wide_) align_cells<int16_t>(seq, len, aln_node, aln_pos);
  else align_cells<int32_t>(seq, len, aln_node, aln_pos);
}

void PoaGraph::add(const uint8_t* seq, uint32_t len) {
  if (len == 0) return;
  align(seq, len, aln_node_, aln_pos_);
  const uint32_t w2 = 2;  // weight 1 at either end of a step
  if (aln_node_.empty()) {  // first sequence: a chain of its own
    uint32_t prev = new_node(seq[0]);
    for (uint32_t k = 1; k < len; ++k) {
      const uint32_t cur = new_node(seq[k]);
      link(prev, cur, w2);
      prev = cur;
    }
    sort_nodes();
    return;
  }
  // a global alignment covers every base, so there is no unaligned head or tail of the sequence to chain up separately
  int64_t prev = -1;
  for (size_t k = 0; k < aln_pos_.size(); ++k) {
    const int32_t pos = aln_pos_[k];
    if (pos < 0) continue;
    const uint8_t c = seq[pos];
    const int32_t at = aln_node_[k];
    uint32_t cur;
    if (at < 0) {
      cur = new_node(c);
    } else if (code_[(size_t)at] == c) {
      cur = (uint32_t)at;
    } else {
      int64_t same = -1;
      for (uint32_t p : peers_[(size_t)at])
        if (code_[p] == c) {
          same = p;
          break;
        }
      if (same >= 0) {
        cur = (uint32_t)same;
      } else {
        cur = new_node(c);
        const std::vector<uint32_t> column = peers_[(size_t)at];  // copy: the lists below are edited
        for (uint32_t p : column) {
          peers_[p].push_back(cur);
          peers_[cur].push_back(p);
        }
        peers_[(size_t)at].push_back(cur);
        peers_[cur].push_back((uint32_t)at);
      }
    }
    if (prev >= 0) link((uint32_t)prev, cur, w2);
    prev = cur;
  }
  sort_nodes();
}

uint32_t PoaGraph::complete_branch(uint32_t rank, std::vector<int64_t>& score, std::vector<int32_t>& pred) const {
  const uint32_t start = order_[rank];
  for (uint32_t e : out_[start])
    for (uint32_t f : in_[head_[e]])
      if (tail_[f] != start) score[tail_[f]] = -1;
  int64_t best = -1;
  for (uint32_t r = rank + 1; r < order_.size(); ++r) {
    const uint32_t v = order_[r];
    score[v] = -1;
    pred[v] = -1;
    for (uint32_t e : in_[v]) {
      const uint32_t t = tail_[e];
      if (score[t] == -1) continue;
      if (score[v] < weight_[e] || (score[v] == weight_[e] && score[(size_t)pred[v]] <= score[t])) {
        score[v] = weight_[e];
        pred[v] = (int32_t)t;
      }
    }
    if (pred[v] >= 0) score[v] += score[(size_t)pred[v]];
    if (best < 0 || score[(size_t)best] < score[v]) best = v;
  }
  return (uint32_t)best;
}

void PoaGraph::consensus(std::string& out) {
  out.clear();
  const uint32_t N = n_nodes();
  if (N == 0) return;
  std::vector<int32_t> pred(N, -1);
  std::vector<int64_t> score(N, -1);
  int64_t best = -1;
  for (uint32_t r = 0; r < N; ++r) {
    const uint32_t v = order_[r];
    for (uint32_t e : in_[v]) {
      const uint32_t t = tail_[e];
      if (score[v] < weight_[e] || (score[v] == weight_[e] && score[(size_t)pred[v]] <= score[t])) {
        score[v] = weight_[e];
        pred[v] = (int32_t)t;
      }
    }
    if (pred[v] >= 0) score[v] += score[(size_t)pred[v]];
    if (best < 0 || score[(size_t)best] < score[v]) best = v;
  }
  uint32_t end = (uint32_t)best;
  while (!out_[end].empty()) end = complete_branch(rank_[end], score, pred);
  for (int64_t v = end; v >= 0; v = pred[(size_t)v]) out.push_back((char)code_[(size_t)v]);
  std::reverse(out.begin(), out.end());
}

// ---- thresholded edit distance ------------------------------------------------------------------------------------------
// Bit-vector recurrence of the unit-cost edit distance matrix (rows = cent_seq, columns = read_seq, first row and column
// 0, 1, 2, ...), 64 rows per word, blocks chained through the horizontal delta of their last row.
int edit_distance(const std::string& a, const std::string& b) {
  const int n = (int)a.size(), m = (int)b.size();
  if (n == 0 || m == 0) return n + m;
  const int nb = (n + 63) / 64;
  static thread_local std::vector<uint64_t> peq, pv, mv;
  peq.assign((size_t)nb * 256, 0);
  pv.assign((size_t)nb, ~0ull);
  mv.assign((size_t)nb, 0ull);
  for (int i = 0; i < n; ++i) peq[(size_t)(i / 64) * 256 + (uint8_t)a[(size_t)i]] |= 1ull << (i % 64);
  const int last_bit = (n - 1) % 64;
  int score = n;
  for (int j = 0; j < m; ++j) {
    const uint8_t c = (uint8_t)b[(size_t)j];
    int hin = 1;  // first row: dp[0][j] = j
    for (int k = 0; k < nb; ++k) {
      uint64_t eq = peq[(size_t)k * 256 + c];
      const uint64_t pvk = pv[(size_t)k], mvk = mv[(size_t)k];
      const uint64_t xv = eq | mvk;
      if (hin < 0) eq |= 1ull;
      const uint64_t xh = (((eq & pvk) + pvk) ^ pvk) | eq;
      uint64_t ph = mvk | ~(xh | pvk);
      uint64_t mh = pvk & xh;
      const int ob = (k == nb - 1) ? last_bit : 63;
      const int hout = (int)((ph >> ob) & 1ull) - (int)((mh >> ob) & 1ull);
      ph = (ph << 1) | (hin > 0 ? 1ull : 0ull);
      mh = (mh << 1) | (hin < 0 ? 1ull : 0ull);
      pv[(size_t)k] = mh | ~(xv | ph);
      mv[(size_t)k] = ph & xv;
      hin = hout;
    }
    score += hin;
  }
  return score;
}

// The same distance when it is at most k, k + 1 otherwise (Ukkonen's cut-off on the block recurrence).  A cell (i, j) with
// a value <= k has |i - j| <= k, and so has every cell of an optimal path to it: column j only needs the blocks that hold
// rows j - k .. j + k.  Blocks above the band are dropped (the first kept block is fed the horizontal delta +1), blocks below
// it are switched on when the band reaches them, with all vertical deltas +1; both model cells outside the band by values that
// are not below the true ones, which leaves every value <= k inside the band exact.
int bounded_edit_distance(const std::string& a, const std::string& b, int k) {
  const int n = (int)a.size(), m = (int)b.size();
  if (std::abs(n - m) > k) return k + 1;
  if (n == 0 || m == 0) return n + m;
  const int nb = (n + 63) / 64;
  static thread_local std::vector<uint64_t> peq, pv, mv;
  peq.assign((size_t)nb * 256, 0);
  pv.assign((size_t)nb, ~0ull);
  mv.assign((size_t)nb, 0ull);
  for (int i = 0; i < n; ++i) peq[(size_t)(uint8_t)a[(size_t)i] * (size_t)nb + (size_t)(i / 64)] |= 1ull << (i % 64);
  int y = (std::min(n, std::max(k, 1)) - 1) / 64;  // last block switched on (rows 1 .. k of column 0)
  long long score = 64ll * (y + 1);                // value at the bottom row of block y
  for (int j = 1; j <= m; ++j) {
    const uint64_t* eqc = peq.data() + (size_t)(uint8_t)b[(size_t)j - 1] * (size_t)nb;
    const int f = (std::max(1, j - k) - 1) / 64;
    const int want = (int)((std::min<long long>(n, (long long)j + k) - 1) / 64);
    while (y < want) {
      ++y;
      pv[(size_t)y] = ~0ull;
      mv[(size_t)y] = 0ull;
      score += 64;
    }
    int hin = 1;
    for (int blk = f; blk <= y; ++blk) {
      uint64_t eq = eqc[blk];
      const uint64_t pvk = pv[(size_t)blk], mvk = mv[(size_t)blk];
      const uint64_t xv = eq | mvk;
      if (hin < 0) eq |= 1ull;
      const uint64_t xh = (((eq & pvk) + pvk) ^ pvk) | eq;
      uint64_t ph = mvk | ~(xh | pvk);
      uint64_t mh = pvk & xh;
      const int hout = (int)(ph >> 63) - (int)(mh >> 63);
      ph = (ph << 1) | (hin > 0 ? 1ull : 0ull);
      mh = (mh << 1) | (hin < 0 ? 1ull : 0ull);
      pv[(size_t)blk] = mh | ~(xv | ph);
      mv[(size_t)blk] = ph & xv;
      hin = hout;
    }
    score += hin;
  }
  // y == nb - 1: |n - m| <= k.  Rows n + 1 .. 64 nb of the last block are padding: take their vertical deltas off again.
  const int pad = 64 * nb - n;
  if (pad > 0) {
    const uint64_t mask = ~0ull << (64 - pad);
    score -= __builtin_popcountll(pv[(size_t)nb - 1] & mask);
    score += __builtin_popcountll(mv[(size_t)nb - 1] & mask);
  }
  return score <= k ? (int)score : k + 1;
}

int thresholded_from_distance(int n, int m, int distance, int T) {
  if (std::abs(n - m) > T) return T + 1;  // :203-206
  if (n == 0) return m;                   // no row is visited
  if (m == 0) return T + 1;               // every row's minimum keeps its starting value of 1000 (> T)
  return distance < T ? distance : (distance == T ? T : T + 1);
}

int thresholded_edit_distance(const std::string& a, const std::string& b, int T) {
  const int n = (int)a.size(), m = (int)b.size();
  if (std::abs(n - m) > T || n == 0 || m == 0) return thresholded_from_distance(n, m, 0, T);
  return thresholded_from_distance(n, m, edit_distance(a, b), T);
}

}  // namespace ltr
