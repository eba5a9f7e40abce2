// Stand-in for rvaser/spoa (un-vendored, unpinned in the reference Makefile:96-103).  Three build modes:
//   default                    compile-only; reaching HaplotypeGenerator::poa aborts loudly (the hot-path oracle's inputs never
//                              do: every allele has >= 2 supporting reads)
//   -DLTR_SPOA_THROW           oracle/hapgen_driver.cpp: "this region needs the assembly" is an answer, not a crash
//   -DLTR_SPOA_RESTATEMENT     the spoa names the reference binds are served by oracle/poa_restatement.hpp (restatement of
//                              spoa's published algorithm, parity UNPINNED: see its header), so that the reference's own
//                              clustering / refinement / support logic around the consensus runs to completion
#ifndef LTR_SHIM_SPOA_HPP
#define LTR_SHIM_SPOA_HPP
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>
#ifdef LTR_SPOA_RESTATEMENT
#include "../../poa_restatement.hpp"
namespace spoa {
enum class AlignmentType { kSW, kNW, kOV };
using Alignment = ltr_poa_oracle::Alignment;
class Graph : public ltr_poa_oracle::Graph {};
class AlignmentEngine {
 public:
  static std::unique_ptr<AlignmentEngine> Create(AlignmentType type, std::int8_t m, std::int8_t n, std::int8_t g) {
    if (type != AlignmentType::kNW) std::abort();  // the only configuration the reference uses
    std::unique_ptr<AlignmentEngine> e(new AlignmentEngine());
    e->m_ = m;
    e->n_ = n;
    e->g_ = g;
    return e;
  }
  Alignment Align(const std::string& s, const Graph& graph, std::int32_t* = nullptr) {
    return ltr_poa_oracle::AlignNW(s, graph, m_, n_, g_);
  }

 private:
  std::int32_t m_, n_, g_;
};
}  // namespace spoa
#else
namespace spoa {
enum class AlignmentType { kSW, kNW, kOV };
using Alignment = std::vector<std::pair<std::int32_t, std::int32_t>>;
class Graph;
class AlignmentEngine {
 public:
  static std::unique_ptr<AlignmentEngine> Create(AlignmentType, std::int8_t, std::int8_t, std::int8_t) {
#ifdef LTR_SPOA_THROW
    throw 1;
#endif
    std::fprintf(stderr, "oracle/_ref: spoa stub reached (POA is outside the hot path)\n");
    std::abort();
  }
  static std::unique_ptr<AlignmentEngine> Create(AlignmentType, std::int8_t, std::int8_t, std::int8_t, std::int8_t) { std::abort(); }
  Alignment Align(const std::string&, const Graph&, std::int32_t* = nullptr) { std::abort(); }
};
class Graph {
 public:
  void AddAlignment(const Alignment&, const std::string&, std::uint32_t = 1) { std::abort(); }
  std::string GenerateConsensus() { std::abort(); }
  std::string GenerateConsensus(std::int32_t) { std::abort(); }
};
}  // namespace spoa
#endif
#endif
