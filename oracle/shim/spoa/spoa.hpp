// Compile-only stand-in for rvaser/spoa (un-vendored, unpinned in the reference Makefile:96-103).  The oracle's
// inputs never reach HaplotypeGenerator::poa (every allele has >= 2 supporting reads); reaching it aborts loudly.
#ifndef LTR_SHIM_SPOA_HPP
#define LTR_SHIM_SPOA_HPP
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>
namespace spoa {
enum class AlignmentType { kSW, kNW, kOV };
using Alignment = std::vector<std::pair<std::int32_t, std::int32_t>>;
class Graph;
class AlignmentEngine {
 public:
  static std::unique_ptr<AlignmentEngine> Create(AlignmentType, std::int8_t, std::int8_t, std::int8_t) {
#ifdef LTR_SPOA_THROW  // oracle/hapgen_driver.cpp: "this region needs the assembly" is an answer, not a crash
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
