// viterbi_host.h -- host-side preparation shared by the C-ABI implementation and by the
// CPU lane emulator of the unit tests: launch constants, boundary tables, task planning.
#pragma once
#include <stdint.h>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "longtr_b200.h"
#include "band_core.cuh"
#include "viterbi_core.cuh"

namespace ltr {

static const int kMaxRowsPerLane = 16;  // largest K instantiated (32*K rows per strip)

struct HostConsts {
  VitConsts C;  // tabI/tabD left NULL: caller points them at host or device copies
  std::vector<double> tabI, tabD;
};

// Constants and boundary tables of align_seq_to_hap (reference HapAligner.cpp:260-280).
// The prefix sums are accumulated by repeated addition exactly like the reference's
// `left_prob += ...` so that non-integer parameters round identically.
inline void make_consts(const ltr_params& p, int tab_len, HostConsts& out) {
  const float MATCH = (float)-0.000100005;  // HapAligner.cpp:261
  const float MISMATCH = -9.0f;             // HapAligner.cpp:260
  VitConsts& C = out.C;
  C.m2m = (double)p.match_match; C.d2m = (double)p.del_match; C.i2m = (double)p.ins_match;
  C.m2i = (double)p.match_ins;   C.i2i = (double)p.ins_ins;
  C.m2d = (double)p.match_del;   C.d2d = (double)p.del_del;
  C.match = (double)MATCH;
  C.mismatch = (double)MISMATCH;
  C.imp = -1000000000.0;
  C.d2d_f = p.del_del;
  C.cut = 35 - p.indel_flank_len;
  if (tab_len < 2) tab_len = 2;
  out.tabI.assign(tab_len, 0.0);
  out.tabD.assign(tab_len, 0.0);
  const float mpi = MATCH + p.match_ins;  // float addition first (HapAligner.cpp:277)
  double left = 0.0;
  out.tabI[0] = C.imp;
  for (int i = 1; i < tab_len; ++i) {
    out.tabI[i] = (double)mpi + left;
    left += (double)p.ins_ins;
  }
  left = 0.0;
  out.tabD[0] = C.imp;
  for (int j = 1; j < tab_len; ++j) {
    out.tabD[j] = (double)p.match_del + left;
    left += (double)p.del_del;
  }
  // MODE_FAST certificate (viterbi_core.cuh): final score > -600 + slack
  //   slack = 2|MISMATCH| + |M2M| + |I2M| + 2|D2D| + 0.01
  C.fast_thr = -600.0 + (2.0 * 9.0 + std::fabs((double)p.match_match) + std::fabs((double)p.ins_match) +
                         2.0 * std::fabs((double)p.del_del) + 0.01);
  C.tabI = nullptr;
  C.tabD = nullptr;
  C.tab_len = tab_len;
}

// The final-score certificate of MODE_FAST holds when every unit of band offset costs the path at least |D2D|:
// all parameters <= 0 and |I2I|, |M2I|, |M2D| >= |D2D| (Dindel defaults and the ONT-like set satisfy it).
// Otherwise every pair is evaluated by the exact kernel.
inline bool fast_certificate_valid(const ltr_params& p) {
  const float v[7] = {p.ins_ins, p.ins_match, p.del_del, p.del_match, p.match_match, p.match_ins, p.match_del};
  for (int i = 0; i < 7; ++i)
    if (!(v[i] <= 0.0f)) return false;
  const double d = std::fabs((double)p.del_del);
  return std::fabs((double)p.ins_ins) >= d && std::fabs((double)p.match_ins) >= d && std::fabs((double)p.match_del) >= d &&
         d * 4096.0 < 1e6;
}

// Row class of a haplotype with n DP rows/columns (n = trimmed length): K rows per lane.
inline int rows_per_lane(int n, int kmax) { return rows_per_lane_hd(n, kmax); }

// Array of PODs whose elements are all written before they are read: no zero fill (these are tens of MB per batch).
template <typename T>
struct PodArray {
  std::unique_ptr<T[]> p;
  size_t n = 0;
  T* ext = nullptr;  // caller-provided storage (the C ABI hands out pinned host memory), not owned
  void resize_uninit(size_t count, void* storage = nullptr) {
    ext = static_cast<T*>(storage);
    if (!ext) p.reset(new T[count ? count : 1]);
    n = count;
  }
  T* data() { return ext ? ext : p.get(); }
  const T* data() const { return ext ? ext : p.get(); }
  size_t size() const { return n; }
  T& operator[](size_t i) { return data()[i]; }
  const T& operator[](size_t i) const { return data()[i]; }
};

// band_w < 0: banding off; 0: automatic margin; > 0: that many diagonals.  The band certificate needs every transition
// parameter <= 0 and the same parameter condition as the final-score certificate (the bail-out is certified from F as well).
// Share of the announced error cost per base that the automatic margin budgets for (LTR_BAND_BUDGET overrides: tuning).
inline double band_budget_factor() {
  static const double f = [] {
    const char* e = getenv("LTR_BAND_BUDGET");
    const double v = e ? atof(e) : 0.45;
    return (v > 0.0 && v < 10.0) ? v : 0.45;
  }();
  return f;
}

inline BandPolicy band_policy(const ltr_params& p, int band_w) {
  BandPolicy b;
  if (band_w < 0 || !fast_certificate_valid(p)) return b;
  b.gap.open = std::min(std::fabs((double)p.match_del), std::fabs((double)p.match_ins));
  b.gap.ext = std::min(std::fabs((double)p.del_del), std::fabs((double)p.ins_ins));
  if (b.gap.open < b.gap.ext) b.gap.open = b.gap.ext;
  if (!(b.gap.ext >= 0.05)) return b;
  // Automatic margin (performance heuristic, results never depend on it): the band must be certifiable for a path that
  // pays, beyond the gap the length difference forces (open + |de| ext), an error budget B(n):
  //   2 open + (|de| + 2w - 1) ext >= open + |de| ext + B   <=>   w >= ((B - open) / ext + 1) / 2.
  // B(n) = 23 log units (two mismatches and a bit; config 3 on a B200: margins of 4 / 7 / 9 / 12 diagonals give
  // 1.89 / 1.84 / 1.78 / 1.60 M loci/s -- flat; 7 it is) + n * 0.6 * r, where r is the error cost per base the
  // transition parameters themselves announce: gaps open with probability 1 - exp(M2M) per base and cost about
  // open + close + ext, and substitutions (cost 9) are assumed as frequent as gaps.  r is 0.002 for the Dindel defaults
  // and 0.30 for the ONT-like set (config 4: margin ~80 for a 760-base haplotype, the measured optimum is 64-96).
  const double p_gap = 1.0 - std::exp((double)p.match_match);
  const double close = std::min(std::fabs((double)p.del_match), std::fabs((double)p.ins_match));
  b.budget0 = 23.0;
  b.budget_per_row = band_budget_factor() * p_gap * (b.gap.open + close + b.gap.ext + 9.0);
  b.w_fixed = band_w > 0 ? band_w : 0;
  static const int share = [] {  // LTR_BAND_MAX_SHARE: tuning
    const char* e = getenv("LTR_BAND_MAX_SHARE");
    const int v = e ? atoi(e) : 80;
    return (v >= 10 && v <= 100) ? v : 80;
  }();
  b.max_share_pct = share;
  b.on = true;
  return b;
}
struct Plan {
  std::vector<std::vector<BandTask>> band_tasks;  // [kBandClasses] tasks of the band kernel; read ranges index unique reads
  std::vector<uint64_t> band_pairs_by_rows;       // [K] band pairs whose haplotype has row class K (capacity of the
                                                  // stream-kernel tasks band_collect_kernel may append)
  uint64_t n_band_pairs = 0;                      // pairs sent to the band kernel (their cells are counted by the kernel)
  BandPolicy band;
  std::vector<std::vector<Task>> tasks;  // [K] -> tasks of class K (index 0 unused); read ranges index UNIQUE reads
  PodArray<uint32_t> hap_locus;          // [n_haps]
  std::vector<unsigned long long> ll_off;  // [n_loci+1] offsets of the caller-visible LL matrices (P_l x H_l)
  uint64_t n_pairs = 0, n_cells = 0;     // as the reference counts them (every pooled read x haplotype)
  uint64_t n_pairs_computed = 0, n_cells_computed = 0;  // after collapsing identical trimmed reads of a locus
  int max_n = 0, max_m = 0;
  std::vector<uint32_t> max_q;  // [K] longest (unique) read stream among the tasks of class K
  std::vector<uint8_t> multi_strip;  // [K] some task of the class needs more than one strip (scratch line hand-off)
  // Identical trimmed reads of a locus give identical log-likelihoods against every haplotype (the kernel is a
  // pure function of the two strings), so each distinct sequence is aligned once and the result fanned out.
  // LongTR pools reads by their +-200 bp sequence (ReadPooler, src/read_pooler.cpp:3-20) but aligns the +-5 bp
  // trim of it (HapAligner.cpp:346-465), so pools that differ only outside the trim window collapse here.
  std::vector<uint32_t> locus_uread_begin;  // [n_loci+1]
  PodArray<uint32_t> uread_off;             // [n_ureads+1]
  uint8_t* uread_bytes = nullptr;           // uread_nbytes bytes: caller-provided staging (pinned) or uread_owned
  size_t uread_nbytes = 0;
  std::unique_ptr<uint8_t[]> uread_owned;   // uninitialised on purpose: first touched by the parallel copy
  PodArray<uint32_t> read_to_uread;         // [n_reads] global unique-read index of every pooled read
  PodArray<uint32_t> read_locus;            // [n_reads]
  std::vector<unsigned long long> ull_off;  // [n_loci+1] offsets of the unique LL matrices (U_l x H_l)
};

inline uint64_t plan_mix(uint64_t a, uint64_t b) {  // 64x64 -> 128 bit multiply, folded
  const unsigned __int128 r = (unsigned __int128)a * b;
  return (uint64_t)r ^ (uint64_t)(r >> 64);
}
// 32 bytes per round on two independent multiply chains (equality is always confirmed with memcmp).
inline uint64_t plan_hash_bytes(const uint8_t* p, uint32_t n) {
  const uint64_t k0 = 0x9E3779B97F4A7C15ull, k1 = 0xD6E8FEB86659FD93ull, k2 = 0xFF51AFD7ED558CCDull, k3 = 0xC4CEB9FE1A85EC53ull;
  uint64_t s0 = k0 ^ n, s1 = k1;
  uint32_t i = 0;
  for (; i + 32 <= n; i += 32) {
    uint64_t w[4];
    std::memcpy(w, p + i, 32);
    s0 = plan_mix(w[0] ^ k2, w[1] ^ s0);
    s1 = plan_mix(w[2] ^ k3, w[3] ^ s1);
  }
  uint64_t w[4] = {0, 0, 0, 0};
  if (i < n) std::memcpy(w, p + i, n - i);
  s0 = plan_mix(w[0] ^ k2, w[1] ^ s0);
  s1 = plan_mix(w[2] ^ k3, w[3] ^ s1);
  return plan_mix(s0 ^ k1, s1 ^ k0);
}

template <typename F>
inline void plan_parallel_for(uint32_t n, int n_threads, F f) {  // f(begin, end, thread)
  if (n_threads <= 1 || n < 2) {
    f(0u, n, 0);
    return;
  }
  std::vector<std::thread> th;
  const uint32_t chunk = (n + (uint32_t)n_threads - 1) / (uint32_t)n_threads;
  for (int t = 0; t < n_threads; ++t) {
    const uint32_t b0 = std::min(n, (uint32_t)t * chunk), b1 = std::min(n, b0 + chunk);
    if (b0 < b1) th.emplace_back(f, b0, b1, t);
  }
  for (auto& x : th) x.join();
}

// Offsets monotone (nothing else in the batch is trusted before this holds): lets the C ABI start the uploads that do not
// depend on the plan while make_plan runs.
inline bool plan_offsets_valid(const ltr_viterbi_batch& b) {
  const uint32_t n_loci = b.n_loci;
  for (uint32_t l = 0; l < n_loci; ++l)
    if (b.locus_hap_begin[l + 1] < b.locus_hap_begin[l] || b.locus_read_begin[l + 1] < b.locus_read_begin[l]) return false;
  const uint32_t n_haps = b.locus_hap_begin[n_loci];
  for (uint32_t h = 0; h < n_haps; ++h)
    if (b.hap_off[h + 1] < b.hap_off[h]) return false;
  return true;
}

// Validates the batch, collapses duplicate reads per locus and builds per-class task lists (heaviest first within a
// class so the persistent warps finish together).  Returns LTR_OK or LTR_ERR_INVALID.
// stage(bytes, slot, user) may provide the buffers of the plan's large arrays (the C ABI hands out pinned host memory so
// that their uploads are asynchronous DMA instead of staged pageable copies); NULL = plain heap memory.
enum { PLAN_SLOT_UREAD_BYTES = 0, PLAN_SLOT_READ_TO_UREAD = 1, PLAN_SLOT_READ_LOCUS = 2, PLAN_SLOT_UREAD_OFF = 3, PLAN_SLOTS = 4 };
typedef uint8_t* (*PlanStageFn)(size_t bytes, int slot, void* user);
inline int make_plan(const ltr_viterbi_batch& b, const ltr_params& p, int kmax, Plan& out, int n_threads = 0,
                     PlanStageFn stage = nullptr, void* stage_user = nullptr, int band_w = -1) {
  const int cut = 35 - p.indel_flank_len;
  static const bool plan_timing = std::getenv("LTR_TIMING") != nullptr;  // diagnostics: phases on stderr
  auto tick = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!plan_timing) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[ltr] plan: %s %.1f ms\n", what, std::chrono::duration<double, std::milli>(now - tick).count());
    tick = now;
  };
  out.band = band_policy(p, band_w);
  out.band_tasks.assign(kBandClasses, std::vector<BandTask>());
  out.band_pairs_by_rows.assign(kmax + 1, 0);
  out.tasks.assign(kmax + 1, std::vector<Task>());
  out.max_q.assign(kmax + 1, 0);
  out.multi_strip.assign(kmax + 1, 0);
  const uint32_t n_loci = b.n_loci;
  const uint32_t n_haps = b.locus_hap_begin[n_loci], n_reads = b.locus_read_begin[n_loci];
  out.hap_locus.resize_uninit(n_haps);
  out.ll_off.assign((size_t)n_loci + 1, 0);
  out.ull_off.assign((size_t)n_loci + 1, 0);
  out.locus_uread_begin.assign((size_t)n_loci + 1, 0);
  out.read_to_uread.resize_uninit(n_reads, stage ? stage((size_t)n_reads * 4 + 16, PLAN_SLOT_READ_TO_UREAD, stage_user) : nullptr);
  out.read_locus.resize_uninit(n_reads, stage ? stage((size_t)n_reads * 4 + 16, PLAN_SLOT_READ_LOCUS, stage_user) : nullptr);
  for (uint32_t l = 0; l < n_loci; ++l)
    if (b.locus_hap_begin[l + 1] < b.locus_hap_begin[l] || b.locus_read_begin[l + 1] < b.locus_read_begin[l])
      return LTR_ERR_INVALID;
  if (n_threads <= 0) {
    // host threads of the plan: up to 16; LTR_PLAN_THREADS caps it (several ranks / batches in flight on one host)
    unsigned cap = 16u;
    if (const char* env = std::getenv("LTR_PLAN_THREADS")) cap = (unsigned)std::max(1, std::atoi(env));
    n_threads = (n_loci >= 4096) ? (int)std::min<unsigned>(cap, std::max(1u, std::thread::hardware_concurrency())) : 1;
  }
  {  // offsets must be monotone (reads non-empty); longest read
    std::vector<int> bad((size_t)std::max(1, n_threads), 0), mx((size_t)std::max(1, n_threads), 0);
    plan_parallel_for(n_reads, n_threads, [&](uint32_t r0, uint32_t r1, int t) {
      int m = 0, bd = 0;
      for (uint32_t r = r0; r < r1; ++r) {
        if (b.read_off[r + 1] <= b.read_off[r]) bd = 1;  // empty read
        else m = std::max<int>(m, (int)(b.read_off[r + 1] - b.read_off[r]));
      }
      bad[(size_t)t] = bd;
      mx[(size_t)t] = m;
    });
    for (size_t t = 0; t < bad.size(); ++t) {
      if (bad[t]) return LTR_ERR_INVALID;
      out.max_m = std::max(out.max_m, mx[t]);
    }
  }
  for (uint32_t h = 0; h < n_haps; ++h)
    if (b.hap_off[h + 1] < b.hap_off[h]) return LTR_ERR_INVALID;

  // ---- pass 1 (parallel over loci): local unique index of every read, unique count / bytes per locus ----
  // The distinct reads of a locus are numbered by increasing length (ties: first occurrence): the band kernel works on
  // rounds of consecutive pairs in lock step and runs of equal band class become one task.
  PodArray<uint32_t> local_u;  // bit 31: the read is the representative (first occurrence) of its sequence
  local_u.resize_uninit(n_reads);
  const uint32_t kRep = 0x80000000u;
  std::vector<uint32_t> ucount(n_loci, 0), ubytes(n_loci, 0);
  auto dedupe = [&](uint32_t l0, uint32_t l1, int) {
    std::vector<uint64_t> hashes;
    std::vector<uint32_t> reps;  // representative read of each unique sequence of the locus
    std::vector<uint32_t> order, rank;
    for (uint32_t l = l0; l < l1; ++l) {
      const uint32_t r0 = b.locus_read_begin[l], r1 = b.locus_read_begin[l + 1];
      hashes.clear();
      reps.clear();
      uint32_t bytes = 0;
      for (uint32_t r = r0; r < r1; ++r) {
        const uint8_t* s = b.read_bytes + b.read_off[r];
        const uint32_t len = b.read_off[r + 1] - b.read_off[r];
        const uint64_t h = plan_hash_bytes(s, len);
        uint32_t u = 0;
        for (; u < reps.size(); ++u) {
          if (hashes[u] != h) continue;
          const uint32_t q = reps[u];
          if (b.read_off[q + 1] - b.read_off[q] == len && std::memcmp(b.read_bytes + b.read_off[q], s, len) == 0) break;
        }
        if (u == reps.size()) {
          reps.push_back(r);
          hashes.push_back(h);
          bytes += len;
        }
        local_u[r] = u;
        out.read_locus[r] = l;
      }
      ucount[l] = (uint32_t)reps.size();
      ubytes[l] = bytes;
      order.resize(reps.size());
      rank.resize(reps.size());
      for (uint32_t u = 0; u < reps.size(); ++u) order[u] = u;
      std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
        return b.read_off[reps[x] + 1] - b.read_off[reps[x]] < b.read_off[reps[y] + 1] - b.read_off[reps[y]];
      });
      for (uint32_t k = 0; k < order.size(); ++k) rank[order[k]] = k;
      for (uint32_t r = r0; r < r1; ++r) local_u[r] = rank[local_u[r]];
      for (uint32_t u = 0; u < reps.size(); ++u) local_u[reps[u]] |= kRep;
    }
  };
  lap("validate");
  plan_parallel_for(n_loci, n_threads, dedupe);
  lap("dedupe");
  // ---- pass 2: prefix sums, unique read bytes ----------------------------------------------------------------
  std::vector<uint64_t> ubyte_off((size_t)n_loci + 1, 0);
  for (uint32_t l = 0; l < n_loci; ++l) {
    out.locus_uread_begin[l + 1] = out.locus_uread_begin[l] + ucount[l];
    ubyte_off[l + 1] = ubyte_off[l] + ubytes[l];
  }
  if (ubyte_off[n_loci] > 0xFFFFFFF0ull) return LTR_ERR_INVALID;
  const uint32_t n_ureads = out.locus_uread_begin[n_loci];
  out.uread_off.resize_uninit((size_t)n_ureads + 1,
                              stage ? stage(((size_t)n_ureads + 1) * 4 + 16, PLAN_SLOT_UREAD_OFF, stage_user) : nullptr);
  out.uread_nbytes = (size_t)ubyte_off[n_loci];
  out.uread_bytes = stage ? stage(out.uread_nbytes + 16, PLAN_SLOT_UREAD_BYTES, stage_user) : nullptr;
  if (out.uread_bytes == nullptr) {
    out.uread_owned.reset(new uint8_t[out.uread_nbytes + 16]);
    out.uread_bytes = out.uread_owned.get();
  }
  plan_parallel_for(n_loci, n_threads, [&](uint32_t l0, uint32_t l1, int) {
   for (uint32_t l = l0; l < l1; ++l) {
    const uint32_t r0 = b.locus_read_begin[l], r1 = b.locus_read_begin[l + 1], u0 = out.locus_uread_begin[l];
    const uint32_t nu = out.locus_uread_begin[l + 1] - u0;
    for (uint32_t r = r0; r < r1; ++r) {
      out.read_to_uread[r] = u0 + (local_u[r] & ~kRep);
      if (local_u[r] & kRep) out.uread_off[u0 + (local_u[r] & ~kRep)] = b.read_off[r + 1] - b.read_off[r];  // length, for now
    }
    uint32_t off = (uint32_t)ubyte_off[l];
    for (uint32_t u = 0; u < nu; ++u) {
      const uint32_t len = out.uread_off[u0 + u];
      out.uread_off[u0 + u] = off;
      off += len;
    }
    for (uint32_t r = r0; r < r1; ++r)
      if (local_u[r] & kRep)
        std::memcpy(out.uread_bytes + out.uread_off[u0 + (local_u[r] & ~kRep)], b.read_bytes + b.read_off[r],
                    b.read_off[r + 1] - b.read_off[r]);
   }
  });
  out.uread_off[n_ureads] = (uint32_t)ubyte_off[n_loci];
  lap("unique bytes");

  // ---- tasks -------------------------------------------------------------------------------------------------
  struct Key { uint64_t cost; Task t; int k; };
  for (uint32_t l = 0; l < n_loci; ++l) {
    const uint32_t H = b.locus_hap_begin[l + 1] - b.locus_hap_begin[l];
    out.ll_off[l + 1] = out.ll_off[l] + (unsigned long long)H * (b.locus_read_begin[l + 1] - b.locus_read_begin[l]);
    out.ull_off[l + 1] = out.ull_off[l] + (unsigned long long)H * (out.locus_uread_begin[l + 1] - out.locus_uread_begin[l]);
  }
  struct Part {
    std::vector<Key> keys;
    std::vector<std::vector<BandTask>> band;
    std::vector<uint64_t> band_by_rows;
    std::vector<uint32_t> max_q;
    std::vector<uint8_t> multi;
    uint64_t n_pairs = 0, n_cells = 0, n_pairs_c = 0, n_cells_c = 0, n_band_pairs = 0;
    int max_n = 0;
  };
  std::vector<Part> parts((size_t)std::max(1, n_threads));
  const BandPolicy& bp = out.band;
  plan_parallel_for(n_loci, n_threads, [&](uint32_t l0, uint32_t l1, int t) {
    Part& P = parts[(size_t)t];
    P.max_q.assign(kmax + 1, 0);
    P.multi.assign(kmax + 1, 0);
    P.band.assign(kBandClasses, std::vector<BandTask>());
    P.band_by_rows.assign(kmax + 1, 0);
    for (uint32_t l = l0; l < l1; ++l) {
      const uint32_t h0 = b.locus_hap_begin[l], h1 = b.locus_hap_begin[l + 1];
      const uint32_t r0 = b.locus_read_begin[l], r1 = b.locus_read_begin[l + 1];
      const uint32_t u0 = out.locus_uread_begin[l], u1 = out.locus_uread_begin[l + 1];
      uint64_t sum_m = 0;
      int min_m = 0x7FFFFFFF, max_m = 0;
      for (uint32_t r = r0; r < r1; ++r) {
        const int m = (int)(b.read_off[r + 1] - b.read_off[r]);
        sum_m += (uint64_t)m;
        min_m = std::min(min_m, m);
        max_m = std::max(max_m, m);
      }
      for (uint32_t h = h0; h < h1; ++h) {
        out.hap_locus[h] = l;
        const int hlen = (int)(b.hap_off[h + 1] - b.hap_off[h]);
        const int n = hlen - 2 * cut;
        if (r1 == r0) continue;
        P.n_pairs += r1 - r0;
        P.n_pairs_c += u1 - u0;
        const bool real = (hlen > 60 && n >= 1);
        int k = 1, strips = 1;
        if (real) {
          P.max_n = std::max(P.max_n, n);
          k = rows_per_lane(n, kmax);
          strips = std::max(1, (n - 1 + 32 * k - 1) / (32 * k));
          if (n - min_m <= 600 && max_m - n <= 600) {
            P.n_cells += (uint64_t)n * sum_m;
          } else {
            for (uint32_t r = r0; r < r1; ++r) {
              const int m = (int)(b.read_off[r + 1] - b.read_off[r]);
              if (std::abs(n - m) <= 600) P.n_cells += (uint64_t)n * (uint64_t)m;
            }
          }
        }
        // runs of consecutive unique reads (sorted by length) with the same band class; class -1 = stream kernel
        uint32_t run_begin = u0;
        int run_class = -2;
        auto close_run = [&](uint32_t run_end) {
          if (run_class == -2 || run_end == run_begin) return;
          if (run_class >= 0) {
            BandTask bt;
            bt.hap = h; bt.read_begin = run_begin; bt.read_end = run_end;
            P.band[(size_t)run_class].push_back(bt);
            P.band_by_rows[(size_t)k] += run_end - run_begin;
            P.n_band_pairs += run_end - run_begin;
            // uncertified pairs come back as stream-kernel tasks over sub-ranges of this run: size its scratch for them
            P.max_q[k] = std::max<uint32_t>(P.max_q[k], out.uread_off[run_end] - out.uread_off[run_begin]);
            if (strips > 1) P.multi[k] = 1;
          } else {
            const uint64_t q = (uint64_t)out.uread_off[run_end] - out.uread_off[run_begin];
            Key key;
            key.cost = real ? (uint64_t)k * strips * (q + 32) : (uint64_t)(run_end - run_begin);
            key.k = k;
            key.t.hap = h; key.t.read_begin = run_begin; key.t.read_end = run_end;
            P.keys.push_back(key);
            if (real) {
              P.max_q[k] = std::max<uint32_t>(P.max_q[k], (uint32_t)q);
              if (strips > 1) P.multi[k] = 1;
            }
          }
        };
        int last_m = -1, c = -1;
        for (uint32_t u = u0; u < u1; ++u) {
          const int m = (int)(out.uread_off[u + 1] - out.uread_off[u]);
          if (m != last_m) {  // reads are sorted by length: equal lengths are neighbours
            c = real ? band_class_of(hlen, n, m, bp) : -1;
            last_m = m;
          }
          // (the cells of banded pairs are counted by the kernel: band_cells of the pairs it evaluated)
          if (c < 0 && real && std::abs(n - m) <= 600) P.n_cells_c += (uint64_t)n * (uint64_t)m;
          if (c != run_class) {
            close_run(u);
            run_begin = u;
            run_class = c;
          }
        }
        close_run(u1);
      }
    }
  });
  lap("tasks");
  std::vector<Key> keys;
  keys.reserve(n_haps);
  for (const Part& P : parts) {
    keys.insert(keys.end(), P.keys.begin(), P.keys.end());
    out.n_pairs += P.n_pairs; out.n_cells += P.n_cells;
    out.n_pairs_computed += P.n_pairs_c; out.n_cells_computed += P.n_cells_c;
    out.n_band_pairs += P.n_band_pairs;
    if (!P.band.empty())
      for (int c = 0; c < kBandClasses; ++c)
        out.band_tasks[(size_t)c].insert(out.band_tasks[(size_t)c].end(), P.band[(size_t)c].begin(), P.band[(size_t)c].end());
    for (size_t k = 0; k < P.band_by_rows.size(); ++k) out.band_pairs_by_rows[k] += P.band_by_rows[k];
    out.max_n = std::max(out.max_n, P.max_n);
    for (size_t k = 0; k < P.max_q.size(); ++k) {
      out.max_q[k] = std::max(out.max_q[k], P.max_q[k]);
      out.multi_strip[k] |= P.multi[k];
    }
  }
  // Small batches (the per-locus entry points): a task streams ALL reads of its locus through one warp, which leaves
  // most of the GPU idle and makes the call latency the length of that stream.  Cut the read ranges so that there
  // are enough tasks to occupy the device; every piece still pays the 31-step pipeline fill once.
  const size_t kWantTasks = 2048;
  if (!keys.empty() && keys.size() < kWantTasks) {
    const uint32_t pieces = (uint32_t)((kWantTasks + keys.size() - 1) / keys.size());
    std::vector<Key> split;
    for (const Key& k : keys) {
      const uint32_t nr = k.t.read_end - k.t.read_begin;
      const uint32_t np = std::max(1u, std::min(pieces, nr));
      for (uint32_t c = 0; c < np; ++c) {
        Key part = k;
        part.t.read_begin = k.t.read_begin + (uint32_t)((uint64_t)nr * c / np);
        part.t.read_end = k.t.read_begin + (uint32_t)((uint64_t)nr * (c + 1) / np);
        part.cost = k.cost / np + 1;
        if (part.t.read_end > part.t.read_begin) split.push_back(part);
      }
    }
    keys.swap(split);
  }
  std::stable_sort(keys.begin(), keys.end(), [](const Key& a, const Key& c) { return a.cost > c.cost; });
  for (const Key& k : keys) out.tasks[k.k].push_back(k.t);
  lap("merge + sort");
  return LTR_OK;
}

// Fan the unique LL matrices (U_l x H_l) out to the caller-visible ones (P_l x H_l); host version for the CPU emulator
// (the product does this on the device, expand_ll_kernel).
inline void expand_ll_host(const ltr_viterbi_batch& b, const Plan& plan, const double* uniq_ll, double* out_ll) {
  const uint32_t n_reads = b.locus_read_begin[b.n_loci];
  for (uint32_t r = 0; r < n_reads; ++r) {
    const uint32_t l = plan.read_locus[r];
    const uint32_t H = b.locus_hap_begin[l + 1] - b.locus_hap_begin[l];
    const double* src = uniq_ll + plan.ull_off[l] + (size_t)(plan.read_to_uread[r] - plan.locus_uread_begin[l]) * H;
    double* dst = out_ll + plan.ll_off[l] + (size_t)(r - b.locus_read_begin[l]) * H;
    for (uint32_t h = 0; h < H; ++h) dst[h] = src[h];
  }
}

}  // namespace ltr
