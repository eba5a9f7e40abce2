// viterbi_host.h -- host-side preparation shared by the C-ABI implementation and by the
// CPU lane emulator of the unit tests: launch constants, boundary tables, task planning.
#pragma once
#include <stdint.h>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "longtr_b200.h"
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
inline int rows_per_lane(int n, int kmax) {
  const int R = n - 1;
  if (R <= 32) return 1;
  const int strips = (R + 32 * kmax - 1) / (32 * kmax);
  const int per = (R + strips - 1) / strips;
  return std::max(1, (per + 31) / 32);
}

struct Plan {
  std::vector<std::vector<Task>> tasks;  // [K] -> tasks of class K (index 0 unused)
  std::vector<uint32_t> hap_locus;       // [n_haps]
  std::vector<unsigned long long> ll_off;  // [n_loci+1]
  uint64_t n_pairs = 0, n_cells = 0;
  int max_n = 0, max_m = 0;
  std::vector<uint32_t> max_q;  // [K] longest read stream among the tasks of class K
};

// Validates the batch and builds per-class task lists (heaviest first within a class so the
// persistent warps finish together).  Returns LTR_OK or LTR_ERR_INVALID.
inline int make_plan(const ltr_viterbi_batch& b, const ltr_params& p, int kmax, Plan& out) {
  const int cut = 35 - p.indel_flank_len;
  out.tasks.assign(kmax + 1, std::vector<Task>());
  out.max_q.assign(kmax + 1, 0);
  const uint32_t n_haps = b.locus_hap_begin[b.n_loci], n_reads = b.locus_read_begin[b.n_loci];
  out.hap_locus.assign(n_haps, 0);
  out.ll_off.assign((size_t)b.n_loci + 1, 0);
  for (uint32_t r = 0; r < n_reads; ++r) {
    if (b.read_off[r + 1] <= b.read_off[r]) return LTR_ERR_INVALID;  // empty read
    out.max_m = std::max<int>(out.max_m, (int)(b.read_off[r + 1] - b.read_off[r]));
  }
  struct Key { uint64_t cost; Task t; int k; };
  std::vector<Key> keys;
  keys.reserve(n_haps);
  for (uint32_t l = 0; l < b.n_loci; ++l) {
    const uint32_t h0 = b.locus_hap_begin[l], h1 = b.locus_hap_begin[l + 1];
    const uint32_t r0 = b.locus_read_begin[l], r1 = b.locus_read_begin[l + 1];
    if (h1 < h0 || r1 < r0) return LTR_ERR_INVALID;
    out.ll_off[l + 1] = out.ll_off[l] + (unsigned long long)(h1 - h0) * (r1 - r0);
    const uint64_t q = (uint64_t)b.read_off[r1] - b.read_off[r0];
    for (uint32_t h = h0; h < h1; ++h) {
      out.hap_locus[h] = l;
      if (b.hap_off[h + 1] < b.hap_off[h]) return LTR_ERR_INVALID;
      const int hlen = (int)(b.hap_off[h + 1] - b.hap_off[h]);
      const int n = hlen - 2 * cut;
      if (r1 == r0) continue;
      int k = 1;
      uint64_t cost = r1 - r0;
      if (hlen > 60 && n >= 1) {
        out.max_n = std::max(out.max_n, n);
        k = rows_per_lane(n, kmax);
        const int strips = std::max(1, (n - 1 + 32 * k - 1) / (32 * k));
        cost = (uint64_t)k * strips * (q + 32);
        out.max_q[k] = std::max<uint32_t>(out.max_q[k], (uint32_t)q);
        for (uint32_t r = r0; r < r1; ++r) {
          const int m = (int)(b.read_off[r + 1] - b.read_off[r]);
          if (std::abs(n - m) <= 600) out.n_cells += (uint64_t)n * (uint64_t)m;
        }
      }
      out.n_pairs += r1 - r0;
      Key key;
      key.cost = cost; key.k = k;
      key.t.hap = h; key.t.read_begin = r0; key.t.read_end = r1;
      keys.push_back(key);
    }
  }
  std::stable_sort(keys.begin(), keys.end(), [](const Key& a, const Key& c) { return a.cost > c.cost; });
  for (const Key& k : keys) out.tasks[k.k].push_back(k.t);
  return LTR_OK;
}

}  // namespace ltr
