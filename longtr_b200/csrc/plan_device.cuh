// plan_device.cuh -- the plan of a job built on the device (what make_plan, viterbi_host.h, builds on the host).
//
// A job's raw pooled reads travel to the device as they are; the kernels of plan_kernels.cu then
//   1. collapse the identical trimmed reads of every locus (the Viterbi score is a pure function of the two strings;
//      LongTR pools reads by their +-200 bp sequence, src/read_pooler.cpp:3-20, but aligns the +-5 bp trim of it,
//      HapAligner.cpp:346-465) and number the distinct ones by increasing length (ties: first occurrence),
//   2. scan the per-locus counts into offsets, compact the distinct reads into one byte stream per locus,
//   3. cut every (haplotype, distinct reads of its locus) row into runs of equal band class and emit them as tasks of
//      the banded kernel (pair lists per band class) or of the full-matrix stream kernel (lists per row class, heaviest
//      cost bucket first),
// with no host involvement: nothing here needs a device-to-host copy, so a job is one stream-ordered sequence of
// copies and kernels (ltr_job_submit / ltr_job_wait).  Numbering, offsets and task sets are identical to make_plan's
// (tests/test_emulator.py::test_device_plan_matches_host_plan holds them against each other through the CPU build of
// this header); only the order of the tasks inside a list differs (cost buckets instead of a full sort).
//
// The per-item functions are written for `nl` cooperating lanes (a warp on the device, one "lane" on the host).
#pragma once
#include <stdint.h>

#include "band_core.cuh"
#include "viterbi_core.cuh"

namespace ltr {

static constexpr uint32_t kPlanRep = 0x80000000u;  // local_u bit 31: first occurrence of its sequence in the locus
static constexpr int kPlanMaxK = 16;               // row classes 1..16 (viterbi_max_rows_per_lane)
static constexpr int kPlanBuckets = 32;            // cost buckets floor(log2 cost) per row class
static constexpr int kPlanSlots = (kPlanMaxK + 1) * kPlanBuckets;

enum {  // words of PlanDev::ctl
  PLAN_CTL_ERR = 0,       // != 0: malformed batch (bit 0: read offsets, bit 1: more than 4 GB of distinct read bytes)
  PLAN_CTL_N_UREADS = 1,  // distinct reads of the job
  PLAN_CTL_N_BAND_TASKS = 2,
  PLAN_CTL_N_BAND_PAIRS = 3,
  PLAN_CTL_BAND_INFO = 4,                                  // [kBandClasses][2]: first pair, number of pairs
  PLAN_CTL_BAND_TASK_COUNT = PLAN_CTL_BAND_INFO + 2 * kBandClasses,     // [kBandClasses] tasks per band class
  PLAN_CTL_BAND_TASK_BASE = PLAN_CTL_BAND_TASK_COUNT + kBandClasses,    // [kBandClasses] first task of the class
  PLAN_CTL_ST_COUNT = PLAN_CTL_BAND_TASK_BASE + kBandClasses,           // [kPlanSlots] pass 0: stream tasks per (row class, bucket)
  PLAN_CTL_ST_BASE = PLAN_CTL_ST_COUNT + kPlanSlots,
  PLAN_CTL_ST_FILL = PLAN_CTL_ST_BASE + kPlanSlots,
  PLAN_CTL_WORDS = PLAN_CTL_ST_FILL + kPlanSlots
};
enum {  // 64-bit statistics, PlanDev::stat
  PLAN_STAT_CELLS = 0,           // reference-defined cells: n*m over every pooled read x haplotype with |n-m| <= 600
  PLAN_STAT_CELLS_STREAM = 1,    // n*m of the distinct pairs planned for the stream kernel
  PLAN_STAT_PAIRS_COMPUTED = 2,  // distinct pairs
  PLAN_STAT_MAX_M = 3,           // longest read
  PLAN_STAT_WORDS = 4
};

struct PlanPair {  // same layout as uint2: (haplotype, distinct read)
  uint32_t x, y;
};

struct PlanDev {
  uint32_t n_loci, n_haps, n_reads;
  uint32_t raw_total;  // bytes of raw reads on the device (= read_off[n_reads] as the host saw it)
  int32_t cut;         // 35 - INDEL_FLANK_LEN
  int32_t kmax;
  BandPolicy band;
  // inputs (device)
  const uint32_t* lhb;        // [n_loci+1]
  const uint32_t* lrb;        // [n_loci+1] raw pooled reads
  const uint32_t* hap_off;    // [n_haps+1]
  const uint32_t* read_off;   // [n_reads+1] raw
  const uint8_t* read_bytes;  // raw, padded on both sides
  const uint8_t* packed;      // NULL, or the raw reads as one 4-bit stream (base b in byte b / 2, even b in the high nibble):
                              // unpacked into read_bytes by the first kernel of the plan
  // scratch
  unsigned long long* rhash;  // [n_reads]
  uint32_t* rlen;             // [n_reads] length of every raw read (0: malformed offsets)
  uint32_t* rep;              // [n_reads] first read of the locus with the same sequence
  uint32_t* rank_of;          // [n_reads] (valid at representatives) rank by (length, first occurrence)
  uint32_t* tmp_len;          // [n_reads] at lrb[l] + rank: length of the distinct read
  uint32_t* tmp_rep;          // [n_reads] at lrb[l] + rank: its representative raw read
  uint32_t* ucount;           // [n_loci]
  uint32_t* ubytes;           // [n_loci]
  uint32_t* ubyte_off;        // [n_loci+1]
  uint32_t* band_task_pos;    // [kBandClasses][n_loci] pass 0: band tasks of (class, locus); after the scan: list position
  uint32_t* band_pair_pos;    // [kBandClasses][n_loci] the same for their pairs
  unsigned long long* scan_partial;  // tile sums of the multi-block scans: 3 * ceil(kBandClasses * n_loci / 1024) + 8 words
  // products
  uint32_t* local_u;          // [n_reads] rank | kPlanRep
  uint32_t* read_locus;       // [n_reads]
  uint32_t* hap_locus;        // [n_haps]
  uint32_t* lub;              // [n_loci+1] distinct reads per locus, prefix sums
  unsigned long long* ull_off;   // [n_loci+1] offsets of the distinct LL matrices (U_l x H_l)
  uint32_t* uread_off;        // [n_reads+1] (n_ureads+1 used)
  uint8_t* uread_bytes;       // compacted distinct reads, padded on both sides
  uint32_t* r2u;              // [n_reads]
  uint32_t* ctl;              // [PLAN_CTL_WORDS]
  unsigned long long* stat;   // [PLAN_STAT_WORDS]
  // task lists
  BandTask* band_tasks;       // all band classes back to back (capacity band_cap tasks)
  PlanPair* band_pairs;       // their pairs, class after class (capacity band_cap)
  uint32_t band_cap;
  Task* st_tasks[kPlanMaxK + 1];   // stream-kernel task list of each row class (NULL: no haplotype of that class)
  uint32_t st_cap[kPlanMaxK + 1];
  uint32_t* st_ntasks[kPlanMaxK + 1];  // where the class keeps its task count (control word 1 of the stream kernel)
};

// ---- small helpers ----------------------------------------------------------------------------------------------
LTR_HD void plan_sync() {
#ifdef LTR_DEVICE_CODE
  __syncwarp();
#endif
}
LTR_HD uint32_t plan_warp_sum(uint32_t v) {
#ifdef LTR_DEVICE_CODE
  return __reduce_add_sync(0xFFFFFFFFu, v);
#else
  return v;
#endif
}
LTR_HD uint32_t plan_atomic_add(uint32_t* p, uint32_t v) {
#ifdef LTR_DEVICE_CODE
  return atomicAdd(p, v);
#else
  const uint32_t o = *p;
  *p = o + v;
  return o;
#endif
}
LTR_HD void plan_atomic_or(uint32_t* p, uint32_t v) {
#ifdef LTR_DEVICE_CODE
  atomicOr(p, v);
#else
  *p |= v;
#endif
}
LTR_HD void plan_atomic_add64(unsigned long long* p, unsigned long long v) {
#ifdef LTR_DEVICE_CODE
  atomicAdd(p, v);
#else
  *p += v;
#endif
}
LTR_HD void plan_atomic_max64(unsigned long long* p, unsigned long long v) {
#ifdef LTR_DEVICE_CODE
  atomicMax(p, v);
#else
  if (*p < v) *p = v;
#endif
}
// Four bytes at any address (the string buffers are padded, reading past the end of a read is harmless).
LTR_HD uint32_t plan_load32(const uint8_t* p) {
#ifdef LTR_DEVICE_CODE
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
  return __funnelshift_r(w[0], w[1], (uint32_t)(a & 3u) * 8u);
#else
  uint32_t v;
  std::memcpy(&v, p, 4);
  return v;
#endif
}
LTR_HD unsigned long long plan_hash(const uint8_t* s, uint32_t len) {  // filter only: equality is confirmed byte by byte
  unsigned long long h0 = 0x9E3779B97F4A7C15ull ^ len, h1 = 0xD6E8FEB86659FD93ull;
  uint32_t i = 0;
  for (; i + 8 <= len; i += 8) {
    h0 = (h0 ^ plan_load32(s + i)) * 0xFF51AFD7ED558CCDull;
    h1 = (h1 ^ plan_load32(s + i + 4)) * 0xC4CEB9FE1A85EC53ull;
  }
  for (; i < len; ++i) h0 = (h0 ^ s[i]) * 0x100000001B3ull;
  h0 ^= h0 >> 29;
  return (h0 + h1) * 0x9E3779B97F4A7C15ull;
}
LTR_HD bool plan_equal(const uint8_t* a, const uint8_t* b, uint32_t len) {
  uint32_t i = 0;
  for (; i + 4 <= len; i += 4)
    if (plan_load32(a + i) != plan_load32(b + i)) return false;
  for (; i < len; ++i)
    if (a[i] != b[i]) return false;
  return true;
}

// ---- step 1: one locus, nl lanes: distinct reads, their order, per-locus counts ----------------------------------------
// bytes / origin: where the raw read bytes are read from -- byte at raw offset o is bytes[o - origin].  The device kernel
// passes a shared-memory copy of the locus' bytes when it fits (coalesced staging instead of 32 lanes streaming 32
// different reads through L1), otherwise (and on the host) the raw buffer itself with origin 0.
// limit: no read of the locus may end behind it (end of the staged span, or of the upload).
LTR_HD void plan_locus_dedupe(const PlanDev& P, uint32_t l, uint32_t lane, uint32_t nl, const uint8_t* bytes, uint32_t origin,
                              uint32_t limit) {
  const uint32_t r0 = P.lrb[l], r1 = P.lrb[l + 1];
  for (uint32_t h = P.lhb[l] + lane; h < P.lhb[l + 1]; h += nl) P.hap_locus[h] = l;
  uint32_t bad = 0, max_m = 0;
  for (uint32_t r = r0 + lane; r < r1; r += nl) {
    const uint32_t o0 = P.read_off[r], o1 = P.read_off[r + 1];
    const bool ok = (o1 > o0) && (o1 <= limit) && (o0 >= origin);  // not empty, inside the upload / the staged span
    const uint32_t len = ok ? o1 - o0 : 0u;
    bad |= ok ? 0u : 1u;
    max_m = len > max_m ? len : max_m;
    P.read_locus[r] = l;
    P.rlen[r] = len;
    P.rhash[r] = plan_hash(bytes + ((ok ? o0 : origin) - origin), len);
  }
  if (bad) plan_atomic_or(P.ctl + PLAN_CTL_ERR, 1u);
  if (max_m) plan_atomic_max64(P.stat + PLAN_STAT_MAX_M, (unsigned long long)max_m);
  plan_sync();
  // first occurrence of every sequence
  for (uint32_t r = r0 + lane; r < r1; r += nl) {
    const unsigned long long h = P.rhash[r];
    const uint32_t len = P.rlen[r];
    const uint8_t* s = bytes + (P.read_off[r] - origin);  // (len == 0: never dereferenced)
    uint32_t rep = r;
    for (uint32_t q = r0; q < r; ++q) {
      if (P.rhash[q] != h || P.rlen[q] != len) continue;
      if (len == 0 || plan_equal(bytes + (P.read_off[q] - origin), s, len)) {
        rep = q;
        break;
      }
    }
    P.rep[r] = rep;
  }
  plan_sync();
  // rank of every representative by (length, first occurrence)
  uint32_t n_rep = 0, n_bytes = 0;
  for (uint32_t r = r0 + lane; r < r1; r += nl) {
    if (P.rep[r] != r) continue;
    const uint32_t len = P.rlen[r];
    uint32_t rank = 0;
    for (uint32_t q = r0; q < r1; ++q) {
      if (P.rep[q] != q) continue;
      const uint32_t lq = P.rlen[q];
      rank += (lq < len || (lq == len && q < r)) ? 1u : 0u;
    }
    P.rank_of[r] = rank;
    ++n_rep;
    n_bytes += len;
  }
  n_rep = plan_warp_sum(n_rep);
  n_bytes = plan_warp_sum(n_bytes);
  if (lane == 0) {
    P.ucount[l] = n_rep;
    P.ubytes[l] = n_bytes;
  }
  plan_sync();
  for (uint32_t r = r0 + lane; r < r1; r += nl) {
    const uint32_t q = P.rep[r];
    const uint32_t k = P.rank_of[q];
    P.local_u[r] = k | (q == r ? kPlanRep : 0u);
    if (q == r) {  // lengths / representatives by rank
      P.tmp_len[r0 + k] = P.rlen[r];
      P.tmp_rep[r0 + k] = r;
    }
  }
}

// ---- step 2: exclusive scans over the loci (host form; the device kernel is a block scan of the same sums) ------------
LTR_HHD void plan_scan_serial(const PlanDev& P) {
  const bool err = P.ctl[PLAN_CTL_ERR] != 0;
  uint32_t nu = 0;
  unsigned long long nb = 0, nll = 0;
  for (uint32_t l = 0; l < P.n_loci; ++l) {
    P.lub[l] = nu;
    P.ubyte_off[l] = (uint32_t)nb;
    P.ull_off[l] = nll;
    const uint32_t c = err ? 0u : P.ucount[l];
    nu += c;
    nb += err ? 0u : P.ubytes[l];
    nll += (unsigned long long)c * (P.lhb[l + 1] - P.lhb[l]);
  }
  P.lub[P.n_loci] = nu;
  P.ubyte_off[P.n_loci] = (uint32_t)nb;
  P.ull_off[P.n_loci] = nll;
  P.ctl[PLAN_CTL_N_UREADS] = nu;
  P.stat[PLAN_STAT_PAIRS_COMPUTED] = nll;
  if (nb > 0xFFFFFFF0ull) P.ctl[PLAN_CTL_ERR] |= 2u;
}

// ---- step 3: one locus, nl lanes: offsets and bytes of its distinct reads, read -> distinct read map ------------------
LTR_HD void plan_locus_fill(const PlanDev& P, uint32_t l, uint32_t lane, uint32_t nl) {
  if (P.ctl[PLAN_CTL_ERR] != 0) return;
  const uint32_t r0 = P.lrb[l], r1 = P.lrb[l + 1];
  const uint32_t u0 = P.lub[l], nu = P.lub[l + 1] - u0;
#ifdef LTR_DEVICE_CODE
  if (nu <= nl) {  // one distinct read per lane: offsets by a warp scan of the lengths
    const uint32_t len = lane < nu ? P.tmp_len[r0 + lane] : 0u;
    uint32_t inc = len;
    for (uint32_t d = 1; d < 32u; d <<= 1) {
      const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, inc, d);
      if (lane >= d) inc += up;
    }
    const uint32_t base = P.ubyte_off[l];
    if (lane < nu) P.uread_off[u0 + lane] = base + inc - len;
    if (l + 1 == P.n_loci && lane == (nu ? nu - 1 : 0)) P.uread_off[u0 + nu] = base + (nu ? inc : 0u);
  } else
#endif
  if (lane == 0) {
    uint32_t off = P.ubyte_off[l];
    for (uint32_t k = 0; k < nu; ++k) {
      P.uread_off[u0 + k] = off;
      off += P.tmp_len[r0 + k];
    }
    if (l + 1 == P.n_loci) P.uread_off[u0 + nu] = off;
  }
  plan_sync();
  for (uint32_t k = 0; k < nu; ++k) {
    const uint8_t* src = P.read_bytes + P.read_off[P.tmp_rep[r0 + k]];
    uint8_t* dst = P.uread_bytes + P.uread_off[u0 + k];
    const uint32_t len = P.tmp_len[r0 + k];
    for (uint32_t i = lane; i < len; i += nl) dst[i] = src[i];
  }
  for (uint32_t r = r0 + lane; r < r1; r += nl) P.r2u[r] = u0 + (P.local_u[r] & ~kPlanRep);
}

// ---- step 4: one locus: every haplotype's runs of distinct reads with the same band class -> tasks ---------------------
// pass 0 counts (per band class and locus: tasks, pairs; per row class and cost bucket: tasks) and accumulates the
// statistics; the scans turn the counts into list positions; pass 1 writes tasks and pairs.  The band lists come out
// class by class in locus order, haplotype by haplotype, reads by length -- the order make_plan produces: the pairs
// of a round of the band kernel (32 / G consecutive pairs walked in lock step) then belong to one locus and have
// similar lengths.  The stream lists are ordered by cost bucket only (a warp takes one task at a time).
LTR_HD void plan_locus_tasks(const PlanDev& P, uint32_t l, int pass) {
  if (P.ctl[PLAN_CTL_ERR] != 0) return;
  const uint32_t r0 = P.lrb[l], r1 = P.lrb[l + 1];
  if (r1 == r0) return;
  const uint32_t u0 = P.lub[l], u1 = P.lub[l + 1];
  unsigned long long cells_ref = 0, cells_stream = 0;
  for (uint32_t g = P.lhb[l]; g < P.lhb[l + 1]; ++g) {
    const int32_t hlen = (int32_t)(P.hap_off[g + 1] - P.hap_off[g]);
    const int32_t n = hlen - 2 * P.cut;
    const bool real = (hlen > 60 && n >= 1);
    int k = 1, strips = 1;
    if (real) {
      k = rows_per_lane_hd(n, P.kmax);
      strips = (n - 1 + 32 * k - 1) / (32 * k);
      if (strips < 1) strips = 1;
    }
    if (pass == 0 && real) {
      for (uint32_t r = r0; r < r1; ++r) {
        const int32_t m = (int32_t)(P.read_off[r + 1] - P.read_off[r]);
        const int32_t d = n - m;
        if ((d < 0 ? -d : d) <= 600) cells_ref += (unsigned long long)n * (unsigned long long)m;
      }
    }
    uint32_t run_begin = u0;
    int run_class = -2;
    int last_m = -1, c = -1;
    for (uint32_t u = u0; u <= u1; ++u) {
      if (u < u1) {
        const int m = (int32_t)(P.uread_off[u + 1] - P.uread_off[u]);
        if (m != last_m) {  // distinct reads are sorted by length: equal lengths are neighbours
          c = real ? band_class_of(hlen, n, m, P.band) : -1;
          last_m = m;
        }
        if (pass == 0 && c < 0 && real) {
          const int32_t d = n - m;
          if ((d < 0 ? -d : d) <= 600) cells_stream += (unsigned long long)n * (unsigned long long)m;
        }
      }
      const int cu = (u < u1) ? c : -3;  // sentinel closes the last run
      if (cu == run_class) continue;
      if (run_class != -2 && u > run_begin) {
        const uint32_t len = u - run_begin;
        if (run_class >= 0) {
          uint32_t* tp = P.band_task_pos + (size_t)run_class * P.n_loci + l;
          uint32_t* pp = P.band_pair_pos + (size_t)run_class * P.n_loci + l;
          if (pass == 0) {
            *tp += 1u;  // the entries of (class, locus) belong to this thread alone
            *pp += len;
          } else {
            const uint32_t t = *tp, pb = *pp;
            *tp = t + 1u;
            *pp = pb + len;
            if (t < P.band_cap && pb + len <= P.band_cap) {
              BandTask bt;
              bt.hap = g;
              bt.read_begin = run_begin;
              bt.read_end = u;
              P.band_tasks[t] = bt;
              for (uint32_t i = 0; i < len; ++i) {
                PlanPair pr;
                pr.x = g;
                pr.y = run_begin + i;
                P.band_pairs[pb + i] = pr;
              }
            }
          }
        } else {
          // cost model of make_plan: rows per lane x strips x (stream length + pipeline fill)
          const unsigned long long q = (unsigned long long)(P.uread_off[u] - P.uread_off[run_begin]);
          const unsigned long long cost = real ? (unsigned long long)k * (unsigned long long)strips * (q + 32ull) : (unsigned long long)len;
          int bucket = 0;
          for (unsigned long long v = cost; v > 1ull; v >>= 1) ++bucket;
          bucket = bucket > kPlanBuckets - 1 ? kPlanBuckets - 1 : bucket;
          const int slot = k * kPlanBuckets + bucket;
          if (pass == 0) {
            plan_atomic_add(P.ctl + PLAN_CTL_ST_COUNT + slot, 1u);
          } else {
            const uint32_t idx = P.ctl[PLAN_CTL_ST_BASE + slot] + plan_atomic_add(P.ctl + PLAN_CTL_ST_FILL + slot, 1u);
            if (P.st_tasks[k] && idx < P.st_cap[k]) {
              Task T;
              T.hap = g;
              T.read_begin = run_begin;
              T.read_end = u;
              P.st_tasks[k][idx] = T;
            }
          }
        }
      }
      run_begin = u;
      run_class = cu;
    }
  }
  if (pass == 0) {
    if (cells_ref) plan_atomic_add64(P.stat + PLAN_STAT_CELLS, cells_ref);
    if (cells_stream) plan_atomic_add64(P.stat + PLAN_STAT_CELLS_STREAM, cells_stream);
  }
}

// ---- step 5: list positions from the counts of pass 0 ---------------------------------------------------------------
// (a) exclusive scan of the [class][locus] band counters (host form; the device kernel is a block scan of the same sums)
LTR_HHD void plan_band_scan_serial(const PlanDev& P) {
  uint32_t nt = 0, np = 0;
  for (int c = 0; c < kBandClasses; ++c) {
    P.ctl[PLAN_CTL_BAND_TASK_BASE + c] = nt;
    P.ctl[PLAN_CTL_BAND_INFO + 2 * c] = np;
    for (uint32_t l = 0; l < P.n_loci; ++l) {
      uint32_t* tp = P.band_task_pos + (size_t)c * P.n_loci + l;
      uint32_t* pp = P.band_pair_pos + (size_t)c * P.n_loci + l;
      const uint32_t a = *tp, b = *pp;
      *tp = nt;
      *pp = np;
      nt += a;
      np += b;
    }
    P.ctl[PLAN_CTL_BAND_TASK_COUNT + c] = nt - P.ctl[PLAN_CTL_BAND_TASK_BASE + c];
    P.ctl[PLAN_CTL_BAND_INFO + 2 * c + 1] = np - P.ctl[PLAN_CTL_BAND_INFO + 2 * c];
  }
  P.ctl[PLAN_CTL_N_BAND_TASKS] = nt < P.band_cap ? nt : P.band_cap;
  P.ctl[PLAN_CTL_N_BAND_PAIRS] = np;
}
// (b) stream lists: positions of the cost buckets, heaviest first (one thread)
LTR_HHD void plan_task_scan(const PlanDev& P) {
  for (int k = 1; k <= P.kmax && k <= kPlanMaxK; ++k) {
    uint32_t running = 0;
    for (int b = kPlanBuckets - 1; b >= 0; --b) {
      P.ctl[PLAN_CTL_ST_BASE + k * kPlanBuckets + b] = running;
      running += P.ctl[PLAN_CTL_ST_COUNT + k * kPlanBuckets + b];
    }
    if (P.st_ntasks[k]) *P.st_ntasks[k] = running < P.st_cap[k] ? running : P.st_cap[k];
  }
}

}  // namespace ltr
