// locus_batcher.cpp -- ltr_genotyper_*: many raw loci -> genotype calls (include/longtr_b200.h).
//
// The front half of SeqStutterGenotyper::genotype (reference src/seq_stutter_genotyper.cpp:485-497, 599-645) for a batch
// of loci in structure-of-arrays form.  Per chunk of loci:
//   prepare (host threads)  pool the reads of every locus by sequence (ReadPooler, src/read_pooler.cpp:3-20), trim every pool
//                           with its CIGAR (HapAligner::trim_alignment, HapAligner.cpp:346-465, run-wise instead of base by
//                           base), write haplotypes (left flank + allele + right flank: Haplotype::get_seq) and trimmed
//                           pools straight into the pinned arrays of an ltr_viterbi_batch / ltr_posterior_batch;
//   submit (one call)       ltr_job_submit_outputs: upload, device plan, Viterbi kernels, pool -> read scatter, posteriors,
//                           removal of uncalled alleles + second pass, download -- asynchronous;
//   finish (host threads)   Genotyper::extract_genotypes_and_likelihoods per locus on the surviving alleles.
// Chunks are sharded over the devices round robin and two chunks per device are in flight, so the host prepares chunk k+1
// while the GPUs work on chunk k.  Results are written by input locus index: the output order is the input (BED) order.
// No per-read std::string / Alignment objects anywhere; no CPU fallback (every chunk goes through the C ABI to a GPU).
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "longtr_b200.h"
#include "longtr_host.h"

namespace {

using Clock = std::chrono::steady_clock;
inline double ms_since(Clock::time_point t0) { return std::chrono::duration<double, std::milli>(Clock::now() - t0).count(); }

// ---- persistent worker pool: parallel_for over [0, n) in contiguous blocks -----------------------------------------------
class WorkerPool {
 public:
  explicit WorkerPool(int n) : n_threads_(std::max(1, n)) {
    for (int t = 1; t < n_threads_; ++t) threads_.emplace_back([this, t] { loop(t); });
  }
  ~WorkerPool() {
    {
      std::lock_guard<std::mutex> g(m_);
      stop_ = true;
      ++epoch_;
    }
    cv_.notify_all();
    for (std::thread& t : threads_) t.join();
  }
  // f(begin, end, thread); blocks of `grain` items are handed out dynamically (loci differ a lot in cost)
  void parallel_for(uint32_t n, uint32_t grain, const std::function<void(uint32_t, uint32_t, int)>& f) {
    if (n == 0) return;
    if (n_threads_ == 1 || n <= grain) {
      f(0, n, 0);
      return;
    }
    {
      std::lock_guard<std::mutex> g(m_);
      fn_ = &f;
      n_ = n;
      grain_ = std::max(1u, grain);
      next_.store(0);
      pending_ = n_threads_ - 1;
      ++epoch_;
    }
    cv_.notify_all();
    work(0);
    std::unique_lock<std::mutex> g(m_);
    done_cv_.wait(g, [this] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  void work(int t) {
    for (;;) {
      const uint32_t b = next_.fetch_add(grain_);
      if (b >= n_) break;
      (*fn_)(b, std::min(n_, b + grain_), t);
    }
  }
  void loop(int t) {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [&] { return epoch_ != seen; });
        seen = epoch_;
        if (stop_) return;
      }
      work(t);
      {
        std::lock_guard<std::mutex> g(m_);
        if (--pending_ == 0) done_cv_.notify_one();
      }
    }
  }
  int n_threads_;
  std::vector<std::thread> threads_;
  std::mutex m_;
  std::condition_variable cv_, done_cv_;
  const std::function<void(uint32_t, uint32_t, int)>* fn_ = nullptr;
  uint32_t n_ = 0, grain_ = 1;
  std::atomic<uint32_t> next_{0};
  int pending_ = 0;
  uint64_t epoch_ = 0;
  bool stop_ = false;
};

// ---- pinned, grow-only host array ----------------------------------------------------------------------------------------
template <typename T>
struct Pinned {
  T* p = nullptr;
  size_t cap = 0;
  bool reserve(size_t n) {
    if (n <= cap) return true;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    const size_t want = n + n / 4 + 64;
    if (cudaHostAlloc(reinterpret_cast<void**>(&p), want * sizeof(T), cudaHostAllocDefault) != cudaSuccess) {
      cudaGetLastError();
      p = nullptr;
      return false;
    }
    cap = want;
    return true;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

// ---- HapAligner::trim_alignment on a BAM-encoded CIGAR, run-wise ------------------------------------------------------------
// The reference (HapAligner.cpp:346-465) pops the CIGAR one base at a time from the front while the read position is left
// of the window, handles the pad (a deletion inside the pad pulls one upstream base back in, insertions stay), and does
// the same from the back.  The same decisions are taken here on whole runs.  Returns false where the reference dies
// (unknown operation) or asserts (trims exceed the read).
enum OpClass { OP_ALIGNED, OP_DEL, OP_READ_ONLY, OP_NONE, OP_BAD };
inline OpClass classify_bam_op(uint32_t op) {
  switch (op & 15u) {
    case 0: case 7: case 8: return OP_ALIGNED;  // M = X
    case 2: return OP_DEL;                      // D
    case 1: case 4: return OP_READ_ONLY;        // I S
    case 5: return OP_NONE;                     // H
    default: return OP_BAD;                     // N P and anything else: the reference's switch prints an error and dies
  }
}

bool trim_packed(const uint32_t* ops, uint32_t n_ops, int32_t aln_start, int32_t aln_stop, int32_t seq_len,
                 int32_t repeat_start, int32_t repeat_end, int32_t pad, int32_t& out_ltrim, int32_t& out_len) {
  const int32_t lo = repeat_start - pad, hi = repeat_end + pad;
  int32_t start_pos = aln_start + 1, end_pos = aln_stop + 1;
  int64_t ltrim = 0, rtrim = 0;
  // live elements are [f, b); fn / bn = bases already consumed of the front / back element
  uint32_t f = 0, b = n_ops;
  uint32_t fn = 0, bn = 0;
  auto front_len = [&]() { return (ops[f] >> 4) - fn - ((f + 1 == b) ? bn : 0u); };
  auto back_len = [&]() { return (ops[b - 1] >> 4) - bn - ((f + 1 == b) ? fn : 0u); };
  auto live = [&]() {
    while (f < b && front_len() == 0) {  // exhausted (or zero-length) front element
      if (f + 1 == b) { f = b; break; }
      ++f;
      fn = 0;
    }
    return f < b;
  };
  auto live_back = [&]() {
    while (f < b && back_len() == 0) {
      if (f + 1 == b) { b = f; break; }
      --b;
      bn = 0;
    }
    return f < b;
  };
  // bases left of the window
  while (start_pos <= lo && live()) {
    const uint32_t run = front_len();
    switch (classify_bam_op(ops[f])) {
      case OP_ALIGNED: {
        const uint32_t t = (uint32_t)std::min<int64_t>(run, (int64_t)lo - start_pos + 1);
        ltrim += t; start_pos += (int32_t)t; fn += t;
        break;
      }
      case OP_DEL: {
        const uint32_t t = (uint32_t)std::min<int64_t>(run, (int64_t)lo - start_pos + 1);
        start_pos += (int32_t)t; fn += t;
        break;
      }
      case OP_READ_ONLY: ltrim += run; fn += run; break;
      case OP_NONE: fn += run; break;
      default: return false;
    }
  }
  // inside the left pad
  for (int32_t mid = start_pos; mid > lo && mid <= lo + pad && live();) {
    switch (classify_bam_op(ops[f])) {
      case OP_ALIGNED: mid++; break;
      case OP_DEL: ltrim--; mid++; break;
      case OP_READ_ONLY: case OP_NONE: break;
      default: return false;
    }
    fn += 1;
  }
  // bases right of the window
  while (end_pos > hi && live_back()) {
    const uint32_t run = back_len();
    switch (classify_bam_op(ops[b - 1])) {
      case OP_ALIGNED: {
        const uint32_t t = (uint32_t)std::min<int64_t>(run, (int64_t)end_pos - hi);
        rtrim += t; end_pos -= (int32_t)t; bn += t;
        break;
      }
      case OP_DEL: {
        const uint32_t t = (uint32_t)std::min<int64_t>(run, (int64_t)end_pos - hi);
        end_pos -= (int32_t)t; bn += t;
        break;
      }
      case OP_READ_ONLY: rtrim += run; bn += run; break;
      case OP_NONE: bn += run; break;
      default: return false;
    }
  }
  // inside the right pad
  for (int32_t mid = end_pos; mid > hi - pad && mid <= hi && live_back();) {
    switch (classify_bam_op(ops[b - 1])) {
      case OP_ALIGNED: mid--; break;
      case OP_DEL: rtrim--; mid--; break;
      case OP_READ_ONLY: case OP_NONE: break;
      default: return false;
    }
    bn += 1;
  }
  if (ltrim < 0) ltrim = 0;
  if (rtrim < 0) rtrim = 0;
  if (ltrim + rtrim > seq_len) return false;  // the reference asserts (HapAligner.cpp:463)
  out_ltrim = (int32_t)ltrim;
  out_len = seq_len - (int32_t)ltrim - (int32_t)rtrim;
  return true;
}

inline uint64_t mix64(uint64_t a, uint64_t b) {
  const unsigned __int128 r = (unsigned __int128)a * b;
  return (uint64_t)r ^ (uint64_t)(r >> 64);
}
inline uint64_t hash_bytes(const uint8_t* p, uint32_t n) {  // filter only: equal pools are confirmed with memcmp
  const uint64_t k0 = 0x9E3779B97F4A7C15ull, k1 = 0xD6E8FEB86659FD93ull, k2 = 0xFF51AFD7ED558CCDull, k3 = 0xC4CEB9FE1A85EC53ull;
  uint64_t s0 = k0 ^ n, s1 = k1;
  uint32_t i = 0;
  for (; i + 32 <= n; i += 32) {
    uint64_t w[4];
    memcpy(w, p + i, 32);
    s0 = mix64(w[0] ^ k2, w[1] ^ s0);
    s1 = mix64(w[2] ^ k3, w[3] ^ s1);
  }
  uint64_t w[4] = {0, 0, 0, 0};
  if (i < n) memcpy(w, p + i, n - i);
  s0 = mix64(w[0] ^ k2, w[1] ^ s0);
  s1 = mix64(w[2] ^ k3, w[3] ^ s1);
  return mix64(s0 ^ k1, s1 ^ k0);
}

// One chunk of loci in flight: the flattened job in pinned memory, its outputs, the per-locus bookkeeping.
struct Chunk {
  uint32_t l0 = 0, l1 = 0;  // loci [l0, l1) of the caller's batch
  int device_slot = 0;
  ltr_job* job = nullptr;
  // per locus of the chunk
  std::vector<int32_t> status;
  std::vector<uint32_t> n_pools, hap_bytes_l, read_bytes_l;
  // per raw read of the chunk (index relative to the chunk's first raw read)
  std::vector<uint32_t> pool_of;              // pool of the read, relative to its locus
  std::vector<uint32_t> pool_rep;             // at [first raw read of the locus + k]: representative read of pool k
  std::vector<int32_t> pool_ltrim, pool_len;  // ... its trim (len < 0: the 10 bp pseudo read)
  // the job, pinned
  Pinned<uint32_t> lhb, lrb, hap_off, read_off, lsb, pool_index, nsamp;
  Pinned<uint8_t> hap_bytes, read_bytes, haploid, mate, kept;
  Pinned<int32_t> label;
  Pinned<double> p1, p2, post, totals;
  Pinned<double> ll;                          // LL of (pool, allele), only when the per-read allele assignment is wanted
  std::vector<unsigned long long> ll_off;     // per locus of the chunk: pools x alleles prefix
  std::vector<unsigned long long> post_off, tot_off;
  uint32_t n_haps = 0, n_preads = 0, n_sreads = 0;
  void release() {
    lhb.release(); lrb.release(); hap_off.release(); read_off.release(); lsb.release(); pool_index.release(); nsamp.release();
    hap_bytes.release(); read_bytes.release(); haploid.release(); mate.release(); kept.release(); label.release();
    p1.release(); p2.release(); post.release(); totals.release(); ll.release();
  }
};

}  // namespace

struct ltr_genotyper {
  std::vector<ltr_ctx*> ctxs;
  WorkerPool* pool = nullptr;
  uint32_t chunk_loci = 20000;
  bool want_read_alleles = false;  // ltr_genotyper_set_read_alleles
  bool want_phased_gls = false;    // ltr_genotyper_set_phased_gls
  std::vector<Chunk*> chunks;  // 2 per device + 1 being prepared
};

namespace {

// Pools, trims and sizes of the loci [l0, l1) (parallel over loci); returns false on allocation failure.
void prepare_pass1(const ltr_locus_batch& B, const ltr_params& prm, Chunk& C, WorkerPool& pool) {
  const uint32_t n = C.l1 - C.l0;
  const uint32_t rbase = B.locus_read_begin[C.l0];
  const uint32_t n_raw = B.locus_read_begin[C.l1] - rbase;
  C.status.assign(n, LTR_OK);
  C.n_pools.assign(n, 0);
  C.hap_bytes_l.assign(n, 0);
  C.read_bytes_l.assign(n, 0);
  C.pool_of.resize(n_raw);
  C.pool_rep.resize(n_raw);
  C.pool_ltrim.resize(n_raw);
  C.pool_len.resize(n_raw);
  pool.parallel_for(n, 64, [&](uint32_t i0, uint32_t i1, int) {
    std::vector<uint64_t> hashes;
    for (uint32_t i = i0; i < i1; ++i) {
      const uint32_t l = C.l0 + i;
      const uint32_t r0 = B.locus_read_begin[l], r1 = B.locus_read_begin[l + 1];
      const uint32_t a0 = B.locus_allele_begin[l], a1 = B.locus_allele_begin[l + 1];
      const uint32_t lf = B.lflank_off[l + 1] - B.lflank_off[l], rf = B.rflank_off[l + 1] - B.rflank_off[l];
      const uint32_t S = B.locus_n_samples[l];
      int32_t st = LTR_OK;
      if (a1 <= a0 || r1 < r0 || B.lflank_off[l + 1] < B.lflank_off[l] || B.rflank_off[l + 1] < B.rflank_off[l]) st = LTR_ERR_INVALID;
      // reads: sample-major, labels in range, phasing terms <= 0 (the reference asserts, genotyper.h:104)
      int32_t prev = 0;
      for (uint32_t r = r0; r < r1 && st == LTR_OK; ++r) {
        const int32_t s = B.read_sample[r];
        if (s < prev || s < 0 || (uint32_t)s >= S || !(B.log_p1[r] <= 0.0) || !(B.log_p2[r] <= 0.0) ||
            B.read_off[r + 1] < B.read_off[r] || B.cigar_off[r + 1] < B.cigar_off[r])
          st = LTR_ERR_INVALID;
        prev = s;
      }
      uint32_t hb = 0;
      for (uint32_t a = a0; a < a1 && st == LTR_OK; ++a) {
        if (B.allele_off[a + 1] < B.allele_off[a]) st = LTR_ERR_INVALID;
        hb += lf + (B.allele_off[a + 1] - B.allele_off[a]) + rf;
      }
      if (st != LTR_OK) {
        C.status[i] = st;
        continue;
      }
      // ReadPooler::add_alignment: one pool per distinct sequence, in order of first appearance
      hashes.clear();
      uint32_t np = 0, rb = 0;
      const uint32_t base = r0 - rbase;
      for (uint32_t r = r0; r < r1; ++r) {
        const uint8_t* s = B.read_bytes + B.read_off[r];
        const uint32_t len = B.read_off[r + 1] - B.read_off[r];
        const uint64_t h = hash_bytes(s, len);
        uint32_t k = 0;
        for (; k < np; ++k) {
          if (hashes[k] != h) continue;
          const uint32_t q = C.pool_rep[base + k];
          if (B.read_off[q + 1] - B.read_off[q] == len && memcmp(B.read_bytes + B.read_off[q], s, len) == 0) break;
        }
        if (k == np) {
          hashes.push_back(h);
          C.pool_rep[base + np] = r;
          ++np;
        }
        C.pool_of[r - rbase] = k;
      }
      // HapAligner::trim_alignment per pool (the pooled alignment is its first read's, read_pooler.cpp:9-11)
      for (uint32_t k = 0; k < np && st == LTR_OK; ++k) {
        const uint32_t q = C.pool_rep[base + k];
        const int32_t len = (int32_t)(B.read_off[q + 1] - B.read_off[q]);
        int32_t ltrim = 0, tlen = 0;
        if (!trim_packed(B.cigar_ops + B.cigar_off[q], B.cigar_off[q + 1] - B.cigar_off[q], B.read_start[q], B.read_stop[q], len,
                         B.repeat_start[l], B.repeat_end[l], prm.indel_flank_len, ltrim, tlen)) {
          st = LTR_ERR_INVALID;
          break;
        }
        if (tlen == 0) {  // HapAligner.cpp:820-823: 10 bp pseudo read from the flanks
          if (lf < 5 || rf < 5) { st = LTR_ERR_INVALID; break; }
          tlen = -10;
        }
        C.pool_ltrim[base + k] = ltrim;
        C.pool_len[base + k] = tlen;
        rb += (uint32_t)(tlen < 0 ? -tlen : tlen);
      }
      if (st != LTR_OK) {
        C.status[i] = st;
        continue;
      }
      C.n_pools[i] = np;
      C.hap_bytes_l[i] = hb;
      C.read_bytes_l[i] = rb;
    }
  });
}

// Offsets (serial prefix sums) and the arrays of the job (parallel over loci).
bool prepare_pass2(const ltr_locus_batch& B, Chunk& C, WorkerPool& pool) {
  const uint32_t n = C.l1 - C.l0;
  const uint32_t rbase = B.locus_read_begin[C.l0];
  if (!C.lhb.reserve(n + 1) || !C.lrb.reserve(n + 1) || !C.lsb.reserve(n + 1) || !C.nsamp.reserve(n + 1) || !C.haploid.reserve(n + 1))
    return false;
  std::vector<uint64_t> hb_off(n + 1, 0), rb_off(n + 1, 0);
  C.post_off.assign(n + 1, 0);
  C.tot_off.assign(n + 1, 0);
  C.lhb.p[0] = C.lrb.p[0] = C.lsb.p[0] = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t l = C.l0 + i;
    const bool ok = C.status[i] == LTR_OK;
    const uint32_t H = ok ? B.locus_allele_begin[l + 1] - B.locus_allele_begin[l] : 0u;
    const uint32_t R = ok ? B.locus_read_begin[l + 1] - B.locus_read_begin[l] : 0u;
    const uint32_t S = ok ? B.locus_n_samples[l] : 0u;
    C.lhb.p[i + 1] = C.lhb.p[i] + H;
    C.lrb.p[i + 1] = C.lrb.p[i] + (ok ? C.n_pools[i] : 0u);
    C.lsb.p[i + 1] = C.lsb.p[i] + R;
    C.nsamp.p[i] = S;
    C.haploid.p[i] = (ok && B.locus_haploid && B.locus_haploid[l]) ? 1 : 0;
    hb_off[i + 1] = hb_off[i] + C.hap_bytes_l[i];
    rb_off[i + 1] = rb_off[i] + C.read_bytes_l[i];
    C.post_off[i + 1] = C.post_off[i] + (unsigned long long)S * H * H;
    C.tot_off[i + 1] = C.tot_off[i] + S;
  }
  C.n_haps = C.lhb.p[n];
  C.n_preads = C.lrb.p[n];
  C.n_sreads = C.lsb.p[n];
  if (hb_off[n] > 0xFFFFFFF0ull || rb_off[n] > 0xFFFFFFF0ull) return false;
  if (!C.hap_off.reserve((size_t)C.n_haps + 1) || !C.hap_bytes.reserve(hb_off[n] + 16) || !C.read_off.reserve((size_t)C.n_preads + 1) ||
      !C.read_bytes.reserve(rb_off[n] + 16) || !C.pool_index.reserve((size_t)C.n_sreads + 1) || !C.label.reserve((size_t)C.n_sreads + 1) ||
      !C.p1.reserve((size_t)C.n_sreads + 1) || !C.p2.reserve((size_t)C.n_sreads + 1) || !C.kept.reserve((size_t)C.n_haps + 1) ||
      !C.post.reserve(C.post_off[n] + 1) || !C.totals.reserve(C.tot_off[n] + 1))
    return false;
  if (B.second_mate && !C.mate.reserve((size_t)C.n_sreads + 1)) return false;
  pool.parallel_for(n, 64, [&](uint32_t i0, uint32_t i1, int) {
    for (uint32_t i = i0; i < i1; ++i) {
      if (C.status[i] != LTR_OK) continue;
      const uint32_t l = C.l0 + i;
      const uint8_t* lfl = B.lflank_bytes + B.lflank_off[l];
      const uint8_t* rfl = B.rflank_bytes + B.rflank_off[l];
      const uint32_t lf = B.lflank_off[l + 1] - B.lflank_off[l], rf = B.rflank_off[l + 1] - B.rflank_off[l];
      // haplotypes in column order: one multi-allele block, so the gray-code counter is the allele index
      // (Haplotype.cpp:157-196); sequence = left flank + allele + right flank (Haplotype::get_seq)
      uint32_t off = (uint32_t)hb_off[i];
      uint32_t h = C.lhb.p[i];
      for (uint32_t a = B.locus_allele_begin[l]; a < B.locus_allele_begin[l + 1]; ++a, ++h) {
        const uint32_t al = B.allele_off[a + 1] - B.allele_off[a];
        C.hap_off.p[h] = off;
        memcpy(C.hap_bytes.p + off, lfl, lf);
        memcpy(C.hap_bytes.p + off + lf, B.allele_bytes + B.allele_off[a], al);
        memcpy(C.hap_bytes.p + off + lf + al, rfl, rf);
        off += lf + al + rf;
      }
      // trimmed pools
      const uint32_t base = B.locus_read_begin[l] - rbase;
      uint32_t roff = (uint32_t)rb_off[i];
      for (uint32_t k = 0; k < C.n_pools[i]; ++k) {
        const uint32_t q = C.pool_rep[base + k];
        C.read_off.p[C.lrb.p[i] + k] = roff;
        const int32_t tl = C.pool_len[base + k];
        if (tl < 0) {
          memcpy(C.read_bytes.p + roff, lfl + lf - 5, 5);
          memcpy(C.read_bytes.p + roff + 5, rfl, 5);
          roff += 10;
        } else {
          memcpy(C.read_bytes.p + roff, B.read_bytes + B.read_off[q] + C.pool_ltrim[base + k], (size_t)tl);
          roff += (uint32_t)tl;
        }
      }
      // sample-reads
      uint32_t sr = C.lsb.p[i];
      for (uint32_t r = B.locus_read_begin[l]; r < B.locus_read_begin[l + 1]; ++r, ++sr) {
        C.pool_index.p[sr] = C.pool_of[r - rbase];
        C.label.p[sr] = B.read_sample[r];
        C.p1.p[sr] = B.log_p1[r];
        C.p2.p[sr] = B.log_p2[r];
        if (B.second_mate) C.mate.p[sr] = B.second_mate[r];
      }
    }
  });
  C.hap_off.p[C.n_haps] = (uint32_t)hb_off[n];
  C.read_off.p[C.n_preads] = (uint32_t)rb_off[n];
  return true;
}

struct CallsOwner {  // storage behind an ltr_batch_calls
  ltr_batch_calls view;
  std::vector<int32_t> status, n_kept, n_pools, gts, n_reads, pls;
  std::vector<uint32_t> lsb, lab;
  std::vector<uint8_t> kept;
  std::vector<double> lpp, lup, gld, stl, gls;
  std::vector<uint64_t> glb;
  std::vector<int32_t> read_allele;  // per raw read of the batch, -1 = not assigned
  std::vector<uint64_t> pglb;        // only with ltr_genotyper_set_phased_gls
  std::vector<double> pgls;
};

// Genotyper::extract_genotypes_and_likelihoods per locus on the surviving alleles (parallel over loci).
void finish_chunk(const ltr_locus_batch& B, Chunk& C, CallsOwner& O, WorkerPool& pool) {
  const uint32_t n = C.l1 - C.l0;
  pool.parallel_for(n, 32, [&](uint32_t i0, uint32_t i1, int) {
    std::vector<int32_t> gts, pls;
    std::vector<double> lpp, lup, hlp, hlu, gls, pgl, gld, stl;
    for (uint32_t i = i0; i < i1; ++i) {
      const uint32_t l = C.l0 + i;
      O.status[l] = C.status[i];
      O.n_pools[l] = (int32_t)C.n_pools[i];
      if (C.status[i] != LTR_OK) continue;
      const uint32_t a0 = B.locus_allele_begin[l], H = B.locus_allele_begin[l + 1] - a0, S = B.locus_n_samples[l];
      const uint8_t* km = C.kept.p + C.lhb.p[i];
      std::vector<int32_t> kept_idx;
      for (uint32_t a = 0; a < H; ++a) {
        O.kept[a0 + a] = km[a];
        if (km[a]) kept_idx.push_back((int32_t)a);
      }
      const int32_t K = (int32_t)kept_idx.size();
      O.n_kept[l] = K;
      const uint32_t s0 = O.lsb[l];
      for (uint32_t r = B.locus_read_begin[l]; r < B.locus_read_begin[l + 1]; ++r) O.n_reads[s0 + (uint32_t)B.read_sample[r]] += 1;
      if (S == 0 || K == 0) continue;
      const bool haploid = C.haploid.p[i] != 0;
      const size_t n_gl = haploid ? (size_t)K : (size_t)K * (K + 1) / 2, n_pgl = haploid ? (size_t)K : (size_t)K * K;
      gts.assign(2 * S, 0); pls.assign(S * n_gl, 0);
      lpp.assign(S, 0); lup.assign(S, 0); hlp.assign(S, 0); hlu.assign(S, 0); gld.assign(S, 0); stl.assign(S, 0);
      gls.assign(S * n_gl, 0); pgl.assign(S * n_pgl, 0);
      ltr_locus_calls c;
      memset(&c, 0, sizeof(c));
      c.best_gts = gts.data(); c.log_phased_posteriors = lpp.data(); c.log_unphased_posteriors = lup.data();
      c.hap_log_phased_posteriors = hlp.data(); c.hap_log_unphased_posteriors = hlu.data(); c.gls = gls.data();
      c.pls = pls.data(); c.phased_gls = pgl.data(); c.gl_diffs = gld.data(); c.sample_total_lls = stl.data();
      const int rc = ltr_extract_calls(haploid ? 1 : 0, (int32_t)S, K, C.post.p + C.post_off[i], C.totals.p + C.tot_off[i], &c);
      if (rc != LTR_OK) {
        O.status[l] = rc;
        continue;
      }
      for (uint32_t s = 0; s < S; ++s) {
        O.gts[2 * (s0 + s)] = kept_idx[(size_t)gts[2 * s]];
        O.gts[2 * (s0 + s) + 1] = kept_idx[(size_t)gts[2 * s + 1]];
        O.lpp[s0 + s] = lpp[s];
        O.lup[s0 + s] = lup[s];
        O.gld[s0 + s] = gld[s];
        O.stl[s0 + s] = stl[s];
        const uint64_t g0 = O.glb[s0 + s];
        for (size_t k = 0; k < n_gl; ++k) {
          O.gls[g0 + k] = gls[s * n_gl + k];
          O.pls[g0 + k] = pls[s * n_gl + k];
        }
        if (!O.pglb.empty()) std::copy(pgl.begin() + (ptrdiff_t)(s * n_pgl), pgl.begin() + (ptrdiff_t)((s + 1) * n_pgl), O.pgls.begin() + (ptrdiff_t)O.pglb[s0 + s]);
      }
      if (!O.read_allele.empty() && !C.ll_off.empty()) {
        // which allele of its sample's genotype a read supports (write_vcf_record, seq_stutter_genotyper.cpp:954-970): the
        // first unless log_p2 + LL[second] >= log_p1 + LL[first]; LLs clamped at -600 like the posterior pass leaves them
        const double* ll = C.ll.p + C.ll_off[i];
        const uint32_t r_first = B.locus_read_begin[C.l0];
        for (uint32_t r = B.locus_read_begin[l]; r < B.locus_read_begin[l + 1]; ++r) {
          const uint32_t s = (uint32_t)B.read_sample[r];
          const int32_t ga = O.gts[2 * (s0 + s)], gb = O.gts[2 * (s0 + s) + 1];
          int32_t best = ga;
          if (!haploid && ga != gb) {
            const double* row = ll + (size_t)C.pool_of[r - r_first] * H;
            const double la = row[ga] < -600 ? -600 : row[ga], lb = row[gb] < -600 ? -600 : row[gb];
            best = (B.log_p1[r] + la > B.log_p2[r] + lb) ? ga : gb;
          }
          O.read_allele[r] = best;
        }
      }
    }
  });
}

}  // namespace

extern "C" {

int ltr_genotyper_create(const int32_t* devices, int32_t n_devices, int32_t host_threads, int32_t chunk_loci,
                         ltr_genotyper** out) {
  if (!out || !devices || n_devices < 1) return LTR_ERR_INVALID;
  *out = nullptr;
  ltr_genotyper* g = new ltr_genotyper();
  for (int d = 0; d < n_devices; ++d) {
    ltr_ctx* ctx = nullptr;
    const int rc = ltr_ctx_create(devices[d], &ctx);
    if (rc != LTR_OK) {
      ltr_genotyper_destroy(g);
      return rc;
    }
    g->ctxs.push_back(ctx);
  }
  int nt = host_threads > 0 ? host_threads : (int)std::max(1u, std::thread::hardware_concurrency());
  g->pool = new WorkerPool(nt);
  if (chunk_loci > 0) g->chunk_loci = (uint32_t)chunk_loci;
  for (size_t k = 0; k < 2 * g->ctxs.size() + 1; ++k) g->chunks.push_back(new Chunk());
  *out = g;
  return LTR_OK;
}

int ltr_genotyper_set_phased_gls(ltr_genotyper* g, int32_t on) {
  if (!g) return LTR_ERR_INVALID;
  g->want_phased_gls = on != 0;
  return LTR_OK;
}

int ltr_genotyper_set_read_alleles(ltr_genotyper* g, int32_t on) {
  if (!g) return LTR_ERR_INVALID;
  g->want_read_alleles = on != 0;
  return LTR_OK;
}

void ltr_genotyper_destroy(ltr_genotyper* g) {
  if (!g) return;
  for (Chunk* c : g->chunks) {
    if (c->job) ltr_job_destroy(g->ctxs[(size_t)c->device_slot], c->job);
    c->release();
    delete c;
  }
  delete g->pool;
  for (ltr_ctx* ctx : g->ctxs) ltr_ctx_destroy(ctx);
  delete g;
}

void ltr_batch_calls_free(ltr_batch_calls* calls) {
  if (calls) delete reinterpret_cast<CallsOwner*>(calls);  // view is the first member
}

int ltr_genotyper_run(ltr_genotyper* g, const ltr_params* params, const ltr_locus_batch* batch, ltr_batch_calls** out) {
  if (!g || !params || !batch || !out) return LTR_ERR_INVALID;
  *out = nullptr;
  const ltr_locus_batch& B = *batch;
  const uint32_t n_loci = B.n_loci;
  if (n_loci && (!B.lflank_off || !B.rflank_off || !B.locus_allele_begin || !B.allele_off || !B.repeat_start || !B.repeat_end ||
                 !B.locus_read_begin || !B.read_off || !B.cigar_off || !B.locus_n_samples))
    return LTR_ERR_INVALID;
  if (params->indel_flank_len < 0 || params->indel_flank_len > 35) return LTR_ERR_INVALID;
  if (params->indel_flank_len < 5) return LTR_ERR_UNSUPPORTED;  // see job_new (abi.cu)
  const auto t_begin = Clock::now();
  static const uint32_t kZero[1] = {0};
  const uint32_t* lab = n_loci ? B.locus_allele_begin : kZero;
  const uint32_t* lrb = n_loci ? B.locus_read_begin : kZero;
  for (uint32_t l = 0; l < n_loci; ++l)
    if (lab[l + 1] < lab[l] || lrb[l + 1] < lrb[l]) return LTR_ERR_INVALID;
  const uint32_t n_reads = lrb[n_loci];
  if (n_reads && (!B.read_start || !B.read_stop || !B.read_sample || !B.log_p1 || !B.log_p2)) return LTR_ERR_INVALID;
  CallsOwner* O = new CallsOwner();
  O->status.assign(n_loci, LTR_OK);
  O->n_kept.assign(n_loci, 0);
  O->n_pools.assign(n_loci, 0);
  O->lsb.assign((size_t)n_loci + 1, 0);
  O->lab.assign(lab, lab + n_loci + 1);
  for (uint32_t l = 0; l < n_loci; ++l) O->lsb[l + 1] = O->lsb[l] + B.locus_n_samples[l];
  const uint32_t n_samples = O->lsb[n_loci];
  O->kept.assign(lab[n_loci], 0);
  O->gts.assign((size_t)2 * n_samples, -1);
  O->n_reads.assign(n_samples, 0);
  O->lpp.assign(n_samples, 0.0); O->lup.assign(n_samples, 0.0); O->gld.assign(n_samples, 0.0); O->stl.assign(n_samples, 0.0);
  O->glb.assign((size_t)n_samples + 1, 0);
  for (uint32_t l = 0; l < n_loci; ++l) {
    const uint64_t H = lab[l + 1] - lab[l];
    const bool hap = B.locus_haploid && B.locus_haploid[l];
    for (uint32_t s = O->lsb[l]; s < O->lsb[l + 1]; ++s) O->glb[s + 1] = O->glb[s] + (hap ? H : H * (H + 1) / 2);
  }
  O->gls.assign(O->glb[n_samples], 0.0);
  O->pls.assign(O->glb[n_samples], 0);
  if (g->want_read_alleles) O->read_allele.assign((size_t)n_reads + 1, -1);
  if (g->want_phased_gls) {
    O->pglb.assign((size_t)n_samples + 1, 0);
    for (uint32_t l = 0; l < n_loci; ++l) {
      const uint64_t H = lab[l + 1] - lab[l];
      const bool hap = B.locus_haploid && B.locus_haploid[l];
      for (uint32_t s = O->lsb[l]; s < O->lsb[l + 1]; ++s) O->pglb[s + 1] = O->pglb[s] + (hap ? H : H * H);
    }
    O->pgls.assign(O->pglb[n_samples], 0.0);
  }

  double prep_ms = 0, wait_ms = 0, post_ms = 0, submit_ms = 0;
  int rc_all = LTR_OK;
  std::deque<Chunk*> inflight;
  std::vector<Chunk*> free_chunks(g->chunks.begin(), g->chunks.end());
  const size_t n_dev = g->ctxs.size();
  auto retire = [&](Chunk* c) {
    const auto t0 = Clock::now();
    const int rc = ltr_job_wait(g->ctxs[(size_t)c->device_slot], c->job);
    wait_ms += ms_since(t0);
    ltr_job_destroy(g->ctxs[(size_t)c->device_slot], c->job);
    c->job = nullptr;
    if (rc != LTR_OK) {
      // the whole job failed (a malformed array only the device sees, or a CUDA error): every locus of the chunk reports it
      for (int32_t& s : c->status)
        if (s == LTR_OK) s = rc;
      if (rc == LTR_ERR_CUDA || rc == LTR_ERR_OOM) rc_all = rc;
    }
    const auto t1 = Clock::now();
    finish_chunk(B, *c, *O, *g->pool);
    post_ms += ms_since(t1);
    free_chunks.push_back(c);
  };
  // Chunk boundaries.  One device: chunk_loci loci per chunk.  Several devices: chunks of about equal estimated work (reads x
  // allele bytes x allele length: VNTR loci differ by two orders of magnitude), at least four per device so that every device
  // has a chunk running and one queued; a new chunk goes to the device with the fewest chunks in flight and whichever chunk
  // finishes first is retired first.
  std::vector<uint32_t> cuts;
  if (n_dev <= 1 || n_loci == 0) {
    for (uint32_t l0 = 0; l0 < n_loci; l0 += g->chunk_loci) cuts.push_back(std::min(n_loci, l0 + g->chunk_loci));
  } else {
    std::vector<double> cost(n_loci);
    double total = 0;
    for (uint32_t l = 0; l < n_loci; ++l) {
      const double H = (double)(lab[l + 1] - lab[l]);
      const double sumA = B.allele_off ? (double)(B.allele_off[lab[l + 1]] - B.allele_off[lab[l]]) : 0.0;
      const double R = (double)(lrb[l + 1] - lrb[l]);
      cost[l] = 1.0 + R * (sumA + 70.0 * H) * (H > 0 ? sumA / H + 70.0 : 0.0);
      total += cost[l];
    }
    const double target = total / (4.0 * (double)n_dev);
    const uint32_t min_loci = 64;
    double acc = 0;
    uint32_t begin = 0;
    for (uint32_t l = 0; l < n_loci; ++l) {
      acc += cost[l];
      const uint32_t n = l + 1 - begin;
      if (n >= g->chunk_loci || (acc >= target && n >= min_loci)) {
        cuts.push_back(l + 1);
        begin = l + 1;
        acc = 0;
      }
    }
    if (begin < n_loci) cuts.push_back(n_loci);
  }
  std::vector<int> dev_load(n_dev, 0);
  auto retire_one = [&]() {  // a finished chunk if there is one, else the oldest
    size_t pick = 0;
    if (n_dev > 1)
      for (size_t k = 0; k < inflight.size(); ++k)
        if (ltr_job_poll(g->ctxs[(size_t)inflight[k]->device_slot], inflight[k]->job) == 1) {
          pick = k;
          break;
        }
    Chunk* c = inflight[pick];
    inflight.erase(inflight.begin() + (long)pick);
    --dev_load[(size_t)c->device_slot];
    retire(c);
  };
  uint32_t l0 = 0;
  for (size_t ci = 0; ci < cuts.size() && rc_all == LTR_OK; l0 = cuts[ci], ++ci) {
    while (free_chunks.empty() || inflight.size() >= 2 * n_dev) retire_one();
    Chunk* c = free_chunks.back();
    free_chunks.pop_back();
    c->l0 = l0;
    c->l1 = cuts[ci];
    c->device_slot = (int)(std::min_element(dev_load.begin(), dev_load.end()) - dev_load.begin());
    ++dev_load[(size_t)c->device_slot];
    const auto t0 = Clock::now();
    prepare_pass1(B, *params, *c, *g->pool);
    const bool ok = prepare_pass2(B, *c, *g->pool);
    prep_ms += ms_since(t0);
    if (!ok) {
      rc_all = LTR_ERR_OOM;
      --dev_load[(size_t)c->device_slot];
      free_chunks.push_back(c);
      break;
    }
    ltr_viterbi_batch vb;
    vb.n_loci = c->l1 - c->l0;
    vb.locus_hap_begin = c->lhb.p; vb.locus_read_begin = c->lrb.p; vb.hap_off = c->hap_off.p; vb.hap_bytes = c->hap_bytes.p;
    vb.read_off = c->read_off.p; vb.read_bytes = c->read_bytes.p;
    ltr_posterior_batch pb;
    memset(&pb, 0, sizeof(pb));
    pb.locus_sread_begin = c->lsb.p; pb.pool_index = c->pool_index.p; pb.sample_label = c->label.p; pb.log_p1 = c->p1.p;
    pb.log_p2 = c->p2.p; pb.locus_n_samples = c->nsamp.p; pb.locus_haploid = c->haploid.p;
    pb.second_mate = B.second_mate ? c->mate.p : nullptr;
    pb.prune_uncalled = 1;
    ltr_job_outputs outs;
    outs.ll = nullptr; outs.post = c->post.p; outs.totals = c->totals.p; outs.kept_mask = c->kept.p;
    c->ll_off.clear();
    if (g->want_read_alleles) {
      const uint32_t nc = c->l1 - c->l0;
      c->ll_off.assign((size_t)nc + 1, 0);
      for (uint32_t i = 0; i < nc; ++i)
        c->ll_off[i + 1] = c->ll_off[i] + (unsigned long long)(c->lrb.p[i + 1] - c->lrb.p[i]) * (c->lhb.p[i + 1] - c->lhb.p[i]);
      if (!c->ll.reserve((size_t)c->ll_off[nc] + 1)) {
        rc_all = LTR_ERR_OOM;
        --dev_load[(size_t)c->device_slot];
        free_chunks.push_back(c);
        break;
      }
      outs.ll = c->ll.p;
    }
    const auto t_submit = Clock::now();
    const int rc = ltr_job_submit_outputs(g->ctxs[(size_t)c->device_slot], params, &vb, &pb, &outs, &c->job);
    submit_ms += ms_since(t_submit);
    if (rc != LTR_OK) {
      for (int32_t& s : c->status)
        if (s == LTR_OK) s = rc;
      if (rc == LTR_ERR_CUDA || rc == LTR_ERR_OOM) rc_all = rc;
      for (uint32_t i = 0; i < c->l1 - c->l0; ++i) O->status[c->l0 + i] = c->status[i];
      --dev_load[(size_t)c->device_slot];
      free_chunks.push_back(c);
      continue;
    }
    inflight.push_back(c);
  }
  while (!inflight.empty()) retire_one();
  ltr_batch_calls& V = O->view;
  V.n_loci = n_loci;
  V.status = O->status.data(); V.locus_sample_begin = O->lsb.data(); V.locus_allele_begin = O->lab.data();
  V.kept_mask = O->kept.data(); V.n_kept = O->n_kept.data(); V.n_pools = O->n_pools.data(); V.gts = O->gts.data();
  V.log_phased_posteriors = O->lpp.data(); V.log_unphased_posteriors = O->lup.data(); V.gl_diffs = O->gld.data();
  V.sample_total_lls = O->stl.data(); V.n_reads = O->n_reads.data(); V.gl_begin = O->glb.data(); V.gls = O->gls.data();
  V.pls = O->pls.data();
  V.prep_ms = prep_ms; V.gpu_wait_ms = wait_ms; V.post_ms = post_ms; V.total_ms = ms_since(t_begin);
  V.submit_ms = submit_ms; V.n_chunks = (uint32_t)cuts.size();
  V.read_allele = O->read_allele.empty() ? nullptr : O->read_allele.data();
  V.pgl_begin = O->pglb.empty() ? nullptr : O->pglb.data();
  V.phased_gls = O->pglb.empty() ? nullptr : O->pgls.data();
  if (rc_all != LTR_OK) {
    delete O;
    return rc_all;
  }
  *out = &O->view;
  return LTR_OK;
}

int32_t ltr_locus_batch_trim_read(const ltr_locus_batch* batch, const ltr_params* params, uint32_t locus, uint32_t read,
                                  uint8_t* out, int32_t cap) {
  if (!batch || !params || !out || locus >= batch->n_loci) return LTR_ERR_INVALID;
  const ltr_locus_batch& B = *batch;
  if (read < B.locus_read_begin[locus] || read >= B.locus_read_begin[locus + 1]) return LTR_ERR_INVALID;
  const int32_t len = (int32_t)(B.read_off[read + 1] - B.read_off[read]);
  int32_t ltrim = 0, tlen = 0;
  if (!trim_packed(B.cigar_ops + B.cigar_off[read], B.cigar_off[read + 1] - B.cigar_off[read], B.read_start[read],
                   B.read_stop[read], len, B.repeat_start[locus], B.repeat_end[locus], params->indel_flank_len, ltrim, tlen))
    return LTR_ERR_INVALID;
  if (tlen == 0) {
    const uint32_t lf = B.lflank_off[locus + 1] - B.lflank_off[locus], rf = B.rflank_off[locus + 1] - B.rflank_off[locus];
    if (lf < 5 || rf < 5 || cap < 10) return LTR_ERR_INVALID;
    memcpy(out, B.lflank_bytes + B.lflank_off[locus] + lf - 5, 5);
    memcpy(out + 5, B.rflank_bytes + B.rflank_off[locus], 5);
    return 10;
  }
  if (tlen > cap) return LTR_ERR_INVALID;
  memcpy(out, B.read_bytes + B.read_off[read] + ltrim, (size_t)tlen);
  return tlen;
}

}  // extern "C"
