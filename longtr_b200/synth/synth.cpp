// synth.cpp -- deterministic synthetic TR loci, emitted directly as flattened batches.
//
// Implements the seeded generators of SURVEY.md section 8(d) for BASELINE.json's synthetic
// configurations (the benchmark and the full-size property tests use them; no dataset can be
// downloaded here):
//   config 3  HiFi STRs : motif 1-6 bp, reference repeat 50-300 bp, alleles ref +- k units
//                         (k in -3..3), 30 reads, substitutions 1e-3, indels 1e-3 (x5 inside
//                         homopolymer runs >= 4), 2-6 candidate haplotypes
//   config 4  VNTRs     : motif 10-60 bp, reference repeat 500-1000 bp, alleles ref +- k units
//                         (k in -4..4), 2-12 candidate haplotypes (motif-level variants),
//                         ONT-like errors (sub 1 %, indel 1.5 %)
//   config 5  homopolymers: run 10-30 bp of A or T, alleles ref +- (-2..2), x10 indel rate in the run
// PRNG = std::mt19937_64 seeded with base_seed + locus index, so any locus range can be produced
// independently (locus-sharded multi-GPU runs draw disjoint ranges of the same job).
//
// Layout produced per locus follows what LongTR hands its hot path: candidate haplotypes are
// 35 bp flank + [5 bp pad + repeat + 5 bp pad] + 35 bp flank (HaplotypeGenerator keeps the pads
// inside the repeat block), reads are pooled by their full +-200 bp sequence
// (ReadPooler, src/read_pooler.cpp:3-20) and then cut to the block +- 5 bp like
// HapAligner::trim_alignment (HapAligner.cpp:346-465) does.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "longtr_synth.h"

namespace {

const char kBases[5] = "ACGT";

struct Rng {
  std::mt19937_64 g;
  explicit Rng(uint64_t s) : g(s) {}
  uint32_t below(uint32_t n) { return (uint32_t)(g() % n); }
  int range(int lo, int hi) { return lo + (int)below((uint32_t)(hi - lo + 1)); }  // inclusive
  double unif() { return (double)(g() >> 11) * (1.0 / 9007199254740992.0); }
  char base() { return kBases[g() & 3]; }
  std::string seq(int n) {
    std::string s((size_t)n, 'A');
    for (int i = 0; i < n; ++i) s[i] = base();
    return s;
  }
};

std::string rand_motif(Rng& r, int period) {
  for (;;) {
    std::string m = r.seq(period);
    bool periodic = false;
    for (int q = 1; q < period && !periodic; ++q) {
      if (period % q) continue;
      bool same = true;
      for (int i = q; i < period && same; ++i) same = (m[i] == m[i - q]);
      periodic = same;
    }
    if (!periodic) return m;
  }
}

std::string repeat_of(const std::string& motif, int units) {
  std::string s;
  s.reserve(motif.size() * (size_t)units);
  for (int u = 0; u < units; ++u) s += motif;
  return s;
}

// Sequencing errors on one segment.  homop_mult multiplies the indel rate inside homopolymer
// runs of length >= 4.
std::string add_errors(Rng& r, const std::string& s, double sub, double indel, double homop_mult) {
  std::string out;
  out.reserve(s.size() + 8);
  const int L = (int)s.size();
  for (int i = 0; i < L; ++i) {
    double ind = indel;
    if (homop_mult != 1.0) {
      int a = i, b = i;
      while (a > 0 && s[a - 1] == s[i]) --a;
      while (b + 1 < L && s[b + 1] == s[i]) ++b;
      if (b - a + 1 >= 4) ind *= homop_mult;
    }
    const double u = r.unif();
    if (u < sub) {
      out.push_back(kBases[(uint32_t)(std::strchr(kBases, s[i]) - kBases + 1 + r.below(3)) & 3]);
    } else if (u < sub + ind * 0.5) {
      continue;  // deletion
    } else if (u < sub + ind) {
      out.push_back(s[i]);
      out.push_back(r.base());  // insertion
    } else {
      out.push_back(s[i]);
    }
  }
  return out;
}

struct LocusOut {
  std::vector<std::string> haps;    // full haplotypes
  std::vector<std::string> reads;   // pooled, trimmed
  std::vector<uint32_t> pool_index; // per sample-read
  std::vector<double> p1, p2;
  // raw form (ltr_synth_generate_loci): the same reads whole, with CIGARs against the reference window
  bool want_raw = false;
  std::string lflank, rflank;
  std::vector<std::string> alleles;            // repeat-block alleles (pads inside), reference first
  std::vector<std::string> raw_reads;
  std::vector<std::vector<uint32_t>> raw_cigars;  // BAM encoding
  std::vector<int32_t> raw_start, raw_stop;
  int32_t repeat_start = 0, repeat_end = 0;
};

// CIGAR under construction: per read base / reference base operations, merged into runs.
struct CigarBuilder {
  std::vector<uint32_t> ops;  // len << 4 | op (BAM: M0 I1 D2 =7 X8)
  void add(uint32_t op, uint32_t n = 1) {
    if (n == 0) return;
    if (!ops.empty() && (ops.back() & 15u) == op) ops.back() += n << 4;
    else ops.push_back((n << 4) | op);
  }
};

// add_errors with the operations it performed, for a stretch of the TRUE haplotype whose bases are either aligned to the
// reference (ref_aligned[i] != 0) or part of an insertion relative to it.  Consumes the generator exactly like add_errors,
// so the read bases are identical to the flattened workload's.
std::string add_errors_ops(Rng& r, const std::string& s, const std::vector<uint8_t>& ref_aligned, size_t a0, double sub,
                           double indel, double homop_mult, CigarBuilder& cg) {
  std::string out;
  out.reserve(s.size() + 8);
  const int L = (int)s.size();
  for (int i = 0; i < L; ++i) {
    double ind = indel;
    if (homop_mult != 1.0) {
      int a = i, b = i;
      while (a > 0 && s[a - 1] == s[i]) --a;
      while (b + 1 < L && s[b + 1] == s[i]) ++b;
      if (b - a + 1 >= 4) ind *= homop_mult;
    }
    const bool al = ref_aligned[a0 + (size_t)i] != 0;
    const double u = r.unif();
    if (u < sub) {
      out.push_back(kBases[(uint32_t)(std::strchr(kBases, s[i]) - kBases + 1 + r.below(3)) & 3]);
      cg.add(al ? 8u : 1u);
    } else if (u < sub + ind * 0.5) {
      if (al) cg.add(2u);  // deletion
      continue;
    } else if (u < sub + ind) {
      out.push_back(s[i]);
      out.push_back(r.base());  // insertion
      cg.add(al ? 7u : 1u);
      cg.add(1u);
    } else {
      out.push_back(s[i]);
      cg.add(al ? 7u : 1u);
    }
  }
  return out;
}

void gen_locus(int config, uint64_t seed, LocusOut& o) {
  Rng r(seed);
  int period, ref_units, kmax, hmin, hmax;
  double sub, indel, homop_mult;
  std::string motif;
  if (config == 4) {
    period = r.range(10, 60);
    motif = rand_motif(r, period);
    const int ref_len = r.range(500, 1000);
    ref_units = std::max(2, ref_len / period);
    kmax = 4; hmin = 2; hmax = 12; sub = 0.01; indel = 0.015; homop_mult = 1.0;
  } else if (config == 5) {
    period = 1;
    motif = std::string(1, (r.g() & 1) ? 'A' : 'T');
    ref_units = r.range(10, 30);
    kmax = 2; hmin = 2; hmax = 5; sub = 1e-3; indel = 1e-3; homop_mult = 10.0;
  } else {
    period = r.range(1, 6);
    motif = rand_motif(r, period);
    const int ref_len = r.range(50, 300);
    ref_units = std::max(2, ref_len / period);
    kmax = 3; hmin = 2; hmax = 6; sub = 1e-3; indel = 1e-3; homop_mult = 5.0;
  }
  const std::string lctx = r.seq(165), rctx = r.seq(165);  // +-200 bp window beyond the 35 bp flanks
  const std::string lflank = r.seq(35), rflank = r.seq(35);
  const std::string lpad = r.seq(5), rpad = r.seq(5);
  auto allele_of = [&](int units, const std::string& mot) {
    return lpad + repeat_of(mot, std::max(1, units)) + rpad;
  };
  int k1 = r.range(-kmax, kmax), k2 = r.range(-kmax, kmax);
  std::vector<std::string> truth;
  truth.push_back(allele_of(ref_units + k1, motif));
  truth.push_back(allele_of(ref_units + k2, motif));
  const int truth_units[2] = {std::max(1, ref_units + k1), std::max(1, ref_units + k2)};
  // candidate set: reference first, then alternates sorted by (length, sequence)
  // (HaplotypeGenerator.cpp:475); decoys are +-1 unit neighbours / motif-level variants
  const int want_h = r.range(hmin, hmax);
  std::vector<std::string> alts;
  const std::string ref_allele = allele_of(ref_units, motif);
  auto add_alt = [&](const std::string& s) {
    if (s != ref_allele && std::find(alts.begin(), alts.end(), s) == alts.end()) alts.push_back(s);
  };
  add_alt(truth[0]);
  add_alt(truth[1]);
  int guard = 0;
  while ((int)alts.size() + 1 < want_h && guard++ < 64) {
    if (config == 4 && r.below(2)) {
      std::string mv = motif;  // motif-level variant: one substituted base in every copy
      mv[r.below((uint32_t)period)] = r.base();
      add_alt(allele_of(ref_units + r.range(-kmax, kmax), mv));
    } else {
      add_alt(allele_of(ref_units + r.range(-kmax - 1, kmax + 1), motif));
    }
  }
  while ((int)alts.size() + 1 > hmax) alts.pop_back();
  std::sort(alts.begin(), alts.end(), [](const std::string& a, const std::string& b) {
    return a.size() != b.size() ? a.size() < b.size() : a < b;
  });
  o.haps.clear();
  o.haps.push_back(lflank + ref_allele + rflank);
  for (const std::string& a : alts) o.haps.push_back(lflank + a + rflank);

  // 30 reads, allele chosen with p = 1/2, HP tag follows the allele (snp_bam_processor.h:16-18)
  std::map<std::string, uint32_t> pools;
  o.reads.clear(); o.pool_index.clear(); o.p1.clear(); o.p2.clear();
  if (o.want_raw) {
    o.lflank = lflank; o.rflank = rflank;
    o.alleles.clear();
    for (const std::string& h : o.haps) o.alleles.push_back(h.substr(35, h.size() - 70));
    o.repeat_start = 1000;
    o.repeat_end = 1000 + (int32_t)ref_allele.size();
    o.raw_reads.clear(); o.raw_cigars.clear(); o.raw_start.clear(); o.raw_stop.clear();
  }
  for (int i = 0; i < 30; ++i) {
    const int a = (int)(r.g() & 1);
    std::string left, mid, right;
    if (!o.want_raw) {
      left = add_errors(r, lctx + lflank.substr(0, 30), sub, indel, homop_mult);
      mid = add_errors(r, lflank.substr(30) + truth[a] + rflank.substr(0, 5), sub, indel, homop_mult);
      right = add_errors(r, rflank.substr(5) + rctx, sub, indel, homop_mult);
    } else {
      // true haplotype = window with the allele in place of the reference allele; against the reference it carries one
      // indel of |d| repeat bases right behind the left pad (left-aligned), everything else is aligned
      const std::string seg1 = lctx + lflank.substr(0, 30), seg2 = lflank.substr(30) + truth[a] + rflank.substr(0, 5),
                        seg3 = rflank.substr(5) + rctx;
      const int d = (truth_units[a] - ref_units) * period;  // > 0: insertion, < 0: deletion
      std::vector<uint8_t> al1(seg1.size(), 1), al2(seg2.size(), 1), al3(seg3.size(), 1);
      const size_t rep0 = 5 + lpad.size();  // first repeat base inside seg2
      if (d > 0)
        for (int k = 0; k < d; ++k) al2[rep0 + (size_t)k] = 0;
      CigarBuilder cg;
      left = add_errors_ops(r, seg1, al1, 0, sub, indel, homop_mult, cg);
      if (d >= 0) {
        mid = add_errors_ops(r, seg2, al2, 0, sub, indel, homop_mult, cg);
      } else {  // the deleted reference bases sit between the left pad and the first repeat base of the read
        const std::string head = seg2.substr(0, rep0), tail = seg2.substr(rep0);
        // (split only for the CIGAR: the generator is consumed base by base in the same order; the homopolymer context of
        // add_errors is evaluated on the whole segment, hence the offsets into seg2)
        CigarBuilder c_mid;
        mid = add_errors_ops(r, seg2, al2, 0, sub, indel, homop_mult, c_mid);
        // re-walk c_mid, inserting the deletion after the operations that consumed the first rep0 haplotype bases
        size_t consumed = 0;
        bool placed = false;
        for (uint32_t op : c_mid.ops) {
          uint32_t n = op >> 4;
          const uint32_t code = op & 15u;
          const bool uses_hap = (code == 7u || code == 8u || code == 2u);  // '=' 'X' 'D' consume an aligned haplotype base
          while (n > 0) {
            if (!placed && consumed == rep0) {
              cg.add(2u, (uint32_t)(-d));
              placed = true;
            }
            uint32_t take = n;
            if (!placed && uses_hap) take = (uint32_t)std::min<size_t>(n, rep0 - consumed);
            if (!placed && !uses_hap) take = n;  // error insertions before the boundary stay where they are
            cg.add(code, take);
            if (uses_hap) consumed += take;
            n -= take;
          }
        }
        if (!placed) cg.add(2u, (uint32_t)(-d));
        (void)head; (void)tail;
      }
      right = add_errors_ops(r, seg3, al3, 0, sub, indel, homop_mult, cg);
      // leading / trailing deletions are not part of an alignment: drop them and move the ends
      int32_t start = 1000 - 200;
      int32_t ref_len = 0;
      std::vector<uint32_t> ops = cg.ops;
      size_t b0 = 0, b1 = ops.size();
      while (b0 < b1 && (ops[b0] & 15u) == 2u) { start += (int32_t)(ops[b0] >> 4); ++b0; }
      while (b1 > b0 && (ops[b1 - 1] & 15u) == 2u) --b1;
      ops.assign(ops.begin() + (long)b0, ops.begin() + (long)b1);
      for (uint32_t op : ops)
        if ((op & 15u) == 7u || (op & 15u) == 8u || (op & 15u) == 2u) ref_len += (int32_t)(op >> 4);
      o.raw_reads.push_back(left + mid + right);
      o.raw_cigars.push_back(ops);
      o.raw_start.push_back(start);
      o.raw_stop.push_back(start + ref_len - 1);
    }
    const std::string key = left + "|" + mid + "|" + right;
    auto it = pools.find(key);
    uint32_t idx;
    if (it == pools.end()) {
      idx = (uint32_t)o.reads.size();
      pools.emplace(key, idx);
      o.reads.push_back(mid.empty() ? lflank.substr(30) + rflank.substr(0, 5) : mid);  // HapAligner.cpp:820-823
    } else {
      idx = it->second;
    }
    o.pool_index.push_back(idx);
    o.p1.push_back(a == 0 ? -0.000001 : -1000.0);
    o.p2.push_back(a == 1 ? -0.000001 : -1000.0);
  }
}

}  // namespace

extern "C" {



void ltr_synth_free(ltr_synth_batch* b) {
  if (!b) return;
  std::free((void*)b->vit.locus_hap_begin); std::free((void*)b->vit.locus_read_begin);
  std::free((void*)b->vit.hap_off); std::free((void*)b->vit.hap_bytes);
  std::free((void*)b->vit.read_off); std::free((void*)b->vit.read_bytes);
  std::free((void*)b->post.locus_sread_begin); std::free((void*)b->post.pool_index);
  std::free((void*)b->post.sample_label); std::free((void*)b->post.log_p1); std::free((void*)b->post.log_p2);
  std::free((void*)b->post.locus_n_samples);
  std::free(b);
}

// --alignment-params recommended for the configuration (ONT-like for config 4).
void ltr_synth_params(int config, ltr_params* p) {
  // Dindel defaults, reference HapAligner.h:118 (set here: this library does not link the product)
  p->ins_ins = -1.0f; p->ins_match = (float)-0.458675; p->del_del = -1.0f; p->del_match = (float)-0.458675;
  p->match_match = (float)-0.00005800168; p->match_ins = (float)-10.448214728; p->match_del = (float)-10.448214728;
  p->indel_flank_len = 5;
  if (config == 4) {
    p->ins_ins = -1.0f; p->ins_match = (float)-0.458675; p->del_del = -1.0f; p->del_match = (float)-0.458675;
    p->match_match = (float)-0.0202027; p->match_ins = (float)-4.60517; p->match_del = (float)-4.60517;
  }
}

int ltr_synth_generate(int config, uint64_t base_seed, uint32_t first_locus, uint32_t n_loci, int n_threads,
                       ltr_synth_batch** out) {
  if (!out || (config != 3 && config != 4 && config != 5)) return LTR_ERR_INVALID;
  std::vector<LocusOut> loci(n_loci);
  if (n_threads < 1) n_threads = 1;
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; ++t)
    th.emplace_back([&, t]() {
      for (uint32_t l = (uint32_t)t; l < n_loci; l += (uint32_t)n_threads)
        gen_locus(config, base_seed + first_locus + l, loci[l]);
    });
  for (auto& x : th) x.join();
  ltr_synth_batch* b = (ltr_synth_batch*)std::calloc(1, sizeof(ltr_synth_batch));
  uint64_t nh = 0, nr = 0, ns = 0, hb = 0, rb = 0;
  for (const LocusOut& o : loci) {
    nh += o.haps.size(); nr += o.reads.size(); ns += o.pool_index.size();
    for (const auto& s : o.haps) hb += s.size();
    for (const auto& s : o.reads) rb += s.size();
  }
  if (hb >= 0xFFFFFFF0ull || rb >= 0xFFFFFFF0ull) { std::free(b); return LTR_ERR_INVALID; }
  uint32_t* lhb = (uint32_t*)std::malloc(sizeof(uint32_t) * ((size_t)n_loci + 1));
  uint32_t* lrb = (uint32_t*)std::malloc(sizeof(uint32_t) * ((size_t)n_loci + 1));
  uint32_t* lsb = (uint32_t*)std::malloc(sizeof(uint32_t) * ((size_t)n_loci + 1));
  uint32_t* hoff = (uint32_t*)std::malloc(sizeof(uint32_t) * (nh + 1));
  uint32_t* roff = (uint32_t*)std::malloc(sizeof(uint32_t) * (nr + 1));
  uint8_t* hbytes = (uint8_t*)std::malloc(hb + 16);
  uint8_t* rbytes = (uint8_t*)std::malloc(rb + 16);
  uint32_t* pool = (uint32_t*)std::malloc(sizeof(uint32_t) * (ns + 1));
  int32_t* label = (int32_t*)std::calloc(ns + 1, sizeof(int32_t));
  double* p1 = (double*)std::malloc(sizeof(double) * (ns + 1));
  double* p2 = (double*)std::malloc(sizeof(double) * (ns + 1));
  uint32_t* nsamp = (uint32_t*)std::malloc(sizeof(uint32_t) * ((size_t)n_loci + 1));
  uint32_t ih = 0, ir = 0, is = 0;
  uint32_t ob = 0, orb = 0;
  lhb[0] = lrb[0] = lsb[0] = 0; hoff[0] = roff[0] = 0;
  for (uint32_t l = 0; l < n_loci; ++l) {
    const LocusOut& o = loci[l];
    for (const auto& s : o.haps) { std::memcpy(hbytes + ob, s.data(), s.size()); ob += (uint32_t)s.size(); hoff[++ih] = ob; }
    for (const auto& s : o.reads) { std::memcpy(rbytes + orb, s.data(), s.size()); orb += (uint32_t)s.size(); roff[++ir] = orb; }
    for (size_t i = 0; i < o.pool_index.size(); ++i, ++is) { pool[is] = o.pool_index[i]; p1[is] = o.p1[i]; p2[is] = o.p2[i]; }
    lhb[l + 1] = ih; lrb[l + 1] = ir; lsb[l + 1] = is; nsamp[l] = 1;
  }
  b->vit.n_loci = n_loci;
  b->vit.locus_hap_begin = lhb; b->vit.locus_read_begin = lrb; b->vit.hap_off = hoff; b->vit.hap_bytes = hbytes;
  b->vit.read_off = roff; b->vit.read_bytes = rbytes;
  b->post.locus_sread_begin = lsb; b->post.pool_index = pool; b->post.sample_label = label;
  b->post.log_p1 = p1; b->post.log_p2 = p2; b->post.locus_n_samples = nsamp; b->post.locus_haploid = nullptr;
  b->n_haps = ih; b->n_reads = ir; b->n_sreads = is; b->hap_nbytes = ob; b->read_nbytes = orb;
  *out = b;
  return LTR_OK;
}


// Raw form of the same workload: whole reads (+-200 bp around the repeat) with CIGARs, flank blocks, candidate alleles --
// what ltr_genotyper_run takes.  Same seeds, same read bases as ltr_synth_generate (the reads there are these reads pooled
// and cut to the repeat +- 5 bp).  One sample per locus.
int ltr_synth_generate_loci(int config, uint64_t base_seed, uint32_t first_locus, uint32_t n_loci, int n_threads,
                            ltr_synth_loci** out) {
  if (!out || (config != 3 && config != 4)) return LTR_ERR_INVALID;
  std::vector<LocusOut> loci(n_loci);
  if (n_threads < 1) n_threads = 1;
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; ++t)
    th.emplace_back([&, t]() {
      for (uint32_t l = (uint32_t)t; l < n_loci; l += (uint32_t)n_threads) {
        loci[l].want_raw = true;
        gen_locus(config, base_seed + first_locus + l, loci[l]);
      }
    });
  for (auto& x : th) x.join();
  uint64_t na = 0, nr = 0, ab = 0, rb = 0, nc = 0;
  for (const LocusOut& o : loci) {
    na += o.alleles.size(); nr += o.raw_reads.size();
    for (const auto& s : o.alleles) ab += s.size();
    for (const auto& s : o.raw_reads) rb += s.size();
    for (const auto& c : o.raw_cigars) nc += c.size();
  }
  if (ab >= 0xFFFFFFF0ull || rb >= 0xFFFFFFF0ull || nc >= 0xFFFFFFF0ull) return LTR_ERR_INVALID;
  ltr_synth_loci* S = (ltr_synth_loci*)std::calloc(1, sizeof(ltr_synth_loci));
  auto u32 = [](uint64_t n) { return (uint32_t*)std::malloc(sizeof(uint32_t) * (n + 1)); };
  auto i32 = [](uint64_t n) { return (int32_t*)std::calloc(n + 1, sizeof(int32_t)); };
  uint32_t *lfo = u32(n_loci), *rfo = u32(n_loci), *lab = u32(n_loci), *ao = u32(na), *lrb = u32(n_loci), *ro = u32(nr),
           *co = u32(nr), *cops = u32(nc), *nsamp = u32(n_loci);
  int32_t *rs = i32(n_loci), *re = i32(n_loci), *rstart = i32(nr), *rstop = i32(nr), *rsample = i32(nr);
  uint8_t* lfb = (uint8_t*)std::malloc((size_t)n_loci * 35 + 16);
  uint8_t* rfb = (uint8_t*)std::malloc((size_t)n_loci * 35 + 16);
  uint8_t* abytes = (uint8_t*)std::malloc(ab + 16);
  uint8_t* rbytes = (uint8_t*)std::malloc(rb + 16);
  double* p1 = (double*)std::malloc(sizeof(double) * (nr + 1));
  double* p2 = (double*)std::malloc(sizeof(double) * (nr + 1));
  lfo[0] = rfo[0] = lab[0] = ao[0] = lrb[0] = ro[0] = co[0] = 0;
  uint32_t ia = 0, ir = 0, oa = 0, orb = 0, oc = 0;
  for (uint32_t l = 0; l < n_loci; ++l) {
    const LocusOut& o = loci[l];
    std::memcpy(lfb + lfo[l], o.lflank.data(), o.lflank.size()); lfo[l + 1] = lfo[l] + (uint32_t)o.lflank.size();
    std::memcpy(rfb + rfo[l], o.rflank.data(), o.rflank.size()); rfo[l + 1] = rfo[l] + (uint32_t)o.rflank.size();
    for (const auto& s : o.alleles) { std::memcpy(abytes + oa, s.data(), s.size()); oa += (uint32_t)s.size(); ao[++ia] = oa; }
    for (size_t i = 0; i < o.raw_reads.size(); ++i) {
      const std::string& s = o.raw_reads[i];
      std::memcpy(rbytes + orb, s.data(), s.size()); orb += (uint32_t)s.size();
      for (uint32_t op : o.raw_cigars[i]) cops[oc++] = op;
      rstart[ir] = o.raw_start[i]; rstop[ir] = o.raw_stop[i]; p1[ir] = o.p1[i]; p2[ir] = o.p2[i];
      ++ir; ro[ir] = orb; co[ir] = oc;
    }
    lab[l + 1] = ia; lrb[l + 1] = ir; rs[l] = o.repeat_start; re[l] = o.repeat_end; nsamp[l] = 1;
  }
  ltr_locus_batch& B = S->batch;
  B.n_loci = n_loci;
  B.lflank_off = lfo; B.lflank_bytes = lfb; B.rflank_off = rfo; B.rflank_bytes = rfb;
  B.locus_allele_begin = lab; B.allele_off = ao; B.allele_bytes = abytes; B.repeat_start = rs; B.repeat_end = re;
  B.locus_read_begin = lrb; B.read_start = rstart; B.read_stop = rstop; B.read_off = ro; B.read_bytes = rbytes;
  B.cigar_off = co; B.cigar_ops = cops; B.read_sample = rsample; B.log_p1 = p1; B.log_p2 = p2; B.second_mate = nullptr;
  B.locus_n_samples = nsamp; B.locus_haploid = nullptr;
  S->n_alleles = ia; S->n_reads = ir; S->n_cigar_ops = oc; S->allele_nbytes = oa; S->read_nbytes = orb;
  *out = S;
  return LTR_OK;
}

void ltr_synth_loci_free(ltr_synth_loci* S) {
  if (!S) return;
  const ltr_locus_batch& B = S->batch;
  const void* ptrs[] = {B.lflank_off, B.lflank_bytes, B.rflank_off, B.rflank_bytes, B.locus_allele_begin, B.allele_off,
                        B.allele_bytes, B.repeat_start, B.repeat_end, B.locus_read_begin, B.read_start, B.read_stop, B.read_off,
                        B.read_bytes, B.cigar_off, B.cigar_ops, B.read_sample, B.log_p1, B.log_p2, B.locus_n_samples};
  for (const void* p : ptrs) std::free(const_cast<void*>(p));
  std::free(S);
}

}  // extern "C"
