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
};

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
  for (int i = 0; i < 30; ++i) {
    const int a = (int)(r.g() & 1);
    const std::string left = add_errors(r, lctx + lflank.substr(0, 30), sub, indel, homop_mult);
    const std::string mid = add_errors(r, lflank.substr(30) + truth[a] + rflank.substr(0, 5), sub, indel, homop_mult);
    const std::string right = add_errors(r, rflank.substr(5) + rctx, sub, indel, homop_mult);
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

}  // extern "C"
