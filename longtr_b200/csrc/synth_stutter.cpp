// synth_stutter.cpp -- BASELINE.json config 5: synthetic homopolymer loci for the --stutter-align-len path.
//
// homopolymer A or T, reference run 10-30 bp, true alleles ref + (-2..2) bases,
// 30 reads per locus spanning a +-200 bp window, HiFi errors (substitution 1e-3, indel 1e-3, x10 inside the run),
// qualities Phred 20-40, CIGAR in '=XID' against the reference window.  PRNG = mt19937_64(base_seed + locus).
// Reads are pooled the way LongTR does (ReadPooler: identical sequence -> one pool, first read's coordinates and
// CIGAR, per-position median quality; src/read_pooler.cpp:3-20) and every pooled read gets the seed base of
// HapAligner::calc_seed_base (HapAligner.cpp:493-542), both through the host mirror (csrc/host).  The result is an
// ltr_stutter_batch plus the per-read coordinates / CIGARs, so the very same loci can be replayed as flat loci
// (include/longtr_b200_locus.h) through the reference on the CPU.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "host/longtr_host.h"
#include "longtr_b200.h"

namespace {

const char kBases[5] = "ACGT";

struct Rng {
  std::mt19937_64 g;
  explicit Rng(uint64_t s) : g(s) {}
  uint32_t below(uint32_t n) { return (uint32_t)(g() % n); }
  int range(int lo, int hi) { return lo + (int)below((uint32_t)(hi - lo + 1)); }
  double unif() { return (double)(g() >> 11) * (1.0 / 9007199254740992.0); }
  char base() { return kBases[g() & 3]; }
  std::string seq(int n) {
    std::string s((size_t)n, 'A');
    for (int i = 0; i < n; ++i) s[i] = base();
    return s;
  }
};

struct ReadOut {
  std::string seq, qual, cigar;
  int32_t start, stop;
};

struct LocusOut {
  std::string lflank, rflank;
  std::vector<std::string> alleles;
  int32_t repeat_start, repeat_end;
  std::vector<ReadOut> pooled;
  std::vector<int32_t> seeds;
  std::vector<uint32_t> pool_index;
  std::vector<double> p1, p2;
};

void append_op(std::string& cigar, char& cur, int& cnt, char op) {
  if (op == cur) {
    cnt++;
    return;
  }
  if (cnt > 0) cigar += std::to_string(cnt) + cur;
  cur = op;
  cnt = 1;
}

// One read of haplotype = window with the reference run replaced by a run of `run_len` bases.
ReadOut simulate_read(Rng& r, const std::string& left, char hb, int ref_run, int run_len, const std::string& right,
                      int32_t win_start) {
  // events against the reference window: (op, base)
  std::vector<std::pair<char, char> > ev;
  for (char c : left) ev.push_back(std::make_pair('=', c));
  const int shared = std::min(ref_run, run_len);
  for (int i = 0; i < shared; ++i) ev.push_back(std::make_pair('=', hb));
  for (int i = shared; i < run_len; ++i) ev.push_back(std::make_pair('I', hb));
  for (int i = shared; i < ref_run; ++i) ev.push_back(std::make_pair('D', '-'));
  for (char c : right) ev.push_back(std::make_pair('=', c));
  const int run_lo = (int)left.size(), run_hi = run_lo + std::max(ref_run, run_len);
  ReadOut out;
  std::string cigar;
  char cur = 0;
  int cnt = 0, n_ref = 0;
  const int n_ev = (int)ev.size();
  for (int k = 0; k < n_ev; ++k) {
    const char op = ev[k].first, c = ev[k].second;
    const bool edge = k < 3 || k >= n_ev - 3;
    if (op == '=' && !edge) {
      const double ind = (k >= run_lo && k < run_hi) ? 1e-2 : 1e-3;
      const double u = r.unif();
      if (u < 1e-3) {
        out.seq.push_back(kBases[(uint32_t)(std::strchr(kBases, c) - kBases + 1 + r.below(3)) & 3]);
        append_op(cigar, cur, cnt, 'X');
        n_ref++;
        continue;
      }
      if (u < 1e-3 + ind * 0.5) {
        append_op(cigar, cur, cnt, 'D');
        n_ref++;
        continue;
      }
      if (u < 1e-3 + ind) {
        out.seq.push_back(c);
        append_op(cigar, cur, cnt, '=');
        n_ref++;
        out.seq.push_back(r.base());
        append_op(cigar, cur, cnt, 'I');
        continue;
      }
    }
    if (op == 'D') {
      append_op(cigar, cur, cnt, 'D');
      n_ref++;
    } else {
      out.seq.push_back(c);
      append_op(cigar, cur, cnt, op);
      if (op == '=') n_ref++;
    }
  }
  if (cnt > 0) cigar += std::to_string(cnt) + cur;
  out.cigar = cigar;
  out.qual.resize(out.seq.size());
  for (size_t i = 0; i < out.seq.size(); ++i) out.qual[i] = (char)(33 + r.range(20, 40));
  out.start = win_start;
  out.stop = win_start + n_ref - 1;
  return out;
}

void gen_locus(uint64_t seed, LocusOut& o) {
  using namespace ltr;
  Rng r(seed);
  const char hb = (r.g() & 1) ? 'A' : 'T';
  const int ref_run = r.range(10, 30);
  const std::string lctx = r.seq(165), rctx = r.seq(165);
  o.lflank = r.seq(35);
  o.rflank = r.seq(35);
  std::string lpad = r.seq(5), rpad = r.seq(5);
  if (lpad[4] == hb) lpad[4] = (hb == 'A') ? 'C' : 'G';  // keep the run length well defined
  if (rpad[0] == hb) rpad[0] = (hb == 'A') ? 'C' : 'G';
  const int k1 = r.range(-2, 2), k2 = r.range(-2, 2);
  const int truth[2] = {ref_run + k1, ref_run + k2};
  auto allele_of = [&](int run) { return lpad + std::string((size_t)run, hb) + rpad; };
  const std::string ref_allele = allele_of(ref_run);
  std::vector<std::string> alts;
  auto add_alt = [&](const std::string& s) {
    if (s != ref_allele && std::find(alts.begin(), alts.end(), s) == alts.end()) alts.push_back(s);
  };
  add_alt(allele_of(truth[0]));
  add_alt(allele_of(truth[1]));
  const int want_h = r.range(2, 5);
  int guard = 0;
  while ((int)alts.size() + 1 < want_h && guard++ < 32) add_alt(allele_of(ref_run + r.range(-3, 3)));
  std::sort(alts.begin(), alts.end(), [](const std::string& a, const std::string& b) {
    return a.size() != b.size() ? a.size() < b.size() : a < b;  // HaplotypeGenerator.cpp:475
  });
  o.alleles.clear();
  o.alleles.push_back(ref_allele);
  for (const std::string& a : alts) o.alleles.push_back(a);
  const int32_t pos0 = 10000;
  o.repeat_start = pos0;
  o.repeat_end = pos0 + (int32_t)ref_allele.size();
  const int32_t win_start = pos0 - 35 - 165;

  ReadPooler pooler;
  o.pool_index.clear();
  o.p1.clear();
  o.p2.clear();
  for (int i = 0; i < 30; ++i) {
    const int a = (int)(r.g() & 1);
    ReadOut rd = simulate_read(r, lctx + o.lflank + lpad, hb, ref_run, truth[a], rpad + o.rflank + rctx, win_start);
    Alignment aln(rd.start, rd.stop, false, false, "read", rd.qual, rd.seq, rd.seq);
    aln.set_cigar_string(rd.cigar.c_str());
    o.pool_index.push_back((uint32_t)pooler.add_alignment(aln));
    o.p1.push_back(a == 0 ? -0.000001 : -1000.0);
    o.p2.push_back(a == 1 ? -0.000001 : -1000.0);
  }
  BaseQuality bq;
  pooler.pool(bq);
  // seeds through the host mirror of HapAligner
  StutterModel model(0.95, 0.05, 0.05, 0.95, 0.01, 0.01, std::string(1, hb));
  HapBlock left(o.repeat_start - 35, o.repeat_start, o.lflank), right(o.repeat_end, o.repeat_end + 35, o.rflank);
  RepeatBlock rep(o.repeat_start, o.repeat_end, o.alleles[0], 1, &model);
  for (size_t k = 1; k < o.alleles.size(); ++k) rep.add_alternate(std::make_pair(o.alleles[k], false));
  std::vector<HapBlock*> blocks;
  blocks.push_back(&left);
  blocks.push_back(&rep);
  blocks.push_back(&right);
  Haplotype hap(blocks);
  std::vector<bool> all((size_t)hap.num_combs(), true);
  std::vector<float> none;
  HapAligner aligner(&hap, all, 5, 20, none, NULL);
  o.pooled.clear();
  o.seeds.clear();
  for (const Alignment& p : pooler.get_alignments()) {
    ReadOut rd;
    rd.seq = p.get_sequence();
    rd.qual = p.get_base_qualities();
    rd.cigar = p.getCigarString();
    rd.start = p.get_start();
    rd.stop = p.get_stop();
    o.pooled.push_back(rd);
    o.seeds.push_back(aligner.calc_seed_base(p));
  }
}

template <typename T>
T* dup(const std::vector<T>& v) {
  T* p = (T*)std::malloc(std::max<size_t>(1, v.size()) * sizeof(T));
  if (!v.empty()) std::memcpy(p, v.data(), v.size() * sizeof(T));
  return p;
}
uint8_t* dups(const std::string& s) {
  uint8_t* p = (uint8_t*)std::malloc(std::max<size_t>(1, s.size()));
  if (!s.empty()) std::memcpy(p, s.data(), s.size());
  return p;
}

}  // namespace

extern "C" {

int ltr_synth_stutter_generate(uint64_t base_seed, uint32_t first_locus, uint32_t n_loci, int n_threads,
                               ltr_synth_stutter_batch** out) {
  if (!out) return LTR_ERR_INVALID;
  std::vector<LocusOut> loci(n_loci);
  if (n_threads < 1) n_threads = 1;
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; ++t)
    th.emplace_back([&, t]() {
      for (uint32_t l = (uint32_t)t; l < n_loci; l += (uint32_t)n_threads) gen_locus(base_seed + first_locus + l, loci[l]);
    });
  for (auto& x : th) x.join();
  std::vector<uint32_t> lab(1, 0), lrb(1, 0), lfo(1, 0), rfo(1, 0), alo(1, 0), rdo(1, 0), cgo(1, 0), lsb(1, 0), pool, nsamp;
  std::vector<int32_t> seeds, motif_len, label, rstart, rstop, rep_start, rep_end;
  std::vector<double> stutter, p1, p2;
  std::string lf, rf, al, rd, ql, cg;
  const double st[6] = {0.95, 0.05, 0.05, 0.95, 0.01, 0.01};  // hipstr_main.cpp:362-363 fixed model
  for (const LocusOut& o : loci) {
    lf += o.lflank; lfo.push_back((uint32_t)lf.size());
    rf += o.rflank; rfo.push_back((uint32_t)rf.size());
    for (const std::string& a : o.alleles) { al += a; alo.push_back((uint32_t)al.size()); }
    lab.push_back((uint32_t)alo.size() - 1);
    for (size_t i = 0; i < o.pooled.size(); ++i) {
      rd += o.pooled[i].seq; ql += o.pooled[i].qual; rdo.push_back((uint32_t)rd.size());
      cg += o.pooled[i].cigar; cgo.push_back((uint32_t)cg.size());
      rstart.push_back(o.pooled[i].start); rstop.push_back(o.pooled[i].stop);
      seeds.push_back(o.seeds[i]);
    }
    lrb.push_back((uint32_t)rdo.size() - 1);
    stutter.insert(stutter.end(), st, st + 6);
    motif_len.push_back(1);
    rep_start.push_back(o.repeat_start); rep_end.push_back(o.repeat_end);
    pool.insert(pool.end(), o.pool_index.begin(), o.pool_index.end());
    p1.insert(p1.end(), o.p1.begin(), o.p1.end());
    p2.insert(p2.end(), o.p2.begin(), o.p2.end());
    label.insert(label.end(), o.pool_index.size(), 0);
    lsb.push_back((uint32_t)pool.size());
    nsamp.push_back(1);
  }
  ltr_synth_stutter_batch* b = (ltr_synth_stutter_batch*)std::calloc(1, sizeof(ltr_synth_stutter_batch));
  b->batch.n_loci = n_loci;
  b->batch.locus_allele_begin = dup(lab); b->batch.locus_read_begin = dup(lrb);
  b->batch.lflank_off = dup(lfo); b->batch.lflank_bytes = dups(lf);
  b->batch.rflank_off = dup(rfo); b->batch.rflank_bytes = dups(rf);
  b->batch.allele_off = dup(alo); b->batch.allele_bytes = dups(al);
  b->batch.stutter = dup(stutter); b->batch.motif_len = dup(motif_len);
  b->batch.read_off = dup(rdo); b->batch.read_bytes = dups(rd); b->batch.qual_bytes = dups(ql);
  b->batch.read_seed = dup(seeds);
  b->batch.realign_allele = NULL; b->batch.realign_read = NULL;
  b->post.locus_sread_begin = dup(lsb); b->post.pool_index = dup(pool); b->post.sample_label = dup(label);
  b->post.log_p1 = dup(p1); b->post.log_p2 = dup(p2); b->post.locus_n_samples = dup(nsamp); b->post.locus_haploid = NULL;
  b->read_start = dup(rstart); b->read_stop = dup(rstop); b->cigar_off = dup(cgo); b->cigar_bytes = dups(cg);
  b->repeat_start = dup(rep_start); b->repeat_end = dup(rep_end);
  b->n_alleles = (uint32_t)alo.size() - 1; b->n_reads = (uint32_t)rdo.size() - 1; b->n_sreads = (uint32_t)pool.size();
  *out = b;
  return LTR_OK;
}

void ltr_synth_stutter_free(ltr_synth_stutter_batch* b) {
  if (!b) return;
  const void* ptrs[] = {b->batch.locus_allele_begin, b->batch.locus_read_begin, b->batch.lflank_off, b->batch.lflank_bytes,
                        b->batch.rflank_off, b->batch.rflank_bytes, b->batch.allele_off, b->batch.allele_bytes,
                        b->batch.stutter, b->batch.motif_len, b->batch.read_off, b->batch.read_bytes, b->batch.qual_bytes,
                        b->batch.read_seed, b->post.locus_sread_begin, b->post.pool_index, b->post.sample_label,
                        b->post.log_p1, b->post.log_p2, b->post.locus_n_samples, b->read_start, b->read_stop,
                        b->cigar_off, b->cigar_bytes, b->repeat_start, b->repeat_end};
  for (const void* p : ptrs) std::free((void*)p);
  std::free(b);
}

}  // extern "C"
