// TEST INFRASTRUCTURE ONLY.  Host-side emulation of one warp of the Viterbi stream kernel:
// the very same lane functions as the CUDA kernel (longtr_b200/csrc/viterbi_core.cuh,
// compiled with LTR_HOST_EMU), with the warp shuffles replaced by array copies.  Lets the
// CPU test-suite check the wavefront / boundary / strip / witness logic against the oracle
// without a GPU.  It is not a CPU fallback: nothing in the product links it.
#define LTR_HOST_EMU 1
#include "viterbi_host.h"

#include <array>
#include <cstdio>
#include <vector>

using namespace ltr;

struct EmuScratch {
  std::vector<XY> sxy;
  std::vector<uint32_t> sb;
  std::vector<unsigned char> smem;
};

template <int K, int MODE>
static void emu_task(const VitConsts& C, const DevBatch& B, const Task& T, FailSink fail, EmuScratch& E) {
  const uint32_t g = T.hap;
  const uint32_t hoff = B.hap_off[g];
  const int32_t hlen = (int32_t)(B.hap_off[g + 1] - hoff);
  const uint32_t l = B.hap_locus[g];
  const uint32_t H = B.locus_hap_begin[l + 1] - B.locus_hap_begin[l];
  const uint32_t rb0 = B.locus_read_begin[l];
  const unsigned long long out_base = B.ll_off[l] + (g - B.locus_hap_begin[l]);
  const int32_t n = hlen - 2 * C.cut;
  if (hlen <= 60 || n < 1) {
    for (uint32_t p = T.read_begin; p < T.read_end; ++p)
      B.out_ll[out_base + (unsigned long long)(p - rb0) * H] = C.imp;
    return;
  }
  const uint8_t* hap = B.hap_bytes + hoff + C.cut;
  if (n == 1) {
    for (uint32_t p = T.read_begin; p < T.read_end; ++p) {
      const uint32_t qb = B.read_off[p];
      const int32_t m = (int32_t)(B.read_off[p + 1] - qb);
      double v;
      if (std::abs(n - m) > 600) v = -700.0;
      else v = single_row_result(C, m, (m - 1 < n) ? hap[m - 1] : 0, B.read_bytes[qb], hap[0]);
      B.out_ll[out_base + (unsigned long long)(p - rb0) * H] = v;
    }
    return;
  }
  StripCtx S;
  S.hap = hap; S.read_bytes = B.read_bytes; S.read_off = B.read_off; S.out_ll = B.out_ll;
  S.out_base = out_base; S.H = H; S.rb0 = rb0; S.hap_index = g;
  S.qs = B.read_off[T.read_begin];
  S.Q = B.read_off[T.read_end] - S.qs;
  S.n = n; S.h0 = hap[0];
  S.fail = fail;
  E.sxy.assign(S.Q + 128, XY());
  E.sb.assign(S.Q + 128, 0);
  E.smem.assign(warp_smem_bytes(K), 0);
  S.sxy = E.sxy.data(); S.sb = E.sb.data();
  S.bnd = reinterpret_cast<XY*>(E.smem.data());
  S.tx = reinterpret_cast<double*>(E.smem.data() + 2 * 32 * sizeof(XY));
  S.tz = S.tx + 2 * K * 32;
  S.txo = S.tz + 2 * K * 32;
  const StripPlan P = plan_strips(n - 1, K);
  int32_t row_start = 1;
  std::vector<LaneStream<K>> lanes(32);
  for (int s = 0; s < P.strips; ++s) {
    const int32_t rows = P.base + (s < P.rem ? 1 : 0);
    S.first_strip = (s == 0);
    S.last_strip = (s == P.strips - 1);
    for (int t = 0; t < 32; ++t) {
      int32_t i0, nrows, tl;
      lane_geometry(K, t, rows, row_start, i0, nrows, tl);
      S.t_last = tl;
      lane_stream_reset<K>(lanes[t], C, S, t, i0, nrows, T.read_begin);
    }
    std::vector<XY> nxt(32);
    std::vector<BoundaryCursor> bc(32);
    for (int t = 0; t < 32; ++t) {
      boundary_cursor_reset(bc[t], S, T.read_begin);
      if (s == 0) { S.bnd[t] = boundary_at(C, S, bc[t], (uint32_t)t); nxt[t] = boundary_at(C, S, bc[t], 32u + (uint32_t)t); }
      else { S.bnd[t] = S.sxy[t]; nxt[t] = S.sxy[32 + t]; }
    }
    std::vector<double> ox(32), oy(32);
    std::vector<uint32_t> ob(32);
    const uint32_t nsteps = S.Q + (uint32_t)S.t_last;
    uint32_t step = 0;
    while (step < nsteps) {  // same event-driven schedule as viterbi_stream_kernel
      const uint32_t chunk_end = (step + 32u < nsteps) ? step + 32u : nsteps;
      while (step < chunk_end) {
        uint32_t nfast = 0xFFFFFFFFu;
        for (int t = 0; t < 32; ++t) nfast = std::min(nfast, lane_plain_distance<K>(lanes[t], S, step - (uint32_t)t));
        nfast = std::min(nfast, chunk_end - step);
        for (int t = 0; t < 32; ++t) { ox[t] = lanes[t].L.Xout; oy[t] = lanes[t].L.Yout; ob[t] = lanes[t].L.Bout; }
        if (nfast > 0u) {
          for (int t = 0; t < 32; ++t)
            lane_fast_step<K, MODE>(lanes[t], C, S, t, step - (uint32_t)t, t ? ox[t - 1] : 0.0, t ? oy[t - 1] : 0.0);
        } else {
          for (int t = 0; t < 32; ++t) {
            const uint32_t pos = step - (uint32_t)t;
            if (pos < S.Q)
              lane_stream_step<K, MODE>(lanes[t], C, S, t, pos, t ? ox[t - 1] : 0.0, t ? oy[t - 1] : 0.0, t ? ob[t - 1] : 0u);
          }
        }
        ++step;
      }
      if (step < nsteps)
        for (int t = 0; t < 32; ++t) {
          S.bnd[((step >> 5) & 1u) * 32u + t] = nxt[t];
          nxt[t] = (s == 0) ? boundary_at(C, S, bc[t], step + 32u + (uint32_t)t) : S.sxy[step + 32u + t];
        }
    }
    row_start += rows;
  }
}

template <int MODE>
static void emu_dispatch(int k, const VitConsts& C, const DevBatch& B, const Task& T, FailSink fail,
                         EmuScratch& E) {
  switch (k) {
#define CASE(KK) case KK: emu_task<KK, MODE>(C, B, T, fail, E); break;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
    CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14) CASE(15) CASE(16)
#undef CASE
    default: std::fprintf(stderr, "emu: bad K %d\n", k); std::abort();
  }
}

// Study hook (tools only): when set, every banded pair appends (n, m, W, w, F_band, certified).
static std::vector<double>* g_band_dump = nullptr;
extern "C" void ltr_emu_band_dump_begin() { delete g_band_dump; g_band_dump = new std::vector<double>(); }
extern "C" uint64_t ltr_emu_band_dump_size() { return g_band_dump ? g_band_dump->size() / 6 : 0; }
extern "C" void ltr_emu_band_dump_get(double* out) {
  if (g_band_dump) std::memcpy(out, g_band_dump->data(), g_band_dump->size() * sizeof(double));
}

// ---- band kernel (band_core.cuh): one round = four pairs in lock step, eight lanes per pair --------------------------
struct EmuTable {
  const double* t;  // [2K][3]
  double x(int q) const { return t[3 * q]; }
  double y(int q) const { return t[3 * q + 1]; }
  double z(int q) const { return t[3 * q + 2]; }
};

template <int K, int G, bool SYM>
static void emu_band_round(const VitConsts& C, const DevBatch& B, const uint32_t (*pairs)[2], uint32_t n_pairs,
                           uint32_t base, const BandGap& gap, uint64_t* n_uncert) {
  constexpr int W = 2 * K * G, NG = 32 / G;  // NG pairs per round
  BandPair R[NG][G];
  BandLane<K> L[NG][G];
  double tab[NG][G][6 * K];
  BandGeom geo[NG];
  bool active[NG];
  double* out[NG];
  int32_t s_pro = 0, s_end_min = 0x7FFFFFFF, s_end_max = 0;
  for (int g = 0; g < NG; ++g) {
    active[g] = base + g < n_pairs;
    const uint32_t pi = active[g] ? base + g : base;
    const uint32_t h = pairs[pi][0], u = pairs[pi][1];
    const uint32_t hoff = B.hap_off[h];
    const int32_t hlen = (int32_t)(B.hap_off[h + 1] - hoff);
    const uint32_t qb = B.read_off[u];
    const int32_t n = hlen - 2 * C.cut, m = (int32_t)(B.read_off[u + 1] - qb);
    geo[g] = band_geometry(n, m, W);
    const uint32_t l = B.hap_locus[h];
    const uint32_t hb0 = B.locus_hap_begin[l], H = B.locus_hap_begin[l + 1] - hb0;
    out[g] = B.out_ll + B.ll_off[l] + (unsigned long long)(u - B.locus_read_begin[l]) * H + (h - hb0);
    for (int t = 0; t < G; ++t) {
      R[g][t].n = n; R[g][t].m = m;
      R[g][t].hap = B.hap_bytes + hoff + C.cut;
      R[g][t].read = B.read_bytes + qb;
      R[g][t].d0 = geo[g].dlo + 2 * K * t;
      for (int q = 0; q < 2 * K; ++q) band_boundary(C, R[g][t], R[g][t].d0 + q, tab[g][t][3 * q], tab[g][t][3 * q + 1], tab[g][t][3 * q + 2]);
      band_lane_reset<K>(L[g][t], C);
    }
    s_pro = std::max(s_pro, band_prologue_steps(geo[g].dlo, W));
    s_end_min = std::min(s_end_min, n + m - 2);
    s_end_max = std::max(s_end_max, n + m - 2);
  }
  double F[NG][G];
  bool got[NG][G];
  for (int g = 0; g < NG; ++g) for (int t = 0; t < G; ++t) { F[g][t] = C.imp; got[g][t] = false; }
  int32_t s = 0;
  auto general = [&]() {
    for (int g = 0; g < NG; ++g) {
      double nb[G];
      for (int t = 0; t < G; ++t) {
        if (s & 1) nb[t] = (t == G - 1) ? C.imp : L[g][t + 1].A[0];
        else nb[t] = (t == 0) ? C.imp : L[g][t - 1].B[K - 1];
      }
      for (int t = 0; t < G; ++t) {
        EmuTable T{tab[g][t]};
        if (s & 1) band_general_step<K, 1, SYM>(L[g][t], C, R[g][t], T, s, nb[t], F[g][t], got[g][t]);
        else band_general_step<K, 0, SYM>(L[g][t], C, R[g][t], T, s, nb[t], F[g][t], got[g][t]);
      }
    }
    ++s;
  };
  if (s + 1 < s_end_min) {  // same schedule as viterbi_band_kernel: plain double steps, fix-ups during the prologue
    int32_t hi[NG][G], ri[NG][G];
    for (int g = 0; g < NG; ++g)
      for (int t = 0; t < G; ++t) {
        band_windows_init<K>(L[g][t], R[g][t], s);
        hi[g][t] = ((s - R[g][t].d0) >> 1) + 1;
        ri[g][t] = ((s + R[g][t].d0) >> 1) + K + 1;
      }
    for (; s + 1 < s_end_min; s += 2) {
      const bool prologue = s < s_pro;
      for (int g = 0; g < NG; ++g) {
        double nb[G];
        int32_t nh[G], nr[G];
        for (int t = 0; t < G; ++t) {
          // the device reads up to ~W/2 + K bytes past and up to W/2 bytes in front of the strings (padded buffers);
          // those values only reach cells outside the matrix
          nh[t] = (int32_t)R[g][t].hap[hi[g][t]++];
          nr[t] = (int32_t)R[g][t].read[ri[g][t]++];
          nb[t] = (t == 0) ? C.imp : L[g][t - 1].B[K - 1];
        }
        for (int t = 0; t < G; ++t) {
          band_fast_even<K, SYM>(L[g][t], C, nb[t]);
          if (prologue) band_fixup<K, 0>(L[g][t], R[g][t], EmuTable{tab[g][t]}, s);
        }
        for (int t = 0; t < G; ++t) nb[t] = (t == G - 1) ? C.imp : L[g][t + 1].A[0];
        for (int t = 0; t < G; ++t) {
          band_fast_odd<K, SYM>(L[g][t], C, nb[t], nh[t], nr[t]);
          if (prologue) band_fixup<K, 1>(L[g][t], R[g][t], EmuTable{tab[g][t]}, s + 1);
        }
      }
    }
  }
  while (s <= s_end_max) general();
  for (int g = 0; g < NG; ++g) {
    if (!active[g]) continue;
    int owners = 0;
    for (int t = 0; t < G; ++t) {
      if (!got[g][t]) continue;
      ++owners;
      const double thr = band_threshold(C, gap, R[g][t].n, R[g][t].m, geo[g].w);
      const bool ok = F[g][t] > thr;
      if (g_band_dump) {
        const double rec[6] = {(double)R[g][t].n, (double)R[g][t].m, (double)W, (double)geo[g].w, F[g][t], ok ? 1.0 : 0.0};
        g_band_dump->insert(g_band_dump->end(), rec, rec + 6);
      }
      *out[g] = ok ? F[g][t] : band_mark(F[g][t]);
      if (!ok) ++*n_uncert;
    }
    if (owners != 1) { std::fprintf(stderr, "emu band: %d owners\n", owners); std::abort(); }
  }
}

static void emu_band_dispatch(int cls, const VitConsts& C, const DevBatch& B, const uint32_t (*pairs)[2], uint32_t n_pairs,
                              uint32_t base, const BandGap& gap, uint64_t* n_uncert) {
  const bool sym = (C.d2m == C.i2m) && (C.m2i == C.m2d);  // as launch_band
#define EMU_BAND(KK, GG)                                                                 \
  do {                                                                                   \
    if (sym) emu_band_round<KK, GG, true>(C, B, pairs, n_pairs, base, gap, n_uncert);    \
    else emu_band_round<KK, GG, false>(C, B, pairs, n_pairs, base, gap, n_uncert);       \
  } while (0)
  switch (cls) {  // as band_kernel_for
    case 0: EMU_BAND(4, 4); break;
    case 1: EMU_BAND(3, 8); break;
    case 2: EMU_BAND(4, 8); break;
    case 3: EMU_BAND(6, 8); break;
    case 4: EMU_BAND(8, 8); break;
    case 5: EMU_BAND(6, 16); break;
    case 6: EMU_BAND(4, 32); break;
    case 7: EMU_BAND(5, 32); break;
    case 8: EMU_BAND(6, 32); break;
    case 9: EMU_BAND(7, 32); break;
    case 10: EMU_BAND(8, 32); break;
    default: std::abort();
  }
#undef EMU_BAND
}

extern "C" int ltr_emu_viterbi_batch_band(const ltr_viterbi_batch* b, const ltr_params* p, int kmax, int use_fast,
                                          int band_w, double* out_ll, uint64_t* n_fallback, uint64_t* band_stats);
static uint64_t g_band_retried = 0;  // pairs that went through the second band round since the last reset
extern "C" uint64_t ltr_emu_band_retried(int reset) {
  const uint64_t v = g_band_retried;
  if (reset) g_band_retried = 0;
  return v;
}

extern "C" int ltr_emu_viterbi_batch(const ltr_viterbi_batch* b, const ltr_params* p, int kmax,
                                     int use_fast, double* out_ll, uint64_t* n_fallback) {
  return ltr_emu_viterbi_batch_band(b, p, kmax, use_fast, -1, out_ll, n_fallback, nullptr);
}

// band_w as ltr_ctx_set_band; band_stats (may be NULL): [0] pairs sent to the band kernel, [1] of which uncertified,
// [2] stream-kernel tasks created for them.
extern "C" int ltr_emu_viterbi_batch_band(const ltr_viterbi_batch* b, const ltr_params* p, int kmax, int use_fast,
                                          int band_w, double* out_ll, uint64_t* n_fallback, uint64_t* band_stats) {
  Plan plan;
  int rc = make_plan(*b, *p, kmax, plan, 0, nullptr, nullptr, use_fast ? band_w : -1);
  if (rc != LTR_OK) return rc;
  HostConsts hc;
  make_consts(*p, std::max(plan.max_n, plan.max_m) + 2, hc);
  hc.C.tabI = hc.tabI.data();
  hc.C.tabD = hc.tabD.data();
  // the warps see the distinct trimmed reads of each locus (Plan); padded copy: the kernel prefetches one byte ahead
  const size_t kPad = 512;  // as the device buffers: readable bytes on both sides of the strings
  std::vector<uint8_t> rbytes(kPad + plan.uread_nbytes + kPad, 0);
  if (plan.uread_nbytes) std::memcpy(rbytes.data() + kPad, plan.uread_bytes, plan.uread_nbytes);
  std::vector<double> uniq_ll((size_t)plan.ull_off[b->n_loci] + 1, 123.0);
  DevBatch B;
  B.hap_bytes = b->hap_bytes; B.hap_off = b->hap_off; B.hap_locus = plan.hap_locus.data();
  B.read_bytes = rbytes.data() + kPad; B.read_off = plan.uread_off.data();
  B.locus_hap_begin = b->locus_hap_begin; B.locus_read_begin = plan.locus_uread_begin.data();
  B.ll_off = plan.ull_off.data(); B.out_ll = uniq_ll.data();
  EmuScratch E;
  uint64_t nfall = 0;
  if (!fast_certificate_valid(*p)) use_fast = 0;  // as ltr_job_create does
  // ---- band kernel first; what it cannot certify goes to the stream kernel's task lists (band_collect_kernel) ------
  std::vector<uint8_t> hbytes(kPad + (size_t)b->hap_off[b->locus_hap_begin[b->n_loci]] + kPad, 0);  // padded like the device copy
  std::memcpy(hbytes.data() + kPad, b->hap_bytes, hbytes.size() - 2 * kPad);
  B.hap_bytes = hbytes.data() + kPad;
  uint64_t bstats[3] = {plan.n_band_pairs, 0, 0};
  static const int retry_rho = [] { const char* e = getenv("LTR_BAND_RETRY_RHO"); return e ? atoi(e) : 80; }();  // as abi.cu
  std::vector<std::array<uint32_t, 2>> retry[kBandClasses];
  for (int c = 0; c < kBandClasses; ++c) {
    std::vector<std::array<uint32_t, 2>> pairs;
    for (const BandTask& bt : plan.band_tasks[(size_t)c])
      for (uint32_t r = bt.read_begin; r < bt.read_end; ++r) pairs.push_back({bt.hap, r});
    const uint32_t np = (uint32_t)pairs.size();
    for (uint32_t base = 0; base < np; base += 32u / (uint32_t)band_class_g(c))
      emu_band_dispatch(c, hc.C, B, reinterpret_cast<const uint32_t(*)[2]>(pairs.data()), np, base,
                        plan.band.gap, &bstats[1]);
    for (const BandTask& bt : plan.band_tasks[(size_t)c]) {  // as band_collect_kernel
      const uint32_t l = plan.hap_locus[bt.hap];
      const uint32_t hb0 = b->locus_hap_begin[l], H = b->locus_hap_begin[l + 1] - hb0, rb0 = plan.locus_uread_begin[l];
      const int n = (int)(b->hap_off[bt.hap + 1] - b->hap_off[bt.hap]) - 2 * hc.C.cut;
      const int kr = rows_per_lane_hd(n, kmax);
      const double* col = uniq_ll.data() + plan.ull_off[l] + (bt.hap - hb0);
      bool in_run = false;
      uint32_t run_begin = 0;
      for (uint32_t r = bt.read_begin; r <= bt.read_end; ++r) {
        bool bad = false;
        if (r < bt.read_end) {
          const double v = col[(size_t)(r - rb0) * H];
          bad = band_marked(v);
          if (bad && retry_rho > 0) {
            const int rc = band_retry_class(hc.C, plan.band.gap, n, (int32_t)(plan.uread_off[r + 1] - plan.uread_off[r]),
                                            band_unmark(v), retry_rho);
            if (rc >= 0) {
              if (rc <= c) { std::fprintf(stderr, "emu band: retry class %d not wider than %d\n", rc, c); std::abort(); }
              bad = false;
              retry[rc].push_back({bt.hap, r});
            }
          }
        }
        if (bad && !in_run) { in_run = true; run_begin = r; }
        if (!bad && in_run) {
          in_run = false;
          Task T; T.hap = bt.hap; T.read_begin = run_begin; T.read_end = r;
          plan.tasks[(size_t)kr].push_back(T);
          ++bstats[2];
        }
      }
    }
  }
  // second band round + band_retry_check_kernel: every retried pair must come back certified
  for (int c = 1; c < kBandClasses; ++c) {
    const uint32_t np = (uint32_t)retry[c].size();
    uint64_t still = 0;
    for (uint32_t base = 0; base < np; base += 32u / (uint32_t)band_class_g(c))
      emu_band_dispatch(c, hc.C, B, reinterpret_cast<const uint32_t(*)[2]>(retry[c].data()), np, base, plan.band.gap, &still);
    if (still) { std::fprintf(stderr, "emu band: %llu retried pairs not certified\n", (unsigned long long)still); std::abort(); }
    g_band_retried += np;
  }
  if (band_stats) std::memcpy(band_stats, bstats, sizeof(bstats));
  const bool sym = (hc.C.d2m == hc.C.i2m) && (hc.C.m2i == hc.C.m2d);  // as launch_viterbi
  for (int k = 1; k <= kmax; ++k) {
    std::vector<Task> fails(plan.n_pairs + 1);
    uint32_t nfail = 0;
    FailSink sink;
    sink.items = fails.data(); sink.count = &nfail; sink.capacity = (uint32_t)fails.size();
    for (const Task& T : plan.tasks[k]) {
      if (use_fast) { if (sym) emu_dispatch<MODE_FAST | MODE_SYM>(k, hc.C, B, T, sink, E); else emu_dispatch<MODE_FAST>(k, hc.C, B, T, sink, E); }
      else { if (sym) emu_dispatch<MODE_FULL | MODE_SYM>(k, hc.C, B, T, sink, E); else emu_dispatch<MODE_FULL>(k, hc.C, B, T, sink, E); }
    }
    nfall += nfail;
    FailSink none; uint32_t zero = 0;
    none.items = nullptr; none.count = &zero; none.capacity = 0;
    for (uint32_t f = 0; f < nfail; ++f) {
      if (sym) emu_dispatch<MODE_FULL | MODE_SYM>(k, hc.C, B, fails[f], none, E);
      else emu_dispatch<MODE_FULL>(k, hc.C, B, fails[f], none, E);
    }
  }
  expand_ll_host(*b, plan, uniq_ll.data(), out_ll);
  if (n_fallback) *n_fallback = nfall;
  return LTR_OK;
}

extern "C" void ltr_emu_band_geometry(int n, int m, int W, int* dlo, int* w, uint64_t* cells) {
  const BandGeom g = band_geometry(n, m, W);
  *dlo = g.dlo;
  *w = g.w;
  *cells = band_cells(n, m, W, g.dlo);
}

// Margin (diagonals) the plan asks of a pair's band class for a haplotype of n rows, or -1 when banding is off.
extern "C" int ltr_emu_band_margin(const ltr_params* p, int band_w, int n) {
  const BandPolicy bp = band_policy(*p, band_w);
  return bp.on ? band_margin_needed(bp, n) : -1;
}

// Plan invariants (viterbi_host.h) for the CPU tests.  Returns 0 when all hold, otherwise the number of the first
// violated one:
//  1 every pooled read maps to a distinct read of its own locus with the same bytes;
//  2 the distinct reads of a locus are numbered by non-decreasing length and are pairwise different;
//  3 every (haplotype, distinct read) pair is covered exactly once by the band tasks and the stream tasks together;
//  4 every pair of a band task has the band class of its list and a certifiable margin; task row classes are right.
// counts: [0] distinct reads, [1] band pairs, [2] stream pairs.
extern "C" int ltr_emu_plan_check(const ltr_viterbi_batch* b, const ltr_params* p, int kmax, int band_w, uint64_t* counts) {
  Plan plan;
  if (make_plan(*b, *p, kmax, plan, 0, nullptr, nullptr, band_w) != LTR_OK) return -1;
  const int cut = 35 - p->indel_flank_len;
  const uint32_t n_loci = b->n_loci;
  const uint32_t n_reads = b->locus_read_begin[n_loci];
  for (uint32_t r = 0; r < n_reads; ++r) {
    const uint32_t l = plan.read_locus[r], u = plan.read_to_uread[r];
    if (!(b->locus_read_begin[l] <= r && r < b->locus_read_begin[l + 1])) return 1;
    if (!(plan.locus_uread_begin[l] <= u && u < plan.locus_uread_begin[l + 1])) return 1;
    const uint32_t len = b->read_off[r + 1] - b->read_off[r];
    if (plan.uread_off[u + 1] - plan.uread_off[u] != len) return 1;
    if (std::memcmp(plan.uread_bytes + plan.uread_off[u], b->read_bytes + b->read_off[r], len) != 0) return 1;
  }
  for (uint32_t l = 0; l < n_loci; ++l)
    for (uint32_t u = plan.locus_uread_begin[l]; u + 1 < plan.locus_uread_begin[l + 1]; ++u) {
      const uint32_t la = plan.uread_off[u + 1] - plan.uread_off[u], lb = plan.uread_off[u + 2] - plan.uread_off[u + 1];
      if (la > lb) return 2;
      for (uint32_t v = u + 1; v < plan.locus_uread_begin[l + 1]; ++v) {
        const uint32_t lv = plan.uread_off[v + 1] - plan.uread_off[v];
        if (lv != la) break;
        if (std::memcmp(plan.uread_bytes + plan.uread_off[u], plan.uread_bytes + plan.uread_off[v], la) == 0) return 2;
      }
    }
  const uint32_t n_haps = b->locus_hap_begin[n_loci];
  std::vector<std::vector<uint8_t> > seen(n_haps);
  for (uint32_t h = 0; h < n_haps; ++h) {
    const uint32_t l = plan.hap_locus[h];
    seen[h].assign(plan.locus_uread_begin[l + 1] - plan.locus_uread_begin[l], 0);
  }
  uint64_t band_pairs = 0, stream_pairs = 0;
  for (int c = 0; c < kBandClasses; ++c)
    for (const BandTask& t : plan.band_tasks[(size_t)c]) {
      const uint32_t l = plan.hap_locus[t.hap];
      const int hlen = (int)(b->hap_off[t.hap + 1] - b->hap_off[t.hap]), n = hlen - 2 * cut;
      for (uint32_t u = t.read_begin; u < t.read_end; ++u) {
        if (u < plan.locus_uread_begin[l] || u >= plan.locus_uread_begin[l + 1]) return 3;
        if (seen[t.hap][u - plan.locus_uread_begin[l]]++) return 3;
        const int m = (int)(plan.uread_off[u + 1] - plan.uread_off[u]);
        if (band_class_of(hlen, n, m, plan.band) != c) return 4;
        if (band_geometry(n, m, band_class_w(c)).w < band_margin_needed(plan.band, n)) return 4;
        ++band_pairs;
      }
    }
  for (int k = 1; k <= kmax; ++k)
    for (const Task& t : plan.tasks[(size_t)k]) {
      const uint32_t l = plan.hap_locus[t.hap];
      const int hlen = (int)(b->hap_off[t.hap + 1] - b->hap_off[t.hap]), n = hlen - 2 * cut;
      if (hlen > 60 && n >= 1 && rows_per_lane(n, kmax) != k) return 4;
      for (uint32_t u = t.read_begin; u < t.read_end; ++u) {
        if (u < plan.locus_uread_begin[l] || u >= plan.locus_uread_begin[l + 1]) return 3;
        if (seen[t.hap][u - plan.locus_uread_begin[l]]++) return 3;
        ++stream_pairs;
      }
    }
  for (uint32_t h = 0; h < n_haps; ++h) {
    const uint32_t l = plan.hap_locus[h];
    if (b->locus_read_begin[l + 1] == b->locus_read_begin[l]) continue;
    for (uint8_t s : seen[h])
      if (s != 1) return 3;
  }
  if (band_pairs != plan.n_band_pairs || band_pairs + stream_pairs != plan.n_pairs_computed) return 3;
  if (counts) {
    counts[0] = plan.locus_uread_begin[n_loci];
    counts[1] = band_pairs;
    counts[2] = stream_pairs;
  }
  return 0;
}

// ---- device plan (plan_device.cuh) run serially on the host, held against make_plan ------------------------------------
#include "plan_device.cuh"

#include <map>
#include <set>
#include <tuple>

// Returns 0 when the device plan's products are identical to make_plan's (numbering of the distinct reads, offsets,
// bytes, read map, statistics, and the SETS of band / stream tasks per class; the order inside a list may differ),
// otherwise the number of the first difference.  -1: make_plan rejected the batch (the device plan must flag it too,
// which is then reported as -2 if it did not).  pad: readable bytes around the string buffers, as on the device.
extern "C" int ltr_emu_device_plan_check(const ltr_viterbi_batch* b, const ltr_params* p, int kmax, int band_w,
                                         uint64_t* counts) {
  Plan plan;
  const int rc_host = make_plan(*b, *p, kmax, plan, 1, nullptr, nullptr, band_w);
  const uint32_t n_loci = b->n_loci;
  const uint32_t n_haps = b->locus_hap_begin[n_loci], n_reads = b->locus_read_begin[n_loci];
  const size_t pad = 512;
  const uint32_t raw_total = n_reads ? b->read_off[n_reads] : 0u;
  std::vector<uint8_t> raw(pad + raw_total + pad, 0), ubytes_buf(pad + raw_total + pad, 0);
  if (raw_total) std::memcpy(raw.data() + pad, b->read_bytes, raw_total);
  std::vector<unsigned long long> rhash(n_reads + 1), ull_off(n_loci + 1), stat(PLAN_STAT_WORDS, 0);
  std::vector<uint32_t> rlen(n_reads + 1), rep(n_reads + 1), rank_of(n_reads + 1), tmp_len(n_reads + 1), tmp_rep(n_reads + 1),
      ucount(n_loci + 1), ubytes(n_loci + 1), ubyte_off(n_loci + 1), local_u(n_reads + 1), read_locus(n_reads + 1),
      hap_locus(n_haps + 1), lub(n_loci + 1), uread_off(n_reads + 2), r2u(n_reads + 1), ctl(PLAN_CTL_WORDS, 0),
      band_task_pos((size_t)kBandClasses * n_loci + 1, 0), band_pair_pos((size_t)kBandClasses * n_loci + 1, 0);
  uint64_t n_pairs = 0;
  for (uint32_t l = 0; l < n_loci; ++l)
    n_pairs += (uint64_t)(b->locus_hap_begin[l + 1] - b->locus_hap_begin[l]) * (b->locus_read_begin[l + 1] - b->locus_read_begin[l]);
  std::vector<BandTask> band_tasks(n_pairs + 1);
  std::vector<PlanPair> band_pairs(n_pairs + 1);
  std::vector<std::vector<Task> > st(kPlanMaxK + 1, std::vector<Task>(n_pairs + 1));
  std::vector<uint32_t> st_n(kPlanMaxK + 1, 0);
  PlanDev P;
  P.n_loci = n_loci; P.n_haps = n_haps; P.n_reads = n_reads; P.raw_total = raw_total;
  P.cut = 35 - p->indel_flank_len; P.kmax = kmax; P.band = band_policy(*p, band_w);
  P.lhb = b->locus_hap_begin; P.lrb = b->locus_read_begin; P.hap_off = b->hap_off; P.read_off = b->read_off;
  P.read_bytes = raw.data() + pad;
  P.rhash = rhash.data(); P.rlen = rlen.data(); P.rep = rep.data(); P.rank_of = rank_of.data();
  P.tmp_len = tmp_len.data(); P.tmp_rep = tmp_rep.data(); P.ucount = ucount.data(); P.ubytes = ubytes.data();
  P.ubyte_off = ubyte_off.data(); P.band_task_pos = band_task_pos.data(); P.band_pair_pos = band_pair_pos.data();
  P.local_u = local_u.data(); P.read_locus = read_locus.data();
  P.hap_locus = hap_locus.data(); P.lub = lub.data(); P.ull_off = ull_off.data(); P.uread_off = uread_off.data();
  P.uread_bytes = ubytes_buf.data() + pad; P.r2u = r2u.data(); P.ctl = ctl.data(); P.stat = stat.data();
  P.band_tasks = band_tasks.data(); P.band_pairs = band_pairs.data(); P.band_cap = (uint32_t)n_pairs;
  for (int k = 0; k <= kPlanMaxK; ++k) {
    P.st_tasks[k] = st[(size_t)k].data();
    P.st_cap[k] = (uint32_t)n_pairs;
    P.st_ntasks[k] = &st_n[(size_t)k];
  }
  for (uint32_t l = 0; l < n_loci; ++l) plan_locus_dedupe(P, l, 0, 1, P.read_bytes, 0u, P.raw_total);
  plan_scan_serial(P);
  for (uint32_t l = 0; l < n_loci; ++l) plan_locus_fill(P, l, 0, 1);
  for (uint32_t l = 0; l < n_loci; ++l) plan_locus_tasks(P, l, 0);
  plan_band_scan_serial(P);
  plan_task_scan(P);
  for (uint32_t l = 0; l < n_loci; ++l) plan_locus_tasks(P, l, 1);
  if (rc_host != LTR_OK) return ctl[PLAN_CTL_ERR] ? -1 : -2;
  if (ctl[PLAN_CTL_ERR]) return 1;
  if (ctl[PLAN_CTL_N_UREADS] != plan.locus_uread_begin[n_loci]) return 2;
  for (uint32_t l = 0; l <= n_loci; ++l)
    if (lub[l] != plan.locus_uread_begin[l] || ull_off[l] != plan.ull_off[l]) return 3;
  for (uint32_t u = 0; u <= plan.locus_uread_begin[n_loci]; ++u)
    if (uread_off[u] != plan.uread_off[u]) return 4;
  if (plan.uread_nbytes && std::memcmp(P.uread_bytes, plan.uread_bytes, plan.uread_nbytes) != 0) return 5;
  for (uint32_t r = 0; r < n_reads; ++r)
    if (r2u[r] != plan.read_to_uread[r] || read_locus[r] != plan.read_locus[r]) return 6;
  for (uint32_t h = 0; h < n_haps; ++h)
    if (hap_locus[h] != plan.hap_locus[h]) return 7;
  if (stat[PLAN_STAT_CELLS] != plan.n_cells || stat[PLAN_STAT_CELLS_STREAM] != plan.n_cells_computed ||
      stat[PLAN_STAT_PAIRS_COMPUTED] != plan.n_pairs_computed || (int)stat[PLAN_STAT_MAX_M] != plan.max_m) return 8;
  if (ctl[PLAN_CTL_N_BAND_PAIRS] != plan.n_band_pairs) return 9;
  typedef std::tuple<uint32_t, uint32_t, uint32_t> T3;
  for (int c = 0; c < kBandClasses; ++c) {
    // same tasks in the same order as make_plan (run single-threaded): locus by locus, haplotype by haplotype
    std::vector<T3> a, d;
    for (const BandTask& t : plan.band_tasks[(size_t)c]) a.push_back(T3(t.hap, t.read_begin, t.read_end));
    const uint32_t t0 = ctl[PLAN_CTL_BAND_TASK_BASE + c], nt = ctl[PLAN_CTL_BAND_TASK_COUNT + c];
    uint32_t np = 0;
    for (uint32_t t = t0; t < t0 + nt; ++t) {
      d.push_back(T3(band_tasks[t].hap, band_tasks[t].read_begin, band_tasks[t].read_end));
      np += band_tasks[t].read_end - band_tasks[t].read_begin;
    }
    if (a != d) return 10;
    if (np != ctl[PLAN_CTL_BAND_INFO + 2 * c + 1]) return 11;
    // the pair list of the class holds exactly the pairs of its tasks
    std::vector<std::pair<uint32_t, uint32_t> > pa, pd;
    for (const BandTask& t : plan.band_tasks[(size_t)c])
      for (uint32_t u = t.read_begin; u < t.read_end; ++u) pa.push_back(std::make_pair(t.hap, u));
    const uint32_t p0 = ctl[PLAN_CTL_BAND_INFO + 2 * c];
    for (uint32_t i = p0; i < p0 + np; ++i) pd.push_back(std::make_pair(band_pairs[i].x, band_pairs[i].y));
    if (pa != pd) return 12;
  }
  uint64_t n_stream_tasks = 0, host_stream_tasks = 0;
  for (int k = 1; k <= kmax; ++k) host_stream_tasks += plan.tasks[(size_t)k].size();
  for (int k = 1; k <= kmax; ++k) {
    std::multiset<T3> a, d;
    for (const Task& t : plan.tasks[(size_t)k]) a.insert(T3(t.hap, t.read_begin, t.read_end));
    for (uint32_t i = 0; i < st_n[(size_t)k]; ++i)
      d.insert(T3(st[(size_t)k][i].hap, st[(size_t)k][i].read_begin, st[(size_t)k][i].read_end));
    {  // the covered (haplotype, distinct read) pairs agree; so do the tasks themselves unless make_plan cut its tasks
       // into pieces (kWantTasks, small batches only)
      std::multiset<std::pair<uint32_t, uint32_t> > pa, pd;
      for (const T3& t : a) for (uint32_t u = std::get<1>(t); u < std::get<2>(t); ++u) pa.insert(std::make_pair(std::get<0>(t), u));
      for (const T3& t : d) for (uint32_t u = std::get<1>(t); u < std::get<2>(t); ++u) pd.insert(std::make_pair(std::get<0>(t), u));
      if (pa != pd) return 13;
      if (host_stream_tasks >= 2048 && a != d) return 13;
    }
    n_stream_tasks += d.size();
    // heaviest cost bucket first
    uint64_t prev_bucket = 64;
    for (uint32_t i = 0; i < st_n[(size_t)k]; ++i) {
      const Task& t = st[(size_t)k][i];
      const int hlen = (int)(b->hap_off[t.hap + 1] - b->hap_off[t.hap]), n = hlen - 2 * P.cut;
      const bool real = hlen > 60 && n >= 1;
      const int strips = real ? std::max(1, (n - 1 + 32 * k - 1) / (32 * k)) : 1;
      const uint64_t q = uread_off[t.read_end] - uread_off[t.read_begin];
      const uint64_t cost = real ? (uint64_t)k * strips * (q + 32) : (uint64_t)(t.read_end - t.read_begin);
      uint64_t bucket = 0;
      for (uint64_t v = cost; v > 1; v >>= 1) ++bucket;
      if (bucket > 31) bucket = 31;
      if (bucket > prev_bucket) return 14;
      prev_bucket = bucket;
    }
  }
  if (counts) {
    counts[0] = ctl[PLAN_CTL_N_UREADS];
    counts[1] = ctl[PLAN_CTL_N_BAND_PAIRS];
    counts[2] = n_stream_tasks;
  }
  return 0;
}
