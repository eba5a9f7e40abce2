// TEST INFRASTRUCTURE ONLY: host-side lane emulator of kernel 2 (homopolymer / stutter path).
// Drives the SAME per-lane functions the device kernel uses (longtr_b200/csrc/stutter_core.cuh, compiled with
// LTR_HOST_EMU) through a sequential simulation of one warp: 32 lanes, a skew of one column per lane, the
// hand-off lines and last-column arrays that stutter_kernel.cu keeps in shared memory.  Used by
// tests/test_emulator_stutter.py to check the warp-level logic against the oracle without a GPU.
#define LTR_HOST_EMU 1
#include <math.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "host/longtr_host.h"
#include "stutter_core.cuh"

using namespace ltr;

namespace {

struct Emu {
  StutConsts C;
  std::vector<double> int_logs, qlc, qlw;
  // "shared memory" of the warp
  std::vector<double> match, lineM, lineD;
  std::vector<uint8_t> qual;
  std::vector<int32_t> um;
  std::vector<uint8_t> seq, blk;
};

template <typename HapChar>
double wavefront_rows(Emu& E, const FlankView& F, std::vector<double>& last, int32_t row0, int32_t nrows,
                      int32_t first_type, HapChar hapc) {
  double left_prob = 0.0;
  const int32_t per_strip = 32 * kStutRows;
  for (int32_t s0 = 0; s0 < nrows; s0 += per_strip) {
    const int32_t rows = std::min(nrows - s0, per_strip);
    const int32_t t_last = (rows - 1) / kStutRows;
    FlankLane Ln[32];
    for (int lane = 0; lane < 32; ++lane) {
      flank_lane_reset(Ln[lane]);
      for (int k = 0; k < kStutRows; ++k) {
        const int32_t r = lane * kStutRows + k;
        if (r < rows) {
          Ln[lane].type[k] = (s0 + r == 0) ? first_type : ROW_NORMAL;
          Ln[lane].hc[k] = hapc(s0 + r);
        }
      }
    }
    const int32_t nsteps = F.L + t_last;
    for (int32_t step = 0; step < nsteps; ++step) {
      double prevM[32], prevD[32];  // values every lane exported at the end of the previous step (shfl source)
      for (int lane = 0; lane < 32; ++lane) {
        prevM[lane] = Ln[lane].outM;
        prevD[lane] = Ln[lane].outD;
      }
      for (int lane = 0; lane < 32; ++lane) {
        const int32_t j = step - lane;
        if (!(lane <= t_last && j >= 0 && j < F.L)) continue;
        double aM = lane ? prevM[lane - 1] : E.lineM[j];
        double aD = lane ? prevD[lane - 1] : E.lineD[j];
        double Mout[kStutRows];
        flank_lane_column(Ln[lane], E.C, F, j, aM, aD, Mout);
        if (lane == t_last) {
          E.lineM[j] = Ln[lane].outM;
          E.lineD[j] = Ln[lane].outD;
        }
        if (j == F.L - 1)
          for (int k = 0; k < kStutRows; ++k)
            if (Ln[lane].type[k] != ROW_OFF) last[row0 + s0 + lane * kStutRows + k] = Mout[k];
      }
    }
    if (first_type == ROW_FIRST && s0 == 0) left_prob = Ln[0].left;
  }
  return left_prob;
}

double run_side(Emu& E, int side, const std::string& read, const std::string& qual, int32_t seed, const std::string& lf,
                const std::string& rf, const std::string& allele, const double* art_lp, std::vector<double>& last) {
  const int32_t N = (int32_t)read.size(), L = side == 0 ? seed : N - seed - 1, B = (int32_t)allele.size();
  E.seq.assign(L, 0);
  E.qual.assign(L, 0);
  E.match.assign(L, 0);
  E.lineM.assign(L, nan(""));
  E.lineD.assign(L, nan(""));
  for (int32_t j = 0; j < L; ++j) {
    const int32_t p = side == 0 ? j : N - 1 - j;
    E.seq[j] = (uint8_t)read[p];
    E.qual[j] = (uint8_t)qual[p];
  }
  E.blk.assign(B, 0);
  for (int32_t i = 0; i < B; ++i) E.blk[i] = (uint8_t)(side == 0 ? allele[i] : allele[B - 1 - i]);
  const int32_t n_del = std::min(B, 6);
  E.um.assign((size_t)6 * B, 0);
  for (int32_t k = 0; k < n_del; ++k) {
    int32_t run = 0;
    for (int32_t i = 0; i < B; ++i) {
      run = (i < k + 1) ? 0 : ((E.blk[i - k - 1] != E.blk[i]) ? 0 : run + 1);
      E.um[(size_t)k * B + i] = run;
    }
  }
  FlankView F;
  F.seq = E.seq.data(); F.qual = E.qual.data(); F.tlc = E.qlc.data(); F.tlw = E.qlw.data(); F.L = L;
  F.blk = E.blk.data(); F.B = B; F.um = E.um.data(); F.n_del = n_del; F.match = E.match.data(); F.art_lp = art_lp;
  for (int32_t p = 0; p < L; ++p) E.match[p] = stutter_match_prob(F, p);
  const std::string& fa = side == 0 ? lf : rf;
  const std::string& fc = side == 0 ? rf : lf;
  const int32_t na = (int32_t)fa.size(), nc = (int32_t)fc.size();
  auto hap_a = [&](int32_t r) { return (int32_t)(uint8_t)(side == 0 ? fa[r] : fa[na - 1 - r]); };
  auto hap_c = [&](int32_t r) { return (int32_t)(uint8_t)(side == 0 ? fc[r] : fc[nc - 1 - r]); };
  const double left_prob = wavefront_rows(E, F, last, 0, na, ROW_FIRST, hap_a);
  std::vector<double> out(L);
  for (int32_t j = 0; j < L; ++j) out[j] = stutter_row_cell(E.C, F, E.lineM.data(), j);
  E.lineD = E.lineM;
  E.lineM = out;
  last[na + B - 1] = E.lineM[L - 1];
  wavefront_rows(E, F, last, na + B, nc, ROW_AFTER_STUTTER, hap_c);
  return left_prob;
}

}  // namespace

// One (read, allele) pair. aln_params: 7 floats (ins_ins .. match_del).  Returns 0 and writes *out_ll.
extern "C" int ltr_emu_stutter_pair(const char* lflank, const char* rflank, const char* allele, const char* read,
                                    const char* qual, int32_t seed, const double* stutter6, int32_t motif_len,
                                    const float* aln_params, double* out_ll) {
  Emu E;
  const std::string lf(lflank), rf(rflank), al(allele), rd(read), ql(qual);
  const int32_t N = (int32_t)rd.size();
  if (al.empty() || lf.empty() || rf.empty() || seed < 1 || seed >= N - 1 || ql.size() != rd.size()) return -3;
  E.int_logs.resize(lf.size() + rf.size() + al.size() + 32);
  for (size_t i = 0; i < E.int_logs.size(); ++i) E.int_logs[i] = int_log((int)i);
  E.qlc.resize(256);
  E.qlw.resize(256);
  BaseQuality().byte_tables(E.qlc.data(), E.qlw.data());
  E.C.i2i = (double)aln_params[0]; E.C.i2m = (double)aln_params[1]; E.C.d2d = (double)aln_params[2];
  E.C.d2m = (double)aln_params[3]; E.C.m2m = (double)aln_params[4]; E.C.m2i = (double)aln_params[5];
  E.C.m2d = (double)aln_params[6];
  E.C.log_thresh = log(0.001);
  E.C.int_logs = E.int_logs.data();
  E.C.qual_lc = E.qlc.data();
  E.C.qual_lw = E.qlw.data();
  StutterModel model(stutter6[0], stutter6[1], stutter6[2], stutter6[3], stutter6[4], stutter6[5],
                     std::string((size_t)motif_len, 'N'));
  RepeatStutterInfo info(1, al, model);
  double art[13];
  for (int D = -6; D <= 6; ++D) art[D + 6] = info.log_prob_pcr_artifact(0, D);
  const int32_t n0 = (int32_t)lf.size(), n2 = (int32_t)rf.size(), B = (int32_t)al.size(), hapsize = n0 + B + n2;
  std::vector<double> lastL(hapsize, nan("")), lastR(hapsize, nan(""));
  const double l_prob = run_side(E, 0, rd, ql, seed, lf, rf, al, art, lastL);
  const double r_prob = run_side(E, 1, rd, ql, seed, lf, rf, al, art, lastR);
  // seed join, as in stutter_pair_kernel
  const int32_t seed_char = (int32_t)(uint8_t)rd[seed];
  const double sc = E.qlc[(uint8_t)ql[seed]], sw = E.qlw[(uint8_t)ql[seed]];
  const double prior = -E.int_logs[n0 + n2];
  std::vector<double> terms;
  terms.push_back(((prior + (seed_char == (uint8_t)lf[0] ? sc : sw)) + l_prob) + lastR[hapsize - 2]);
  terms.push_back(((prior + (seed_char == (uint8_t)rf[n2 - 1] ? sc : sw)) + r_prob) + lastL[hapsize - 2]);
  for (int32_t i = 1; i <= hapsize - 2; ++i) {
    if (i >= n0 && i < n0 + B) continue;
    const int32_t hc = (uint8_t)(i < n0 ? lf[i] : rf[i - n0 - B]);
    terms.push_back(((prior + (seed_char == hc ? sc : sw)) + lastL[i - 1]) + lastR[hapsize - 2 - i]);
  }
  double mx = terms[0];
  for (double v : terms) mx = smax(mx, v);
  double total = 0.0;
  for (int k = (int)terms.size() - 1; k >= 0; --k) total += lse_term(E.C, terms[k], mx);  // reversed on purpose: order-free
  *out_ll = lse_finish(mx, total);
  return 0;
}
