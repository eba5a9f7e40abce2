// TEST INFRASTRUCTURE ONLY -- the reference's EMStutterGenotyper (src/em_stutter_genotyper.{h,cpp}, compiled IN PLACE) behind a
// flat C entry point: length-based EM for the stutter model of one locus, as GenotyperBamProcessor::learn_stutter_model
// calls it (src/genotyper_bam_processor.cpp:170-225: ref_allele 0, train(MAX_EM_ITER, ABS_LL_CONVERGE, FRAC_LL_CONVERGE)).
// Private members are reached with -fno-access-control; the iteration log (disp_stats) gives the number of iterations and
// the log-likelihood of each.  The reference's sources are not touched.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <sstream>
#include <string>
#include <vector>

#include "em_stutter_genotyper.h"
#include "mathops.h"

// reads sample-major: sample s owns reads_per_sample[s] consecutive entries of bp_diff / log_p1 / log_p2.
// out_params: in_geom, in_up, in_down, out_geom, out_up, out_down; out_lls[max_iter]: LL per iteration; returns trained (0/1).
extern "C" int ltr_ref_em_train(uint32_t n_samples, const int32_t* reads_per_sample, const int32_t* bp_diff, const double* log_p1,
                                const double* log_p2, int32_t motif_len, int32_t haploid, int32_t max_iter, double abs_conv,
                                double frac_conv, double* out_params, int32_t* out_n_iter, double* out_lls,
                                double* out_log_gt_priors, int32_t* out_n_alleles) {
  static bool init = false;
  if (!init) {
    precompute_integer_logs();
    init = true;
  }
  std::vector<std::vector<int> > bps(n_samples);
  std::vector<std::vector<double> > p1(n_samples), p2(n_samples);
  std::vector<std::string> names;
  size_t k = 0;
  for (uint32_t s = 0; s < n_samples; ++s) {
    names.push_back("S" + std::to_string(s));
    for (int32_t r = 0; r < reads_per_sample[s]; ++r, ++k) {
      bps[s].push_back(bp_diff[k]);
      p1[s].push_back(log_p1[k]);
      p2[s].push_back(log_p2[k]);
    }
  }
  EMStutterGenotyper g(haploid != 0, std::string((size_t)motif_len, 'A'), bps, p1, p2, names, 0);
  std::ostringstream log;
  log.precision(17);
  const bool trained = g.train(max_iter, abs_conv, frac_conv, true, log);
  int n_iter = 0;
  {
    std::istringstream in(log.str());
    std::string line;
    while (std::getline(in, line)) {
      if (line.compare(0, 10, "Iteration ") == 0) {
        const size_t p = line.find("LL = ");
        if (p != std::string::npos && n_iter < max_iter + 1) out_lls[n_iter] = atof(line.c_str() + p + 5);
        ++n_iter;
      }
    }
  }
  *out_n_iter = n_iter;
  StutterModel* m = g.stutter_model_;
  out_params[0] = m->in_geom_;  out_params[1] = m->in_up_;  out_params[2] = m->in_down_;
  out_params[3] = m->out_geom_; out_params[4] = m->out_up_; out_params[5] = m->out_down_;
  *out_n_alleles = g.num_alleles_;
  for (int a = 0; a < g.num_alleles_ && a < 64; ++a) out_log_gt_priors[a] = g.log_gt_priors_[a];
  return trained ? 1 : 0;
}
