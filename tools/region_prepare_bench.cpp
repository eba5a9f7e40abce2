// Host-only timing of the per-region preparation of ltr_regions_run (ltr_region_collect + ltr_candidate_alleles) on T threads.
//   g++ -O2 -std=c++17 -Iinclude tools/region_prepare_bench.cpp -Llongtr_b200/csrc -llongtr_b200 -Wl,-rpath,$PWD/longtr_b200/csrc -pthread -o /tmp/region_prepare_bench
//   /tmp/region_prepare_bench <bam> <regions.txt: start stop period per line> <chrom.txt> <threads> [no_assembly]
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

#include "longtr_b200.h"

static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv) {
  if (argc < 5) return 2;
  ltr_bam* bam = nullptr;
  if (ltr_bam_open(argv[1], nullptr, &bam) != LTR_OK) return 3;
  if (!ltr_bam_has_index(bam)) ltr_bam_build_index(bam);
  std::vector<ltr_region> regions;
  {
    std::ifstream f(argv[2]);
    ltr_region r;
    while (f >> r.start >> r.stop >> r.period) regions.push_back(r);
  }
  std::string chrom;
  {
    std::ifstream f(argv[3]);
    f >> chrom;
  }
  const int T = atoi(argv[4]);
  const uint32_t flags = argc > 5 ? LTR_CAND_FLAG_NO_ASSEMBLY : 0u;
  ltr_region_params rp;
  ltr_region_params_default(&rp);
  for (int rep = 0; rep < 3; ++rep) {
    std::atomic<uint32_t> next(0);
    std::vector<double> t_collect(T, 0), t_cand(T, 0);
    std::atomic<uint32_t> n_asm(0), n_ok(0);
    const double t0 = now_ms();
    auto work = [&](int t) {
      const ltr_bam* bams[1] = {bam};
      for (uint32_t r0 = next.fetch_add(4); r0 < regions.size(); r0 = next.fetch_add(4))
      for (uint32_t r = r0; r < std::min<uint32_t>((uint32_t)regions.size(), r0 + 4); ++r) {
        ltr_region_reads* reads = nullptr;
        double a = now_ms();
        int rc = ltr_region_collect(bams, 1, "chrS", regions[r].start, regions[r].stop, (const uint8_t*)chrom.data(), 0,
                                    (int64_t)chrom.size(), &rp, &reads);
        double b = now_ms();
        t_collect[t] += b - a;
        if (rc == LTR_OK && reads->n_reads) {
          ltr_candidates* c = nullptr;
          rc = ltr_candidate_alleles_flags(reads, regions[r].start, regions[r].stop, regions[r].period,
                                           (const uint8_t*)chrom.data(), 0, (int64_t)chrom.size(), 5, flags, &c);
          t_cand[t] += now_ms() - b;
          if (rc == LTR_OK) {
            ++n_ok;
            if (c->assembly_threshold > 0) ++n_asm;
          }
          ltr_candidates_free(c);
        }
        ltr_region_reads_free(reads);
      }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < T; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    const double wall = now_ms() - t0;
    double sc = 0, sa = 0;
    for (int t = 0; t < T; ++t) sc += t_collect[t], sa += t_cand[t];
    printf("threads %d  wall %.1f ms  %.0f regions/s   per region: collect %.3f ms, candidates %.3f ms  (ok %u, assembled %u)\n", T,
           wall, regions.size() / wall * 1e3, sc / regions.size(), sa / regions.size(), n_ok.load(), n_asm.load());
  }
  ltr_bam_close(bam);
  return 0;
}
