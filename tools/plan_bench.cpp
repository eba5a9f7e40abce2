// Host-only timing of make_plan (longtr_b200/csrc/viterbi_host.h) on a synthetic workload -- no GPU needed.
//   g++ -O2 -std=c++17 -Iinclude -Ilongtr_b200/csrc tools/plan_bench.cpp -o /tmp/plan_bench -ldl -lpthread
//   LTR_TIMING=1 /tmp/plan_bench <config 3|4> <n_loci> <threads>
// (LTR_HOST_EMU only selects the host flavour of the shared headers; nothing here is a CPU fallback of the product.)
#define LTR_HOST_EMU 1
#include "viterbi_host.h"
#include <chrono>
#include <cstdio>
#include <dlfcn.h>
using namespace ltr;
int main(int argc, char** argv) {
  void* h = dlopen(argc > 4 ? argv[4] : "longtr_b200/csrc/liblongtr_b200.so", RTLD_NOW);
  if (!h) { printf("%s\n", dlerror()); return 1; }
  auto gen = (int (*)(int, uint64_t, uint32_t, uint32_t, int, ltr_synth_batch**))dlsym(h, "ltr_synth_generate");
  auto par = (void (*)(int, ltr_params*))dlsym(h, "ltr_synth_params");
  int cfg = argc > 1 ? atoi(argv[1]) : 3; int nl = argc > 2 ? atoi(argv[2]) : 100000; int nt = argc > 3 ? atoi(argv[3]) : 8;
  ltr_synth_batch* sb = nullptr;
  gen(cfg, 20260103, 0, nl, 8, &sb);
  ltr_params p; par(cfg, &p);
  for (int it = 0; it < 3; ++it) {
    Plan plan;
    auto t0 = std::chrono::steady_clock::now();
    make_plan(sb->vit, p, 16, plan, nt, nullptr, nullptr, 0);
    auto t1 = std::chrono::steady_clock::now();
    size_t nt=0; for (auto& v: plan.band_tasks) nt+=v.size(); size_t ns=0; for (auto& v: plan.tasks) ns+=v.size();
    printf("band tasks %zu stream tasks %zu\n", nt, ns);
    printf("plan %.1f ms  band pairs %llu of %llu\n", std::chrono::duration<double, std::milli>(t1 - t0).count(),
           (unsigned long long)plan.n_band_pairs, (unsigned long long)plan.n_pairs_computed);
  }
}
