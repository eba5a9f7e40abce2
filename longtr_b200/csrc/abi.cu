// abi.cu -- implementation of the C ABI declared in include/longtr_b200.h.
//
// Host side of the drop-in boundary: validates and plans a flattened batch of loci
// (viterbi_host.h), keeps it resident in HBM, launches the sm_100a kernels on the context's
// CUDA streams (one stream per row class so the persistent grids overlap), and moves results
// back.  There is no CPU fallback anywhere in this file: without a CUDA device the context
// cannot be created and every entry point fails.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "kernels.h"
#include "longtr_b200.h"
#include "viterbi_host.h"

using namespace ltr;

#include "ctx.h"

struct ClassState {
  int k = 0;
  uint32_t n_tasks = 0;   // tasks of the plan; band_collect_kernel may append up to task_cap
  uint32_t task_cap = 0;
  uint32_t fail_cap = 0;
  uint32_t grid_fast = 0, grid_full = 0;
  DeviceBuffer tasks, fails, ctrl;  // ctrl: [0] fast cursor [1] n_tasks [2] fail count [3] full cursor
  DeviceBuffer sxy, sb;
  uint32_t scratch_stride = 0;
  bool force_full = false;
};

struct ltr_job {
  ltr_params params;
  Plan plan;  // its large arrays (unique read bytes, read maps, offsets) live in the context's pinned staging buffers
              // and are valid during ltr_job_create only; afterwards only the scalars and task lists are used
  HostConsts hc;
  uint32_t n_loci = 0, n_haps = 0, n_reads = 0;
  uint64_t n_ll = 0, n_post = 0, n_tot = 0;
  DeviceBuffer hap_bytes, hap_off, hap_locus, read_bytes, read_off, lhb, lrb, ll_off, out_ll, tabI, tabD;
  DeviceBuffer lub, r2u, rlocus, ull_off, uniq_ll;  // unique-read bookkeeping (Plan)
  // posterior inputs
  bool has_post = false;
  DeviceBuffer lsb, pool, label, p1, p2, nsamp, haploid, post_off, tot_off, post, totals, int_logs;
  uint32_t n_int_logs = 0;
  std::vector<ClassState> classes;
  // banded kernel (band_kernel.cu): tasks of all band classes back to back, their pair lists, control words
  struct BandClass {
    int cls = 0;  // band class index (band_class_k / band_class_g)
    uint32_t task_begin = 0, n_tasks = 0, pair_begin = 0, n_pairs = 0, grid = 0;
  };
  std::vector<BandClass> band_classes;
  uint32_t n_band_tasks = 0;
  DeviceBuffer band_tasks, band_cum, band_pairs, band_ctrl;  // band_ctrl: u32[8] cursors, u32[2] counters, pad, u64[2] stats
  uint64_t plan_cells_computed = 0;
  // Memsets and kernels of ltr_job_create are issued AFTER every host-to-device copy: they execute on the SMs / in
  // stream order behind whatever another context's persistent kernels are doing, and a pageable copy enqueued behind
  // them would block the calling thread for that long (batches in flight on other host threads).
  std::vector<std::pair<void*, size_t>> deferred_zero;
  ltr_job_stats stats;
};
static const size_t kBandCtrlBytes = 72;  // u32[12], u64 uncertified pairs, u64 their n*m cells, u64 band cells evaluated
static const size_t kBandBucketOff = 128, kBandBucketWords = 17 * 32;  // then count / base / fill of band_collect_kernel
static const size_t kBandCtrlAlloc = kBandBucketOff + 3 * kBandBucketWords * sizeof(uint32_t);

namespace {

// The string buffers (haplotypes, reads) carry kStringPad readable bytes on either side: the stream kernel prefetches one
// byte ahead, the band kernel's character windows run up to W/2 + K bytes (W <= 512) ahead of a string's end and,
// during the prologue, up to W/2 bytes in front of its start (values that only reach cells outside the matrix).
static const size_t kStringPad = 512;

template <typename T>
int upload(ltr_ctx* ctx, DeviceBuffer& buf, const T* src, size_t count, size_t pad_bytes, uint64_t* h2d,
           std::vector<std::pair<void*, size_t>>* deferred_zero = nullptr) {
  const size_t bytes = count * sizeof(T);
  const size_t front = pad_bytes;  // padded buffers are padded on both sides; data starts at p + pad_bytes
  LTR_CUDA(ctx, buf.alloc(front + bytes + pad_bytes));
  if (pad_bytes) {
    if (deferred_zero) {
      deferred_zero->push_back(std::make_pair(buf.p, front));
      deferred_zero->push_back(std::make_pair((void*)((char*)buf.p + front + bytes), pad_bytes));
    } else {
      LTR_CUDA(ctx, cudaMemsetAsync(buf.p, 0, front, ctx->main_stream));
      LTR_CUDA(ctx, cudaMemsetAsync((char*)buf.p + front + bytes, 0, pad_bytes, ctx->main_stream));
    }
  }
  if (bytes) LTR_CUDA(ctx, cudaMemcpyAsync((char*)buf.p + front, src, bytes, cudaMemcpyHostToDevice, ctx->main_stream));
  if (h2d) *h2d += bytes;
  return LTR_OK;
}

}  // namespace

extern "C" {

void ltr_params_default(ltr_params* p) {
  // Dindel defaults, reference HapAligner.h:118
  p->ins_ins = -1.0f;
  p->ins_match = (float)-0.458675;
  p->del_del = -1.0f;
  p->del_match = (float)-0.458675;
  p->match_match = (float)-0.00005800168;
  p->match_ins = (float)-10.448214728;
  p->match_del = (float)-10.448214728;
  p->indel_flank_len = 5;
}

const char* ltr_strerror(int code) {
  switch (code) {
    case LTR_OK: return "ok";
    case LTR_ERR_NO_DEVICE: return "no usable CUDA device (there is no CPU fallback)";
    case LTR_ERR_CUDA: return "CUDA runtime error (see ltr_last_error)";
    case LTR_ERR_INVALID: return "invalid argument or malformed batch";
    case LTR_ERR_OOM: return "out of device memory";
    case LTR_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown error";
  }
}

const char* ltr_last_error(const ltr_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }

const char* ltr_version(void) { return "longtr_b200 0.1 (sm_100a)"; }

int ltr_ctx_create(int device, ltr_ctx** out) {
  if (!out) return LTR_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
    cudaGetLastError();
    return LTR_ERR_NO_DEVICE;
  }
  if (cudaSetDevice(device) != cudaSuccess) return LTR_ERR_NO_DEVICE;
  ltr_ctx* ctx = new ltr_ctx();
  ctx->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    delete ctx;
    return LTR_ERR_NO_DEVICE;
  }
  ctx->sm_count = prop.multiProcessorCount;
  {
    cudaMemPool_t pool = nullptr;  // keep freed blocks cached: job create/destroy never goes back to the driver
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  }
  cudaError_t e = cudaStreamCreateWithFlags(&ctx->main_stream, cudaStreamNonBlocking);
  for (int i = 0; i < kNumStreams && e == cudaSuccess; ++i) {
    e = cudaStreamCreateWithFlags(&ctx->streams[i], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_stream[i], cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_start);
  if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_vit);
  if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_end);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_init, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_collect, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    ltr_ctx_destroy(ctx);
    return LTR_ERR_CUDA;
  }
  for (int k = 1; k <= viterbi_max_rows_per_lane(); ++k) {
    ctx->blocks_per_sm[MODE_FAST][k] = viterbi_blocks_per_sm(k, MODE_FAST);
    ctx->blocks_per_sm[MODE_FULL][k] = viterbi_blocks_per_sm(k, MODE_FULL);
    if (ctx->blocks_per_sm[MODE_FAST][k] <= 0 || ctx->blocks_per_sm[MODE_FULL][k] <= 0) {
      // kernel image not loadable on this device (not sm_100a?)
      ltr_ctx_destroy(ctx);
      return LTR_ERR_NO_DEVICE;
    }
  }
  for (int c = 0; c < kBandClasses; ++c) {
    ctx->band_blocks_per_sm[c] = band_blocks_per_sm(c);
    if (ctx->band_blocks_per_sm[c] <= 0) {
      ltr_ctx_destroy(ctx);
      return LTR_ERR_NO_DEVICE;
    }
  }
  if (const char* env = getenv("LTR_BAND")) ctx->band_w = atoi(env);  // diagnostics: initial ltr_ctx_set_band value
  *out = ctx;
  return LTR_OK;
}

int ltr_ctx_set_band(ltr_ctx* ctx, int32_t half_width) {
  if (!ctx) return LTR_ERR_INVALID;
  ctx->band_w = half_width;
  return LTR_OK;
}

void ltr_ctx_destroy(ltr_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  for (int i = 0; i < kNumStreams; ++i) {
    if (ctx->streams[i]) cudaStreamDestroy(ctx->streams[i]);
    if (ctx->ev_stream[i]) cudaEventDestroy(ctx->ev_stream[i]);
  }
  if (ctx->main_stream) cudaStreamDestroy(ctx->main_stream);
  if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
  if (ctx->ev_vit) cudaEventDestroy(ctx->ev_vit);
  if (ctx->ev_end) cudaEventDestroy(ctx->ev_end);
  if (ctx->ev_init) cudaEventDestroy(ctx->ev_init);
  if (ctx->ev_collect) cudaEventDestroy(ctx->ev_collect);
  for (int i = 0; i < 4; ++i)
    if (ctx->stage[i]) cudaFreeHost(ctx->stage[i]);
  delete ctx;
}

void ltr_job_destroy(ltr_ctx* ctx, ltr_job* job) {
  if (!job) return;
  if (ctx) cudaSetDevice(ctx->device);
  DeviceBuffer* bufs[] = {&job->hap_bytes, &job->hap_off, &job->hap_locus, &job->read_bytes, &job->read_off,
                          &job->lhb, &job->lrb, &job->ll_off, &job->out_ll, &job->tabI, &job->tabD,
                          &job->lub, &job->r2u, &job->rlocus, &job->ull_off, &job->uniq_ll,
                          &job->lsb, &job->pool, &job->label, &job->p1, &job->p2, &job->nsamp,
                          &job->haploid, &job->post_off, &job->tot_off, &job->post, &job->totals,
                          &job->int_logs, &job->band_tasks, &job->band_cum, &job->band_pairs, &job->band_ctrl};
  for (DeviceBuffer* b : bufs) b->free();
  for (ClassState& c : job->classes) {
    c.tasks.free(); c.fails.free(); c.ctrl.free(); c.sxy.free(); c.sb.free();
  }
  delete job;
}

int ltr_job_create(ltr_ctx* ctx, const ltr_params* params, const ltr_viterbi_batch* b,
                   const ltr_posterior_batch* post, ltr_job** out) {
  if (!ctx || !params || !b || !out) return LTR_ERR_INVALID;
  *out = nullptr;
  if (b->n_loci && (!b->locus_hap_begin || !b->locus_read_begin || !b->hap_off || !b->read_off))
    return LTR_ERR_INVALID;
  if (params->indel_flank_len < 0 || params->indel_flank_len > 35) return LTR_ERR_INVALID;
  LTR_CUDA(ctx, cudaSetDevice(ctx->device));
  AllocScope alloc_scope(ctx->main_stream);
  ltr_job* job = new ltr_job();
  std::memset(&job->stats, 0, sizeof(job->stats));
  job->params = *params;
  const int kmax = viterbi_max_rows_per_lane();
  static const uint32_t kZero[2] = {0, 0};
  ltr_viterbi_batch bb = *b;
  if (bb.n_loci == 0) {
    bb.locus_hap_begin = bb.locus_read_begin = bb.hap_off = bb.read_off = kZero;
  }
  // unique read bytes are staged in pinned host memory owned by the context (grow-only, reused by the next job)
  struct Stage {
    static uint8_t* get(size_t bytes, int slot, void* user) {
      ltr_ctx* c = static_cast<ltr_ctx*>(user);
      if (slot < 0 || slot >= 4) return nullptr;
      if (bytes > c->stage_bytes[slot]) {
        if (c->stage[slot]) cudaFreeHost(c->stage[slot]);
        c->stage[slot] = nullptr;
        c->stage_bytes[slot] = 0;
        const size_t want = bytes + bytes / 4 + (1u << 20);
        if (cudaHostAlloc(&c->stage[slot], want, cudaHostAllocDefault) != cudaSuccess) {
          cudaGetLastError();
          c->stage[slot] = nullptr;
          return nullptr;  // make_plan falls back to pageable memory
        }
        c->stage_bytes[slot] = want;
      }
      return static_cast<uint8_t*>(c->stage[slot]);
    }
  };
  static const bool timing = getenv("LTR_TIMING") != nullptr;  // diagnostics: host-side phases of job creation on stderr
  const auto t_begin = std::chrono::steady_clock::now();
  if (!plan_offsets_valid(bb)) { delete job; return LTR_ERR_INVALID; }
  job->n_loci = bb.n_loci;
  job->n_haps = bb.locus_hap_begin[bb.n_loci];
  job->n_reads = bb.locus_read_begin[bb.n_loci];
  uint64_t* h2d = &job->stats.h2d_bytes;
  // Uploads that do not depend on the plan are enqueued first: with pinned caller buffers they overlap make_plan.
  {
    int rc_up = upload(ctx, job->hap_bytes, bb.hap_bytes, (size_t)bb.hap_off[job->n_haps], kStringPad, h2d, &job->deferred_zero);
    if (rc_up == LTR_OK) rc_up = upload(ctx, job->hap_off, bb.hap_off, (size_t)job->n_haps + 1, 0, h2d);
    if (rc_up == LTR_OK) rc_up = upload(ctx, job->lhb, bb.locus_hap_begin, (size_t)job->n_loci + 1, 0, h2d);
    if (rc_up == LTR_OK) rc_up = upload(ctx, job->lrb, bb.locus_read_begin, (size_t)job->n_loci + 1, 0, h2d);
    if (rc_up != LTR_OK) { ltr_job_destroy(ctx, job); return rc_up; }
  }
  int rc = make_plan(bb, *params, kmax, job->plan, 0, &Stage::get, ctx, ctx->band_w);
  const auto t_plan = std::chrono::steady_clock::now();
  if (rc != LTR_OK) { ltr_job_destroy(ctx, job); return rc; }
  Plan& plan = job->plan;
  job->n_ll = plan.ll_off[bb.n_loci];
  job->stats.n_pairs = plan.n_pairs;
  job->stats.n_cells = plan.n_cells;
  job->stats.n_pairs_computed = plan.n_pairs_computed;
  job->stats.n_cells_computed = plan.n_cells_computed;
  job->plan_cells_computed = plan.n_cells_computed;  // stream-kernel pairs of the plan (banded pairs: counted by the kernel)
  job->stats.n_band_pairs = plan.n_band_pairs;
  make_consts(*params, std::max(plan.max_n, plan.max_m) + 2, job->hc);

#define LTR_TRY(expr)                     \
  do {                                    \
    int rc__ = (expr);                    \
    if (rc__ != LTR_OK) {                 \
      ltr_job_destroy(ctx, job);          \
      return rc__;                        \
    }                                     \
  } while (0)
#define LTR_CUDA_J(call)                                   \
  do {                                                     \
    cudaError_t e__ = (call);                              \
    if (e__ != cudaSuccess) {                              \
      int rc__ = fail_cuda(ctx, e__, #call);               \
      ltr_job_destroy(ctx, job);                           \
      return rc__;                                         \
    }                                                      \
  } while (0)

  // Only the distinct trimmed reads of each locus travel to the device (Plan, viterbi_host.h); the kernels fill the
  // unique LL matrices and expand_ll_kernel fans them out to the caller-visible aln_probs layout.
  // padding (here and for hap_bytes above): the stream kernel prefetches one byte, the band kernel's character
  // windows run up to W/2 + K bytes ahead
  LTR_TRY(upload(ctx, job->read_bytes, plan.uread_bytes, plan.uread_nbytes, kStringPad, h2d, &job->deferred_zero));
  LTR_TRY(upload(ctx, job->read_off, plan.uread_off.data(), plan.uread_off.size(), 0, h2d));
  LTR_TRY(upload(ctx, job->lub, plan.locus_uread_begin.data(), plan.locus_uread_begin.size(), 0, h2d));
  LTR_TRY(upload(ctx, job->r2u, plan.read_to_uread.data(), plan.read_to_uread.size(), 0, h2d));
  LTR_TRY(upload(ctx, job->rlocus, plan.read_locus.data(), plan.read_locus.size(), 0, h2d));
  LTR_TRY(upload(ctx, job->hap_locus, plan.hap_locus.data(), plan.hap_locus.size(), 0, h2d));
  LTR_TRY(upload(ctx, job->ll_off, plan.ll_off.data(), plan.ll_off.size(), 0, h2d));
  LTR_TRY(upload(ctx, job->ull_off, plan.ull_off.data(), plan.ull_off.size(), 0, h2d));
  LTR_CUDA_J(job->uniq_ll.alloc(plan.ull_off[bb.n_loci] * sizeof(double)));
  LTR_TRY(upload(ctx, job->tabI, job->hc.tabI.data(), job->hc.tabI.size(), 0, h2d));
  LTR_TRY(upload(ctx, job->tabD, job->hc.tabD.data(), job->hc.tabD.size(), 0, h2d));
  LTR_CUDA_J(job->out_ll.alloc(job->n_ll * sizeof(double)));
  job->hc.C.tabI = job->tabI.as<double>();
  job->hc.C.tabD = job->tabD.as<double>();

  for (int k = 1; k <= kmax; ++k) {
    const uint64_t band_extra = plan.band_pairs_by_rows[k];  // tasks band_collect_kernel may append (<= pairs)
    if (plan.tasks[k].empty() && band_extra == 0) continue;
    if (plan.tasks[k].size() + band_extra > 0xFFFFFFF0ull) { ltr_job_destroy(ctx, job); return LTR_ERR_INVALID; }
    ClassState cs;
    cs.k = k;
    cs.force_full = !fast_certificate_valid(*params);  // parameters outside the certificate's condition: exact kernel only
    cs.n_tasks = (uint32_t)plan.tasks[k].size();
    cs.task_cap = (uint32_t)(plan.tasks[k].size() + band_extra);
    uint64_t pairs = band_extra;
    for (const Task& t : plan.tasks[k]) pairs += t.read_end - t.read_begin;
    cs.fail_cap = (uint32_t)std::min<uint64_t>(pairs, 1u << 22);
    const uint32_t warps_per_block = viterbi_block_threads() / 32;
    const uint64_t want_blocks = ((uint64_t)cs.n_tasks + band_extra + warps_per_block - 1) / warps_per_block;
    cs.grid_fast = (uint32_t)std::min<uint64_t>(want_blocks, (uint64_t)(ctx->sm_count * ctx->blocks_per_sm[MODE_FAST][k]));
    cs.grid_full = (uint32_t)(ctx->sm_count * std::min(2, ctx->blocks_per_sm[MODE_FULL][k]));
    const uint32_t grid_max = std::max(cs.grid_fast, cs.grid_full);
    LTR_CUDA_J(cs.tasks.alloc((size_t)cs.task_cap * sizeof(Task)));
    if (cs.n_tasks) {
      LTR_CUDA_J(cudaMemcpyAsync(cs.tasks.p, plan.tasks[k].data(), (size_t)cs.n_tasks * sizeof(Task),
                                 cudaMemcpyHostToDevice, ctx->main_stream));
      *h2d += (size_t)cs.n_tasks * sizeof(Task);
    }
    LTR_CUDA_J(cs.fails.alloc((size_t)cs.fail_cap * sizeof(Task)));
    LTR_CUDA_J(cs.ctrl.alloc(4 * sizeof(uint32_t)));
    // per-warp scratch line: row-0 boundary of the read stream / strip hand-off (viterbi_core.cuh)
    {
      // only haplotypes cut into several strips hand rows over through the scratch line (row class kmax only)
      cs.scratch_stride = plan.multi_strip[k] ? viterbi_scratch_entries(plan.max_q[k]) : 1u;
      const size_t warps = (size_t)grid_max * warps_per_block;
      LTR_CUDA_J(cs.sxy.alloc(warps * cs.scratch_stride * sizeof(XY)));
      LTR_CUDA_J(cs.sb.alloc(warps * cs.scratch_stride * sizeof(uint32_t)));
      job->deferred_zero.push_back(std::make_pair(cs.sxy.p, cs.sxy.bytes));
    }
    job->classes.push_back(cs);
  }

  if (plan.n_band_pairs) {
    std::vector<BandTask> all;
    std::vector<uint32_t> cum;
    uint64_t npairs = 0;
    for (int c = 0; c < kBandClasses; ++c) {
      const std::vector<BandTask>& v = plan.band_tasks[(size_t)c];
      if (v.empty()) continue;
      ltr_job::BandClass bc;
      bc.cls = c;
      bc.task_begin = (uint32_t)all.size();
      bc.n_tasks = (uint32_t)v.size();
      bc.pair_begin = (uint32_t)npairs;
      for (const BandTask& t : v) {
        all.push_back(t);
        cum.push_back((uint32_t)npairs);
        npairs += t.read_end - t.read_begin;
      }
      if (npairs > 0xFFFFFFF0ull) { ltr_job_destroy(ctx, job); return LTR_ERR_INVALID; }
      bc.n_pairs = (uint32_t)npairs - bc.pair_begin;
      const uint32_t warps_per_block = (uint32_t)band_block_threads() / 32;
      const uint32_t ppr = 32u / (uint32_t)band_class_g(c);
      const uint32_t rounds = (bc.n_pairs + ppr - 1) / ppr;
      bc.grid = std::min<uint32_t>((rounds + warps_per_block - 1) / warps_per_block,
                                   (uint32_t)(ctx->sm_count * ctx->band_blocks_per_sm[bc.cls]));
      job->band_classes.push_back(bc);
    }
    job->n_band_tasks = (uint32_t)all.size();
    LTR_TRY(upload(ctx, job->band_tasks, all.data(), all.size(), 0, h2d));
    LTR_TRY(upload(ctx, job->band_cum, cum.data(), cum.size(), 0, h2d));
    LTR_CUDA_J(job->band_pairs.alloc((size_t)npairs * sizeof(uint2)));
    LTR_CUDA_J(job->band_ctrl.alloc(kBandCtrlAlloc));
  }

  if (post) {
    job->has_post = true;
    const uint32_t n_sreads = post->locus_sread_begin[bb.n_loci];
    std::vector<unsigned long long> post_off((size_t)bb.n_loci + 1, 0), tot_off((size_t)bb.n_loci + 1, 0);
    uint32_t max_h = 1;
    for (uint32_t l = 0; l < bb.n_loci; ++l) {
      const uint32_t H = bb.locus_hap_begin[l + 1] - bb.locus_hap_begin[l];
      const uint32_t S = post->locus_n_samples[l];
      const uint32_t P = bb.locus_read_begin[l + 1] - bb.locus_read_begin[l];
      max_h = std::max(max_h, H);
      post_off[l + 1] = post_off[l] + (unsigned long long)S * H * H;
      tot_off[l + 1] = tot_off[l] + S;
      for (uint32_t r = post->locus_sread_begin[l]; r < post->locus_sread_begin[l + 1]; ++r)
        if (post->pool_index[r] >= P || post->sample_label[r] < 0 || (uint32_t)post->sample_label[r] >= S) {
          ltr_job_destroy(ctx, job);
          return LTR_ERR_INVALID;
        }
    }
    job->n_post = post_off[bb.n_loci];
    job->n_tot = tot_off[bb.n_loci];
    // INT_LOGS (mathops.cpp:14-22) with the host's libm so the priors match the reference bit for bit
    std::vector<double> logs((size_t)max_h + 2);
    logs[0] = -1000.0;
    for (uint32_t i = 1; i < logs.size(); ++i) logs[i] = log((double)i);
    job->n_int_logs = (uint32_t)logs.size();
    LTR_TRY(upload(ctx, job->int_logs, logs.data(), logs.size(), 0, h2d));
    LTR_TRY(upload(ctx, job->lsb, post->locus_sread_begin, (size_t)bb.n_loci + 1, 0, h2d));
    LTR_TRY(upload(ctx, job->pool, post->pool_index, n_sreads, 0, h2d));
    LTR_TRY(upload(ctx, job->label, post->sample_label, n_sreads, 0, h2d));
    LTR_TRY(upload(ctx, job->p1, post->log_p1, n_sreads, 0, h2d));
    LTR_TRY(upload(ctx, job->p2, post->log_p2, n_sreads, 0, h2d));
    LTR_TRY(upload(ctx, job->nsamp, post->locus_n_samples, bb.n_loci, 0, h2d));
    if (post->locus_haploid) LTR_TRY(upload(ctx, job->haploid, post->locus_haploid, bb.n_loci, 0, h2d));
    LTR_TRY(upload(ctx, job->post_off, post_off.data(), post_off.size(), 0, h2d));
    LTR_TRY(upload(ctx, job->tot_off, tot_off.data(), tot_off.size(), 0, h2d));
    LTR_CUDA_J(job->post.alloc(job->n_post * sizeof(double)));
    LTR_CUDA_J(job->totals.alloc(job->n_tot * sizeof(double)));
  }
  // every copy from host memory is enqueued: wait for those only (the caller may release its arrays when we return)
  LTR_CUDA_J(cudaEventRecord(ctx->ev_init, ctx->main_stream));
  for (const std::pair<void*, size_t>& z : job->deferred_zero)
    LTR_CUDA_J(cudaMemsetAsync(z.first, 0, z.second, ctx->main_stream));
  job->deferred_zero.clear();
  if (job->n_band_tasks)
    LTR_CUDA_J(launch_band_expand(job->band_tasks.as<BandTask>(), job->band_cum.as<uint32_t>(), job->n_band_tasks,
                                  job->band_pairs.as<uint2>(), ctx->main_stream));
  const auto t_enq = std::chrono::steady_clock::now();
  LTR_CUDA_J(cudaEventSynchronize(ctx->ev_init));
  if (timing) {
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
      return std::chrono::duration<double, std::milli>(b - a).count();
    };
    fprintf(stderr, "[ltr] job_create: plan %.1f ms, alloc+enqueue %.1f ms, copy drain %.1f ms (h2d %.1f MB)\n",
            ms(t_begin, t_plan), ms(t_plan, t_enq), ms(t_enq, std::chrono::steady_clock::now()), *h2d / 1e6);
  }
  *out = job;
  return LTR_OK;
#undef LTR_TRY
#undef LTR_CUDA_J
}

static DevBatch job_dev_batch(const ltr_job* job) {
  DevBatch B;
  B.hap_bytes = job->hap_bytes.as<uint8_t>() + kStringPad;
  B.hap_off = job->hap_off.as<uint32_t>();
  B.hap_locus = job->hap_locus.as<uint32_t>();
  B.read_bytes = job->read_bytes.as<uint8_t>() + kStringPad;
  B.read_off = job->read_off.as<uint32_t>();
  B.locus_hap_begin = job->lhb.as<uint32_t>();
  B.locus_read_begin = job->lub.as<uint32_t>();            // unique reads
  B.ll_off = job->ull_off.as<unsigned long long>();
  B.out_ll = job->uniq_ll.as<double>();
  return B;
}

// Launches the stream kernels of every row class on its stream.  init_ctrl: write the control words first (on the
// class stream); otherwise the caller has initialised them (band phase: band_collect_kernel appends tasks).
// only_forced: re-run of the classes whose fail list overflowed; ntasks_override[c] = tasks incl. appended ones.
static int run_classes(ltr_ctx* ctx, ltr_job* job, bool only_forced, bool init_ctrl,
                       const std::vector<uint32_t>* ntasks_override) {
  const DevBatch B = job_dev_batch(job);
  int si = 0;
  for (size_t ci = 0; ci < job->classes.size(); ++ci) {
    ClassState& cs = job->classes[ci];
    if (only_forced && !cs.force_full) continue;
    static const bool serial = getenv("LTR_SERIAL_CLASSES") != nullptr;  // diagnostics: one row class at a time
    cudaStream_t st = ctx->streams[serial ? 0 : (si % kNumStreams)];
    ++si;
    if (init_ctrl) {
      const uint32_t ctrl_init[4] = {0u, ntasks_override ? (*ntasks_override)[ci] : cs.n_tasks, 0u, 0u};
      LTR_CUDA(ctx, cudaMemcpyAsync(cs.ctrl.p, ctrl_init, sizeof(ctrl_init), cudaMemcpyHostToDevice, st));
    }
    uint32_t* ctrl = cs.ctrl.as<uint32_t>();
    FailSink sink;
    sink.items = cs.fails.as<Task>();
    sink.count = ctrl + 2;
    sink.capacity = cs.fail_cap;
    FailSink none;
    none.items = nullptr;
    none.count = ctrl + 2;
    none.capacity = 0;
    if (cs.force_full) {
      // parameters outside the certificate's condition, or the fail list overflowed on an earlier run: exact kernel only
      LTR_CUDA(ctx, launch_viterbi(cs.k, MODE_FULL, (int)cs.grid_fast, st, job->hc.C, B, cs.tasks.as<Task>(),
                                   ctrl + 1, cs.task_cap, ctrl + 0, none, cs.sxy.as<XY>(),
                                   cs.sb.as<uint32_t>(), cs.scratch_stride));
      job->stats.n_launches += 1;
    } else {
      LTR_CUDA(ctx, launch_viterbi(cs.k, MODE_FAST, (int)cs.grid_fast, st, job->hc.C, B, cs.tasks.as<Task>(),
                                   ctrl + 1, cs.task_cap, ctrl + 0, sink, cs.sxy.as<XY>(),
                                   cs.sb.as<uint32_t>(), cs.scratch_stride));
      LTR_CUDA(ctx, launch_viterbi(cs.k, MODE_FULL, (int)cs.grid_full, st, job->hc.C, B, cs.fails.as<Task>(),
                                   ctrl + 2, cs.fail_cap, ctrl + 3, none, cs.sxy.as<XY>(),
                                   cs.sb.as<uint32_t>(), cs.scratch_stride));
      job->stats.n_launches += 2;
    }
  }
  return LTR_OK;
}

// Band phase: the banded kernels of every band class, then band_collect_kernel, all ordered before the stream kernels.
static int run_band_phase(ltr_ctx* ctx, ltr_job* job) {
  const DevBatch B = job_dev_batch(job);
  for (ClassState& cs : job->classes) {
    const uint32_t ctrl_init[4] = {0u, cs.n_tasks, 0u, 0u};
    LTR_CUDA(ctx, cudaMemcpyAsync(cs.ctrl.p, ctrl_init, sizeof(ctrl_init), cudaMemcpyHostToDevice, ctx->main_stream));
  }
  LTR_CUDA(ctx, cudaMemsetAsync(job->band_ctrl.p, 0, kBandCtrlAlloc, ctx->main_stream));
  LTR_CUDA(ctx, cudaEventRecord(ctx->ev_init, ctx->main_stream));
  uint32_t* bctrl = job->band_ctrl.as<uint32_t>();
  static const bool no_abandon = getenv("LTR_BAND_NO_ABANDON") != nullptr;  // diagnostics
  for (size_t i = 0; i < job->band_classes.size(); ++i) {
    const ltr_job::BandClass& bc = job->band_classes[i];
    cudaStream_t st = ctx->streams[i % kNumStreams];
    LTR_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_init, 0));
    BandArgs A;
    A.pairs = job->band_pairs.as<uint2>() + bc.pair_begin;
    A.n_pairs = bc.n_pairs;
    A.cursor = bctrl + i;
    A.counters = bctrl + 8;
    A.cells_evaluated = reinterpret_cast<unsigned long long*>(job->band_ctrl.as<char>() + 64);
    A.gap = job->plan.band.gap;
    A.abandon_after = no_abandon ? 0u : 4096u;
    LTR_CUDA(ctx, launch_band(bc.cls, (int)bc.grid, st, job->hc.C, B, A));
    job->stats.n_launches += 1;
    LTR_CUDA(ctx, cudaEventRecord(ctx->ev_stream[i % kNumStreams], st));
    LTR_CUDA(ctx, cudaStreamWaitEvent(ctx->main_stream, ctx->ev_stream[i % kNumStreams], 0));
  }
  BandCollect S;
  for (int k = 0; k < 17; ++k) {
    S.tasks[k] = nullptr;
    S.count[k] = bctrl + 10;  // never read: capacity 0
    S.cap[k] = 0;
  }
  for (ClassState& cs : job->classes) {
    S.tasks[cs.k] = cs.tasks.as<Task>();
    S.count[cs.k] = cs.ctrl.as<uint32_t>() + 1;
    S.cap[cs.k] = cs.task_cap;
  }
  S.kmax = viterbi_max_rows_per_lane();
  S.n_uncertified = reinterpret_cast<unsigned long long*>(job->band_ctrl.as<char>() + 48);
  S.cells_uncertified = reinterpret_cast<unsigned long long*>(job->band_ctrl.as<char>() + 56);
  S.bucket_count = reinterpret_cast<uint32_t*>(job->band_ctrl.as<char>() + kBandBucketOff);
  S.bucket_base = S.bucket_count + kBandBucketWords;
  S.bucket_fill = S.bucket_base + kBandBucketWords;
  LTR_CUDA(ctx, launch_band_collect(job->hc.C, B, job->band_tasks.as<BandTask>(), job->n_band_tasks, S,
                                    ctx->main_stream));
  job->stats.n_launches += 3;
  LTR_CUDA(ctx, cudaEventRecord(ctx->ev_collect, ctx->main_stream));
  return LTR_OK;
}

int ltr_job_run(ltr_ctx* ctx, ltr_job* job) {
  if (!ctx || !job) return LTR_ERR_INVALID;
  LTR_CUDA(ctx, cudaSetDevice(ctx->device));
  job->stats.n_launches = 0;
  job->stats.n_fallback = 0;
  job->stats.n_band_uncertified = 0;
  const int n_used = (int)std::min<size_t>(kNumStreams, job->classes.size());  // streams run_classes touches
  const bool band = !job->band_classes.empty();
  LTR_CUDA(ctx, cudaEventRecord(ctx->ev_start, ctx->main_stream));
  if (band) {
    int rcb = run_band_phase(ctx, job);
    if (rcb != LTR_OK) return rcb;
    for (int i = 0; i < n_used; ++i) LTR_CUDA(ctx, cudaStreamWaitEvent(ctx->streams[i], ctx->ev_collect, 0));
  } else {
    for (int i = 0; i < n_used; ++i) LTR_CUDA(ctx, cudaStreamWaitEvent(ctx->streams[i], ctx->ev_start, 0));
  }
  int rc = run_classes(ctx, job, false, !band, nullptr);
  if (rc != LTR_OK) return rc;
  for (int i = 0; i < n_used; ++i) {
    LTR_CUDA(ctx, cudaEventRecord(ctx->ev_stream[i], ctx->streams[i]));
    LTR_CUDA(ctx, cudaStreamWaitEvent(ctx->main_stream, ctx->ev_stream[i], 0));
  }
  // fail counts back to the host: fallback statistics + overflow detection
  std::vector<uint32_t> ctrl(job->classes.size() * 4, 0);
  for (size_t c = 0; c < job->classes.size(); ++c)
    LTR_CUDA(ctx, cudaMemcpyAsync(&ctrl[c * 4], job->classes[c].ctrl.p, 4 * sizeof(uint32_t),
                                  cudaMemcpyDeviceToHost, ctx->main_stream));
  unsigned long long band_words[kBandCtrlBytes / 8] = {0};
  if (band)
    LTR_CUDA(ctx, cudaMemcpyAsync(band_words, job->band_ctrl.p, kBandCtrlBytes, cudaMemcpyDeviceToHost, ctx->main_stream));
  LTR_CUDA(ctx, cudaStreamSynchronize(ctx->main_stream));
  job->stats.n_band_uncertified = band_words[6];
  // cells evaluated: full matrices of the stream-kernel pairs (planned + uncertified) + the bands actually evaluated
  job->stats.n_cells_computed = job->plan_cells_computed + band_words[7] + band_words[8];
  bool rerun = false;
  std::vector<uint32_t> ntasks(job->classes.size(), 0);
  for (size_t c = 0; c < job->classes.size(); ++c) {
    ClassState& cs = job->classes[c];
    ntasks[c] = std::min(ctrl[c * 4 + 1], cs.task_cap);
    if (cs.force_full) continue;
    job->stats.n_fallback += ctrl[c * 4 + 2];
    if (ctrl[c * 4 + 2] > cs.fail_cap) {
      cs.force_full = true;
      rerun = true;
    }
  }
  if (rerun) {
    for (int i = 0; i < kNumStreams; ++i) LTR_CUDA(ctx, cudaStreamWaitEvent(ctx->streams[i], ctx->ev_start, 0));
    rc = run_classes(ctx, job, true, true, &ntasks);
    if (rc != LTR_OK) return rc;
    for (int i = 0; i < kNumStreams; ++i) {
      LTR_CUDA(ctx, cudaEventRecord(ctx->ev_stream[i], ctx->streams[i]));
      LTR_CUDA(ctx, cudaStreamWaitEvent(ctx->main_stream, ctx->ev_stream[i], 0));
    }
  }
  {
    ExpandArgs E;
    E.n_reads = job->n_reads;
    E.read_locus = job->rlocus.as<uint32_t>();
    E.read_to_uread = job->r2u.as<uint32_t>();
    E.locus_hap_begin = job->lhb.as<uint32_t>();
    E.locus_read_begin = job->lrb.as<uint32_t>();
    E.locus_uread_begin = job->lub.as<uint32_t>();
    E.ll_off = job->ll_off.as<unsigned long long>();
    E.ull_off = job->ull_off.as<unsigned long long>();
    E.uniq_ll = job->uniq_ll.as<double>();
    E.out_ll = job->out_ll.as<double>();
    LTR_CUDA(ctx, launch_expand_ll(E, ctx->main_stream));
    if (job->n_reads) job->stats.n_launches += 1;
  }
  LTR_CUDA(ctx, cudaEventRecord(ctx->ev_vit, ctx->main_stream));
  if (job->has_post) {
    DevPosterior P;
    P.n_loci = job->n_loci;
    P.locus_hap_begin = job->lhb.as<uint32_t>();
    P.locus_sread_begin = job->lsb.as<uint32_t>();
    P.pool_index = job->pool.as<uint32_t>();
    P.sample_label = job->label.as<int32_t>();
    P.log_p1 = job->p1.as<double>();
    P.log_p2 = job->p2.as<double>();
    P.locus_n_samples = job->nsamp.as<uint32_t>();
    P.locus_haploid = job->haploid.p ? job->haploid.as<uint8_t>() : nullptr;
    P.ll_off = job->ll_off.as<unsigned long long>();
    P.post_off = job->post_off.as<unsigned long long>();
    P.tot_off = job->tot_off.as<unsigned long long>();
    P.ll = job->out_ll.as<double>();
    P.int_logs = job->int_logs.as<double>();
    P.n_int_logs = job->n_int_logs;
    P.log_one_half = log(0.5);
    P.post = job->post.as<double>();
    P.totals = job->totals.as<double>();
    LTR_CUDA(ctx, launch_posteriors(P, ctx->main_stream));
    job->stats.n_launches += 1;
  }
  LTR_CUDA(ctx, cudaEventRecord(ctx->ev_end, ctx->main_stream));
  LTR_CUDA(ctx, cudaStreamSynchronize(ctx->main_stream));
  float ms_total = 0.f, ms_vit = 0.f;
  LTR_CUDA(ctx, cudaEventElapsedTime(&ms_total, ctx->ev_start, ctx->ev_end));
  LTR_CUDA(ctx, cudaEventElapsedTime(&ms_vit, ctx->ev_start, ctx->ev_vit));
  job->stats.kernel_ms = ms_total;
  job->stats.viterbi_ms = ms_vit;
  return LTR_OK;
}

void ltr_job_sizes(const ltr_job* job, uint64_t* n_ll, uint64_t* n_post, uint64_t* n_totals) {
  if (n_ll) *n_ll = job ? job->n_ll : 0;
  if (n_post) *n_post = job ? job->n_post : 0;
  if (n_totals) *n_totals = job ? job->n_tot : 0;
}

int ltr_job_download(ltr_ctx* ctx, ltr_job* job, double* out_ll, double* out_post, double* out_totals) {
  if (!ctx || !job) return LTR_ERR_INVALID;
  LTR_CUDA(ctx, cudaSetDevice(ctx->device));
  job->stats.d2h_bytes = 0;
  if (out_ll && job->n_ll) {
    LTR_CUDA(ctx, cudaMemcpyAsync(out_ll, job->out_ll.p, job->n_ll * sizeof(double), cudaMemcpyDeviceToHost,
                                  ctx->main_stream));
    job->stats.d2h_bytes += job->n_ll * sizeof(double);
  }
  if (out_post && job->has_post && job->n_post) {
    LTR_CUDA(ctx, cudaMemcpyAsync(out_post, job->post.p, job->n_post * sizeof(double), cudaMemcpyDeviceToHost,
                                  ctx->main_stream));
    job->stats.d2h_bytes += job->n_post * sizeof(double);
  }
  if (out_totals && job->has_post && job->n_tot) {
    LTR_CUDA(ctx, cudaMemcpyAsync(out_totals, job->totals.p, job->n_tot * sizeof(double), cudaMemcpyDeviceToHost,
                                  ctx->main_stream));
    job->stats.d2h_bytes += job->n_tot * sizeof(double);
  }
  LTR_CUDA(ctx, cudaStreamSynchronize(ctx->main_stream));
  return LTR_OK;
}

void ltr_job_get_stats(const ltr_job* job, ltr_job_stats* stats) {
  if (job && stats) *stats = job->stats;
}

int ltr_viterbi_ll(ltr_ctx* ctx, const ltr_params* params, const ltr_viterbi_batch* batch, double* out_ll,
                   ltr_job_stats* stats) {
  ltr_job* job = nullptr;
  int rc = ltr_job_create(ctx, params, batch, nullptr, &job);
  if (rc != LTR_OK) return rc;
  rc = ltr_job_run(ctx, job);
  if (rc == LTR_OK) rc = ltr_job_download(ctx, job, out_ll, nullptr, nullptr);
  if (stats) ltr_job_get_stats(job, stats);
  ltr_job_destroy(ctx, job);
  return rc;
}

int ltr_posteriors(ltr_ctx* ctx, int haploid, int32_t n_samples, int32_t n_reads, int32_t n_alleles,
                   double* ll, const double* log_p1, const double* log_p2, const int32_t* sample_label,
                   double* post, double* totals, double* total_ll) {
  if (!ctx || !ll || !log_p1 || !log_p2 || !sample_label || !post || !totals) return LTR_ERR_INVALID;
  if (n_samples <= 0 || n_alleles <= 0 || n_reads < 0) return LTR_ERR_INVALID;
  LTR_CUDA(ctx, cudaSetDevice(ctx->device));
  AllocScope alloc_scope(ctx->main_stream);
  for (int r = 0; r < n_reads; ++r)
    if (sample_label[r] < 0 || sample_label[r] >= n_samples) return LTR_ERR_INVALID;
  const size_t H = (size_t)n_alleles, R = (size_t)n_reads, S = (size_t)n_samples;
  DeviceBuffer d_ll, d_p1, d_p2, d_lab, d_pool, d_lhb, d_lsb, d_ns, d_hap, d_off, d_post, d_tot, d_logs;
  std::vector<uint32_t> pool(R);
  for (size_t r = 0; r < R; ++r) pool[r] = (uint32_t)r;
  const uint32_t lhb[2] = {0u, (uint32_t)H}, lsb[2] = {0u, (uint32_t)R}, ns[1] = {(uint32_t)S};
  const uint8_t hp[1] = {(uint8_t)(haploid ? 1 : 0)};
  const unsigned long long offs[6] = {0ull, (unsigned long long)(R * H), 0ull, (unsigned long long)(S * H * H),
                                      0ull, (unsigned long long)S};
  std::vector<double> logs(H + 2);
  logs[0] = -1000.0;
  for (size_t i = 1; i < logs.size(); ++i) logs[i] = log((double)i);
  int rc = LTR_OK;
  auto up = [&](DeviceBuffer& b, const void* src, size_t bytes) {
    if (rc != LTR_OK) return;
    cudaError_t e = b.alloc(bytes);
    if (e == cudaSuccess && bytes) e = cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, ctx->main_stream);
    if (e != cudaSuccess) rc = fail_cuda(ctx, e, "ltr_posteriors upload");
  };
  up(d_ll, ll, R * H * 8); up(d_p1, log_p1, R * 8); up(d_p2, log_p2, R * 8); up(d_lab, sample_label, R * 4);
  up(d_pool, pool.data(), R * 4); up(d_lhb, lhb, 8); up(d_lsb, lsb, 8); up(d_ns, ns, 4); up(d_hap, hp, 1);
  up(d_off, offs, sizeof(offs)); up(d_logs, logs.data(), logs.size() * 8);
  if (rc == LTR_OK) {
    cudaError_t e = d_post.alloc(S * H * H * 8);
    if (e == cudaSuccess) e = d_tot.alloc(S * 8);
    if (e != cudaSuccess) rc = fail_cuda(ctx, e, "ltr_posteriors alloc");
  }
  if (rc == LTR_OK) {
    DevPosterior P;
    P.n_loci = 1;
    P.locus_hap_begin = d_lhb.as<uint32_t>();
    P.locus_sread_begin = d_lsb.as<uint32_t>();
    P.pool_index = d_pool.as<uint32_t>();
    P.sample_label = d_lab.as<int32_t>();
    P.log_p1 = d_p1.as<double>();
    P.log_p2 = d_p2.as<double>();
    P.locus_n_samples = d_ns.as<uint32_t>();
    P.locus_haploid = d_hap.as<uint8_t>();
    P.ll_off = d_off.as<unsigned long long>();
    P.post_off = d_off.as<unsigned long long>() + 2;
    P.tot_off = d_off.as<unsigned long long>() + 4;
    P.ll = d_ll.as<double>();
    P.int_logs = d_logs.as<double>();
    P.n_int_logs = (uint32_t)logs.size();
    P.log_one_half = log(0.5);
    P.post = d_post.as<double>();
    P.totals = d_tot.as<double>();
    cudaError_t e = launch_posteriors(P, ctx->main_stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(post, d_post.p, S * H * H * 8, cudaMemcpyDeviceToHost, ctx->main_stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(totals, d_tot.p, S * 8, cudaMemcpyDeviceToHost, ctx->main_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->main_stream);
    if (e != cudaSuccess) rc = fail_cuda(ctx, e, "ltr_posteriors run");
  }
  DeviceBuffer* all[] = {&d_ll, &d_p1, &d_p2, &d_lab, &d_pool, &d_lhb, &d_lsb, &d_ns, &d_hap, &d_off, &d_post,
                         &d_tot, &d_logs};
  for (DeviceBuffer* b : all) b->free();
  if (rc != LTR_OK) return rc;
  // the reference clamps log_aln_probs_ in place (genotyper.cpp:57-58); mirror that on the caller's array
  for (size_t i = 0; i < R * H; ++i)
    if (ll[i] < -600.0) ll[i] = -600.0;
  if (total_ll) {
    double t = 0.0;  // sum(), mathops.cpp:24-29
    for (size_t s = 0; s < S; ++s) t += totals[s];
    *total_ll = t;
  }
  return LTR_OK;
}

}  // extern "C"
