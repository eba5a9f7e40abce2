// abi.cu -- implementation of the C ABI declared in include/longtr_b200.h.
//
// Host side of the drop-in boundary.  A job (one flattened batch of loci) is ONE stream-ordered sequence on the device:
//   upload (h2d stream) -> plan -> banded kernels -> collect -> full-matrix stream kernels -> fan-out -> posteriors
//   (main + class streams) -> download (d2h stream)
// with nothing in between that needs the host: task counts, fail lists and statistics stay in device memory and come
// back with the results.  ltr_job_submit enqueues all of it and returns; ltr_job_wait blocks on the last event.
// The plan (de-duplication of the trimmed reads, band classes, task lists) is built on the device for large batches
// (plan_kernels.cu) and on the host for the small ones of the per-locus entry points (make_plan, viterbi_host.h), where
// a handful of launches matter more than the plan's cost.  There is no CPU fallback anywhere in this file: without a
// CUDA device the context cannot be created and every entry point fails.
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

namespace {

struct ClassState {
  int k = 0;
  uint32_t n_tasks = 0;   // host plan: tasks of the plan (device plan: counted on the device); band_collect_kernel may
  uint32_t task_cap = 0;  // append up to task_cap
  uint32_t fail_cap = 0;
  uint32_t grid_fast = 0, grid_full = 0;
  DeviceBuffer tasks, fails;
  DeviceBuffer sxy, sb;
  uint32_t scratch_stride = 0;
  bool force_full = false;
};

// What comes back from the device with every run (pinned host memory, one block per job).
struct JobResult {
  uint32_t cls_ctrl[(kPlanMaxK + 1) * 4];  // per row class: fast cursor, tasks, fail count, full cursor
  unsigned long long band_words[13];       // band_ctrl: cursors, counters, statistics
  uint32_t plan_ctl[PLAN_CTL_BAND_TASK_COUNT];
  unsigned long long plan_stat[PLAN_STAT_WORDS];
};

// u32[16] (round cursor per band class, [12] pairs evaluated, [13] of which uncertified, [14] [15] scratch), then from
// byte 64: u64 uncertified pairs, u64 n*m cells of those re-run over the full matrix, u64 band cells evaluated, u64 pairs
// given a second band round, u64 of which still uncertified
static_assert(kBandClasses <= 12, "band_ctrl layout");
static const size_t kBandStatOff = 64, kBandCtrlBytes = kBandStatOff + 5 * 8;
static const size_t kBandBucketOff = 128, kBandBucketWords = 17 * 32;  // then count / base / fill of band_collect_kernel
// second band round: retry_count[16], retry_fill[16], retry_info[32], cursors[16]
static const size_t kBandRetryOff = kBandBucketOff + 3 * kBandBucketWords * sizeof(uint32_t), kBandRetryWords = 80;
static const size_t kBandCtrlAlloc = kBandRetryOff + kBandRetryWords * sizeof(uint32_t);

// Second band round (band_retry_class): a retry must cost less than this percentage of the full matrix; 0 switches the
// second round off (LTR_BAND_RETRY_RHO, diagnostics).
static int band_retry_rho() {
  static const int rho = [] {
    const char* e = getenv("LTR_BAND_RETRY_RHO");
    const int v = e ? atoi(e) : 80;
    return v < 0 ? 0 : (v > 100 ? 100 : v);
  }();
  return rho;
}

// The string buffers (haplotypes, reads) carry kStringPad readable bytes on either side: the stream kernel prefetches one
// byte ahead, the band kernel's character windows run up to W/2 + K bytes (W <= 512) ahead of a string's end and,
// during the prologue, up to W/2 bytes in front of its start (values that only reach cells outside the matrix).
static const size_t kStringPad = 512;

// Batches with at least this many pooled reads are planned on the device (ltr_ctx_set_plan overrides).
static const uint32_t kDevicePlanMinReads = 4096;

}  // namespace

struct ltr_job {
  ltr_ctx* ctx = nullptr;
  int lane = 0;
  bool device_plan = false;
  ltr_params params;
  Plan plan;  // host plan: its large arrays live in the context's pinned staging buffers and are valid during job setup
              // only; afterwards only the scalars and task lists are used
  HostConsts hc;
  BandPolicy band;
  uint32_t n_loci = 0, n_haps = 0, n_reads = 0, raw_bytes = 0;
  uint64_t n_ll = 0, n_post = 0, n_tot = 0, n_upairs_cap = 0;
  DeviceBuffer hap_bytes, hap_off, hap_locus, read_bytes, read_off, lhb, lrb, ll_off, out_ll, tabI, tabD;
  DeviceBuffer lub, r2u, rlocus, ull_off, uniq_ll;  // distinct-read bookkeeping (host or device plan)
  DeviceBuffer raw_bytes_d, raw_off_d, plan_scratch, plan_ctl, plan_stat;  // device plan only
  DeviceBuffer raw_packed_d;           // device plan, read encoding 1: the 4-bit stream as uploaded
  std::vector<uint8_t> host_unpacked;  // host plan, read encoding 1: the reads as bytes
  PlanDev pd;
  // posterior inputs
  bool has_post = false;
  DeviceBuffer lsb, pool, label, p1, p2, nsamp, haploid, post_off, tot_off, post, totals, int_logs;
  DeviceBuffer mate, aligned, kept_mask, kept_index;  // optional: mate pairs, voting reads, removal of uncalled alleles
  bool prune = false;
  uint32_t n_int_logs = 0, n_sreads = 0;
  std::vector<ClassState> classes;
  DeviceBuffer cls_ctrl;  // u32[(kPlanMaxK+1)*4]
  // banded kernel (band_kernel.cu): tasks of all band classes back to back, their pair lists, control words
  struct BandLaunch {
    int cls = 0;
    uint32_t grid = 0;
  };
  std::vector<BandLaunch> band_launches;
  uint32_t band_cap = 0;  // capacity of band_tasks / band_pairs
  DeviceBuffer band_tasks, band_cum, band_pairs, band_ctrl, band_meta;  // band_meta (host plan): info[16], n_band_tasks
  DeviceBuffer band_retry;  // pair lists of the second band round (band_collect_kernel)
  uint32_t band_retry_cap = 0;
  const uint32_t* band_info_dev = nullptr;
  const uint32_t* n_band_tasks_dev = nullptr;
  uint64_t plan_cells_computed = 0;
  // Memsets of job setup are issued on the main stream of the lane, never between host-to-device copies.
  std::vector<std::pair<void*, size_t>> deferred_zero;
  cudaEvent_t ev_h2d = nullptr, ev_start = nullptr, ev_plan = nullptr, ev_vit = nullptr, ev_end = nullptr,
              ev_stats = nullptr, ev_done = nullptr;
  JobResult* res = nullptr;
  bool pending = false;  // submitted, not yet waited for
  bool ran = false;
  bool drained = true;   // nothing of this job is queued on the device any more (destroy need not wait)
  cudaStream_t st_h2d = nullptr, st_d2h = nullptr;  // device plan: the lane's copy streams; host plan (small jobs): main
  ltr_job_stats stats;
};

namespace {

template <typename T>
int upload(ltr_ctx* ctx, cudaStream_t st, DeviceBuffer& buf, const T* src, size_t count, size_t pad_bytes, uint64_t* h2d,
           std::vector<std::pair<void*, size_t>>* deferred_zero = nullptr) {
  const size_t bytes = count * sizeof(T);
  const size_t front = pad_bytes;  // padded buffers are padded on both sides; data starts at p + pad_bytes
  LTR_CUDA(ctx, buf.alloc(front + bytes + pad_bytes));
  if (pad_bytes && deferred_zero) {
    deferred_zero->push_back(std::make_pair(buf.p, front));
    deferred_zero->push_back(std::make_pair((void*)((char*)buf.p + front + bytes), pad_bytes));
  }
  if (bytes) LTR_CUDA(ctx, cudaMemcpyAsync((char*)buf.p + front, src, bytes, cudaMemcpyHostToDevice, st));
  if (h2d) *h2d += bytes;
  return LTR_OK;
}

JobResult* take_result_block(ltr_ctx* ctx) {
  if (!ctx->result_blocks.empty()) {
    JobResult* r = static_cast<JobResult*>(ctx->result_blocks.back());
    ctx->result_blocks.pop_back();
    return r;
  }
  void* p = nullptr;
  if (cudaHostAlloc(&p, sizeof(JobResult), cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return static_cast<JobResult*>(p);
}

cudaError_t take_event(ltr_ctx* ctx, bool timing, cudaEvent_t* out) {
  std::vector<cudaEvent_t>& pool = timing ? ctx->timing_events : ctx->plain_events;
  if (!pool.empty()) {
    *out = pool.back();
    pool.pop_back();
    return cudaSuccess;
  }
  return timing ? cudaEventCreate(out) : cudaEventCreateWithFlags(out, cudaEventDisableTiming);
}

void destroy_lane(JobLane& L) {
  for (int i = 0; i < kNumStreams; ++i) {
    if (L.cls[i]) cudaStreamDestroy(L.cls[i]);
    if (L.ev_cls[i]) cudaEventDestroy(L.ev_cls[i]);
  }
  if (L.main) cudaStreamDestroy(L.main);
  if (L.h2d) cudaStreamDestroy(L.h2d);
  if (L.d2h) cudaStreamDestroy(L.d2h);
  if (L.ev_init) cudaEventDestroy(L.ev_init);
  if (L.ev_collect) cudaEventDestroy(L.ev_collect);
  L = JobLane();
}

cudaError_t create_lane(JobLane& L) {
  cudaError_t e = cudaStreamCreateWithFlags(&L.main, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&L.h2d, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&L.d2h, cudaStreamNonBlocking);
  for (int i = 0; i < kNumStreams && e == cudaSuccess; ++i) {
    e = cudaStreamCreateWithFlags(&L.cls[i], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&L.ev_cls[i], cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&L.ev_init, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&L.ev_collect, cudaEventDisableTiming);
  return e;
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

const char* ltr_version(void) { return "longtr_b200 0.2 (sm_100a)"; }

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
  for (int i = 0; i < kLanes && e == cudaSuccess; ++i) e = create_lane(ctx->lanes[i]);
  if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_start);
  if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_end);
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
  if (const char* env = getenv("LTR_PLAN")) ctx->plan_mode = atoi(env);  // diagnostics: initial ltr_ctx_set_plan value
  *out = ctx;
  return LTR_OK;
}

int ltr_ctx_set_band(ltr_ctx* ctx, int32_t half_width) {
  if (!ctx) return LTR_ERR_INVALID;
  ctx->band_w = half_width;
  return LTR_OK;
}

int ltr_ctx_set_read_encoding(ltr_ctx* ctx, int32_t encoding) {
  if (!ctx || encoding < 0 || encoding > 1) return LTR_ERR_INVALID;
  ctx->read_encoding = encoding;
  return LTR_OK;
}

int ltr_ctx_set_plan(ltr_ctx* ctx, int32_t mode) {
  if (!ctx || mode < 0 || mode > 2) return LTR_ERR_INVALID;
  ctx->plan_mode = mode;
  return LTR_OK;
}

void ltr_ctx_destroy(ltr_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (int i = 0; i < kLanes; ++i) destroy_lane(ctx->lanes[i]);
  if (ctx->main_stream) cudaStreamDestroy(ctx->main_stream);
  if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
  if (ctx->ev_end) cudaEventDestroy(ctx->ev_end);
  for (int i = 0; i < 4; ++i)
    if (ctx->stage[i]) cudaFreeHost(ctx->stage[i]);
  for (void* p : ctx->result_blocks) cudaFreeHost(p);
  for (cudaEvent_t e : ctx->timing_events) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->plain_events) cudaEventDestroy(e);
  delete ctx;
}

void ltr_job_destroy(ltr_ctx* ctx, ltr_job* job) {
  if (!job) return;
  if (!ctx) ctx = job->ctx;
  if (ctx) cudaSetDevice(ctx->device);
  if (ctx && !job->drained) {
    // a job may be destroyed while its work is still queued (error paths, a caller that gives up): drain its lane first
    JobLane& L = ctx->lanes[job->lane];
    cudaStreamSynchronize(L.h2d);
    cudaStreamSynchronize(L.main);
    for (int i = 0; i < kNumStreams; ++i) cudaStreamSynchronize(L.cls[i]);
    cudaStreamSynchronize(L.d2h);
  }
  DeviceBuffer* bufs[] = {&job->hap_bytes, &job->hap_off, &job->hap_locus, &job->read_bytes, &job->read_off,
                          &job->lhb, &job->lrb, &job->ll_off, &job->out_ll, &job->tabI, &job->tabD,
                          &job->lub, &job->r2u, &job->rlocus, &job->ull_off, &job->uniq_ll,
                          &job->raw_bytes_d, &job->raw_packed_d, &job->raw_off_d, &job->plan_scratch, &job->plan_ctl, &job->plan_stat,
                          &job->lsb, &job->pool, &job->label, &job->p1, &job->p2, &job->nsamp,
                          &job->haploid, &job->post_off, &job->tot_off, &job->post, &job->totals,
                          &job->int_logs, &job->mate, &job->aligned, &job->kept_mask, &job->kept_index, &job->cls_ctrl, &job->band_tasks, &job->band_cum, &job->band_pairs,
                          &job->band_ctrl, &job->band_meta, &job->band_retry};
  for (DeviceBuffer* b : bufs) b->free();
  for (ClassState& c : job->classes) {
    c.tasks.free(); c.fails.free(); c.sxy.free(); c.sb.free();
  }
  cudaEvent_t tev[] = {job->ev_start, job->ev_plan, job->ev_vit, job->ev_end};
  cudaEvent_t pev[] = {job->ev_h2d, job->ev_stats, job->ev_done};
  for (cudaEvent_t e : tev)
    if (e) { if (ctx) ctx->timing_events.push_back(e); else cudaEventDestroy(e); }
  for (cudaEvent_t e : pev)
    if (e) { if (ctx) ctx->plain_events.push_back(e); else cudaEventDestroy(e); }
  if (job->res && ctx) ctx->result_blocks.push_back(job->res);
  delete job;
}

}  // extern "C"

namespace {

#define LTR_TRY(expr)        \
  do {                       \
    int rc__ = (expr);       \
    if (rc__ != LTR_OK) return rc__; \
  } while (0)

// Row-class bookkeeping shared by both plan modes: buffers of one class, sized for `task_cap` tasks.
int setup_class(ltr_ctx* ctx, ltr_job* job, int k, uint32_t n_tasks, uint64_t task_cap, uint64_t pairs, bool exact_grid,
                bool multi_strip, uint32_t max_q) {
  if (task_cap > 0xFFFFFFF0ull || pairs > 0xFFFFFFF0ull) return LTR_ERR_INVALID;
  ClassState cs;
  cs.k = k;
  cs.force_full = !fast_certificate_valid(job->params);  // parameters outside the certificate's condition: exact kernel only
  cs.n_tasks = n_tasks;
  cs.task_cap = (uint32_t)task_cap;
  cs.fail_cap = (uint32_t)pairs;  // every pair can fail the final-score certificate: the list never overflows
  const uint32_t warps_per_block = viterbi_block_threads() / 32;
  const uint32_t full_grid = (uint32_t)(ctx->sm_count * ctx->blocks_per_sm[MODE_FAST][k]);
  const uint64_t want_blocks = (task_cap + warps_per_block - 1) / warps_per_block;
  cs.grid_fast = exact_grid ? (uint32_t)std::min<uint64_t>(want_blocks, full_grid) : full_grid;
  if (cs.grid_fast == 0) cs.grid_fast = 1;
  cs.grid_full = (uint32_t)(ctx->sm_count * std::min(2, ctx->blocks_per_sm[MODE_FULL][k]));
  const uint32_t grid_max = std::max(cs.grid_fast, cs.grid_full);
  LTR_CUDA(ctx, cs.tasks.alloc((size_t)cs.task_cap * sizeof(Task)));
  LTR_CUDA(ctx, cs.fails.alloc((size_t)cs.fail_cap * sizeof(Task)));
  // per-warp scratch line: strip hand-off of haplotypes cut into several strips (viterbi_core.cuh)
  cs.scratch_stride = multi_strip ? viterbi_scratch_entries(max_q) : 1u;
  const size_t warps = (size_t)grid_max * warps_per_block;
  LTR_CUDA(ctx, cs.sxy.alloc(warps * cs.scratch_stride * sizeof(XY)));
  LTR_CUDA(ctx, cs.sb.alloc(warps * cs.scratch_stride * sizeof(uint32_t)));
  job->deferred_zero.push_back(std::make_pair(cs.sxy.p, cs.sxy.bytes));
  job->classes.push_back(cs);
  return LTR_OK;
}

int setup_posteriors(ltr_ctx* ctx, ltr_job* job, const ltr_viterbi_batch& bb, const ltr_posterior_batch* post,
                     cudaStream_t st, bool validate_reads_on_host) {
  uint64_t* h2d = &job->stats.h2d_bytes;
  const uint32_t n_loci = bb.n_loci;
  if (n_loci && (!post->locus_sread_begin || !post->locus_n_samples)) return LTR_ERR_INVALID;
  static const uint32_t kZero1[1] = {0};
  const uint32_t* lsb = n_loci ? post->locus_sread_begin : kZero1;
  for (uint32_t l = 0; l < n_loci; ++l)
    if (lsb[l + 1] < lsb[l]) return LTR_ERR_INVALID;
  const uint32_t n_sreads = lsb[n_loci];
  if (n_sreads && (!post->pool_index || !post->sample_label || !post->log_p1 || !post->log_p2)) return LTR_ERR_INVALID;
  job->has_post = true;
  job->n_sreads = n_sreads;
  std::vector<unsigned long long> post_off((size_t)n_loci + 1, 0), tot_off((size_t)n_loci + 1, 0);
  uint32_t max_h = 1;
  for (uint32_t l = 0; l < n_loci; ++l) {
    const uint32_t H = bb.locus_hap_begin[l + 1] - bb.locus_hap_begin[l];
    const uint32_t S = post->locus_n_samples[l];
    max_h = std::max(max_h, H);
    const unsigned long long shh = (unsigned long long)S * H * H;
    if (S > (1u << 20) || shh > (1ull << 31)) return LTR_ERR_INVALID;  // the kernels index a locus' entries with 32 bits
    post_off[l + 1] = post_off[l] + shh;
    tot_off[l + 1] = tot_off[l] + S;
    if (validate_reads_on_host) {
      const uint32_t P = bb.locus_read_begin[l + 1] - bb.locus_read_begin[l];
      for (uint32_t r = lsb[l]; r < lsb[l + 1]; ++r)
        if (post->pool_index[r] >= P || post->sample_label[r] < 0 || (uint32_t)post->sample_label[r] >= S)
          return LTR_ERR_INVALID;
    }
  }
  if (post_off[n_loci] > (1ull << 40)) return LTR_ERR_INVALID;
  job->n_post = post_off[n_loci];
  job->n_tot = tot_off[n_loci];
  // INT_LOGS (mathops.cpp:14-22) with the host's libm so the priors match the reference bit for bit
  std::vector<double> logs((size_t)max_h + 2);
  logs[0] = -1000.0;
  for (uint32_t i = 1; i < logs.size(); ++i) logs[i] = log((double)i);
  job->n_int_logs = (uint32_t)logs.size();
  LTR_TRY(upload(ctx, st, job->int_logs, logs.data(), logs.size(), 0, h2d));
  LTR_TRY(upload(ctx, st, job->lsb, lsb, (size_t)n_loci + 1, 0, h2d));
  LTR_TRY(upload(ctx, st, job->pool, post->pool_index, n_sreads, 0, h2d));
  LTR_TRY(upload(ctx, st, job->label, post->sample_label, n_sreads, 0, h2d));
  LTR_TRY(upload(ctx, st, job->p1, post->log_p1, n_sreads, 0, h2d));
  LTR_TRY(upload(ctx, st, job->p2, post->log_p2, n_sreads, 0, h2d));
  LTR_TRY(upload(ctx, st, job->nsamp, post->locus_n_samples, n_loci, 0, h2d));
  if (post->locus_haploid) LTR_TRY(upload(ctx, st, job->haploid, post->locus_haploid, n_loci, 0, h2d));
  if (post->second_mate) LTR_TRY(upload(ctx, st, job->mate, post->second_mate, n_sreads, 0, h2d));
  if (post->read_aligned) LTR_TRY(upload(ctx, st, job->aligned, post->read_aligned, n_sreads, 0, h2d));
  job->prune = post->prune_uncalled != 0;
  if (job->prune) {
    LTR_CUDA(ctx, job->kept_mask.alloc((size_t)job->n_haps));
    LTR_CUDA(ctx, job->kept_index.alloc((size_t)job->n_haps * 4));
  }
  LTR_TRY(upload(ctx, st, job->post_off, post_off.data(), post_off.size(), 0, h2d));
  LTR_TRY(upload(ctx, st, job->tot_off, tot_off.data(), tot_off.size(), 0, h2d));
  LTR_CUDA(ctx, job->post.alloc(job->n_post * sizeof(double)));
  LTR_CUDA(ctx, job->totals.alloc(job->n_tot * sizeof(double)));
  return LTR_OK;
}

// ---- job setup, host plan (small batches): make_plan on the host, exact sizes and grids -------------------------------
int setup_host_plan(ltr_ctx* ctx, ltr_job* job, const ltr_viterbi_batch& bb, cudaStream_t st) {
  const int kmax = viterbi_max_rows_per_lane();
  uint64_t* h2d = &job->stats.h2d_bytes;
  // the plan's large arrays are staged in pinned host memory owned by the context (grow-only, reused by the next job)
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
  LTR_TRY(make_plan(bb, job->params, kmax, job->plan, 0, &Stage::get, ctx, ctx->band_w));
  Plan& plan = job->plan;
  job->band = plan.band;
  job->n_ll = plan.ll_off[bb.n_loci];
  job->stats.n_pairs = plan.n_pairs;
  job->stats.n_cells = plan.n_cells;
  job->stats.n_pairs_computed = plan.n_pairs_computed;
  job->stats.n_cells_computed = plan.n_cells_computed;
  job->plan_cells_computed = plan.n_cells_computed;  // stream-kernel pairs of the plan (banded pairs: counted by the kernel)
  job->stats.n_band_pairs = plan.n_band_pairs;
  make_consts(job->params, std::max(plan.max_n, plan.max_m) + 2, job->hc);
  // Only the distinct trimmed reads of each locus travel to the device; the kernels fill the distinct LL matrices and
  // expand_ll_kernel fans them out to the caller-visible aln_probs layout.
  LTR_TRY(upload(ctx, st, job->read_bytes, plan.uread_bytes, plan.uread_nbytes, kStringPad, h2d, &job->deferred_zero));
  LTR_TRY(upload(ctx, st, job->read_off, plan.uread_off.data(), plan.uread_off.size(), 0, h2d));
  LTR_TRY(upload(ctx, st, job->lub, plan.locus_uread_begin.data(), plan.locus_uread_begin.size(), 0, h2d));
  LTR_TRY(upload(ctx, st, job->r2u, plan.read_to_uread.data(), plan.read_to_uread.size(), 0, h2d));
  LTR_TRY(upload(ctx, st, job->rlocus, plan.read_locus.data(), plan.read_locus.size(), 0, h2d));
  LTR_TRY(upload(ctx, st, job->hap_locus, plan.hap_locus.data(), plan.hap_locus.size(), 0, h2d));
  LTR_TRY(upload(ctx, st, job->ll_off, plan.ll_off.data(), plan.ll_off.size(), 0, h2d));
  LTR_TRY(upload(ctx, st, job->ull_off, plan.ull_off.data(), plan.ull_off.size(), 0, h2d));
  job->n_upairs_cap = plan.ull_off[bb.n_loci];
  LTR_CUDA(ctx, job->uniq_ll.alloc(job->n_upairs_cap * sizeof(double)));
  for (int k = 1; k <= kmax; ++k) {
    const uint64_t band_extra = plan.band_pairs_by_rows[k];  // tasks band_collect_kernel may append (<= pairs)
    if (plan.tasks[k].empty() && band_extra == 0) continue;
    uint64_t pairs = band_extra;
    for (const Task& t : plan.tasks[k]) pairs += t.read_end - t.read_begin;
    LTR_TRY(setup_class(ctx, job, k, (uint32_t)plan.tasks[k].size(), plan.tasks[k].size() + band_extra, pairs, true,
                        plan.multi_strip[k] != 0, plan.max_q[k]));
    ClassState& cs = job->classes.back();
    if (cs.n_tasks) {
      LTR_CUDA(ctx, cudaMemcpyAsync(cs.tasks.p, plan.tasks[k].data(), (size_t)cs.n_tasks * sizeof(Task),
                                    cudaMemcpyHostToDevice, st));
      *h2d += (size_t)cs.n_tasks * sizeof(Task);
    }
  }
  if (plan.n_band_pairs) {
    std::vector<BandTask> all;
    std::vector<uint32_t> cum;
    uint32_t meta[2 * kBandClasses + 1];
    std::memset(meta, 0, sizeof(meta));
    uint64_t npairs = 0;
    for (int c = 0; c < kBandClasses; ++c) {
      const std::vector<BandTask>& v = plan.band_tasks[(size_t)c];
      meta[2 * c] = (uint32_t)npairs;
      if (v.empty()) continue;
      for (const BandTask& t : v) {
        all.push_back(t);
        cum.push_back((uint32_t)npairs);
        npairs += t.read_end - t.read_begin;
      }
      if (npairs > 0xFFFFFFF0ull) return LTR_ERR_INVALID;
      meta[2 * c + 1] = (uint32_t)npairs - meta[2 * c];
      const uint32_t warps_per_block = (uint32_t)band_block_threads() / 32;
      const uint32_t ppr = 32u / (uint32_t)band_class_g(c);
      const uint32_t rounds = (meta[2 * c + 1] + ppr - 1) / ppr;
      ltr_job::BandLaunch bl;
      bl.cls = c;
      bl.grid = std::min<uint32_t>((rounds + warps_per_block - 1) / warps_per_block,
                                   (uint32_t)(ctx->sm_count * ctx->band_blocks_per_sm[c]));
      job->band_launches.push_back(bl);
    }
    meta[2 * kBandClasses] = (uint32_t)all.size();
    job->band_cap = (uint32_t)all.size();
    LTR_TRY(upload(ctx, st, job->band_tasks, all.data(), all.size(), 0, h2d));
    LTR_TRY(upload(ctx, st, job->band_cum, cum.data(), cum.size(), 0, h2d));
    LTR_TRY(upload(ctx, st, job->band_meta, meta, 2 * kBandClasses + 1, 0, h2d));
    LTR_CUDA(ctx, job->band_pairs.alloc((size_t)npairs * sizeof(uint2)));
    LTR_CUDA(ctx, job->band_ctrl.alloc(kBandCtrlAlloc));
    if (band_retry_rho() > 0) {
      job->band_retry_cap = (uint32_t)npairs;
      LTR_CUDA(ctx, job->band_retry.alloc((size_t)npairs * sizeof(uint2)));
    }
    job->band_info_dev = job->band_meta.as<uint32_t>();
    job->n_band_tasks_dev = job->band_meta.as<uint32_t>() + 2 * kBandClasses;
  }
  return LTR_OK;
}

// ---- job setup, device plan: raw reads up, upper bounds for every list, full persistent grids -------------------------
int setup_device_plan(ltr_ctx* ctx, ltr_job* job, const ltr_viterbi_batch& bb, cudaStream_t st) {
  const int kmax = viterbi_max_rows_per_lane();
  uint64_t* h2d = &job->stats.h2d_bytes;
  const uint32_t n_loci = bb.n_loci, n_haps = job->n_haps, n_reads = job->n_reads;
  const int cut = 35 - job->params.indel_flank_len;
  job->band = band_policy(job->params, ctx->band_w);
  // What the host can know from the locus- and haplotype-level arrays alone (O(n_loci + n_haps)): sizes of the
  // caller-visible matrices, the row classes in use and upper bounds for their lists.
  std::vector<unsigned long long> ll_off((size_t)n_loci + 1, 0);
  uint64_t cap_k[kPlanMaxK + 1] = {0};
  uint32_t maxq_k[kPlanMaxK + 1] = {0};
  bool multi_k[kPlanMaxK + 1] = {false};
  int max_n = 1;
  bool any_band = false;
  for (uint32_t l = 0; l < n_loci; ++l) {
    const uint32_t h0 = bb.locus_hap_begin[l], h1 = bb.locus_hap_begin[l + 1];
    const uint32_t P = bb.locus_read_begin[l + 1] - bb.locus_read_begin[l];
    ll_off[l + 1] = ll_off[l] + (unsigned long long)(h1 - h0) * P;
    if (P == 0) continue;
    uint32_t lbytes = 0;
    bool have_bytes = false;
    for (uint32_t h = h0; h < h1; ++h) {
      const int hlen = (int)(bb.hap_off[h + 1] - bb.hap_off[h]);
      const int n = hlen - 2 * cut;
      int k = 1;
      if (hlen > 60 && n >= 1) {
        max_n = std::max(max_n, n);
        k = rows_per_lane(n, kmax);
        any_band = any_band || (n >= 2);
        const int strips = std::max(1, (n - 1 + 32 * k - 1) / (32 * k));
        if (strips > 1) {
          multi_k[k] = true;
          if (!have_bytes) {  // bytes of the locus' raw reads: upper bound for any read stream of the locus
            const uint32_t o0 = bb.read_off[bb.locus_read_begin[l]], o1 = bb.read_off[bb.locus_read_begin[l + 1]];
            lbytes = o1 > o0 ? o1 - o0 : 0u;
            have_bytes = true;
          }
          maxq_k[k] = std::max(maxq_k[k], lbytes);
        }
      }
      cap_k[k] += P;
    }
  }
  job->n_ll = ll_off[n_loci];
  job->n_upairs_cap = job->n_ll;
  job->stats.n_pairs = job->n_ll;
  if (job->n_ll > 0xFFFFFFF0ull) return LTR_ERR_INVALID;  // pair lists are indexed with 32 bits
  // every pair with |n - m| > 600 is answered before any table is read (HapAligner.cpp:249-252): columns up to
  // max_n + 600 are the most a table entry can be asked for
  make_consts(job->params, max_n + 604, job->hc);
  if (ctx->read_encoding == 1) {
    // the reads travel as one 4-bit stream (half the bytes); the plan's first kernel expands them into raw_bytes_d
    LTR_CUDA(ctx, job->raw_bytes_d.alloc(kStringPad + (size_t)job->raw_bytes + kStringPad));
    job->deferred_zero.push_back(std::make_pair(job->raw_bytes_d.p, kStringPad));
    job->deferred_zero.push_back(std::make_pair((void*)((char*)job->raw_bytes_d.p + kStringPad + job->raw_bytes), kStringPad));
    LTR_TRY(upload(ctx, st, job->raw_packed_d, bb.read_bytes, ((size_t)job->raw_bytes + 1) / 2, 0, h2d));
  } else {
    LTR_TRY(upload(ctx, st, job->raw_bytes_d, bb.read_bytes, (size_t)job->raw_bytes, kStringPad, h2d, &job->deferred_zero));
  }
  LTR_TRY(upload(ctx, st, job->raw_off_d, bb.read_off, (size_t)n_reads + 1, 0, h2d));
  LTR_TRY(upload(ctx, st, job->ll_off, ll_off.data(), ll_off.size(), 0, h2d));
  // products of the plan kernels
  LTR_CUDA(ctx, job->read_bytes.alloc(kStringPad + (size_t)job->raw_bytes + kStringPad));
  job->deferred_zero.push_back(std::make_pair(job->read_bytes.p, kStringPad));
  // (the pad behind the distinct bytes is wherever they end: the whole tail is cleared, it is written again by the fill)
  LTR_CUDA(ctx, job->read_off.alloc(((size_t)n_reads + 2) * 4));
  LTR_CUDA(ctx, job->lub.alloc(((size_t)n_loci + 1) * 4));
  LTR_CUDA(ctx, job->r2u.alloc((size_t)n_reads * 4));
  LTR_CUDA(ctx, job->rlocus.alloc((size_t)n_reads * 4));
  LTR_CUDA(ctx, job->hap_locus.alloc((size_t)n_haps * 4));
  LTR_CUDA(ctx, job->ull_off.alloc(((size_t)n_loci + 1) * 8));
  LTR_CUDA(ctx, job->uniq_ll.alloc(job->n_upairs_cap * sizeof(double)));
  LTR_CUDA(ctx, job->plan_ctl.alloc(PLAN_CTL_WORDS * 4));
  LTR_CUDA(ctx, job->plan_stat.alloc(PLAN_STAT_WORDS * 8));
  // scratch of the plan kernels, one allocation
  const size_t nr = ((size_t)n_reads + 2 + 1) & ~(size_t)1, nl = ((size_t)n_loci + 2 + 1) & ~(size_t)1;
  const size_t n_partial = 3 * ((size_t)kBandClasses * nl / 1024 + 2) + 8;  // tile sums of the scans (plan_kernels.cu)
  LTR_CUDA(ctx, job->plan_scratch.alloc(nr * 8 + n_partial * 8 + nr * 4 * 6 + nl * 4 * 3 + nl * 4 * 2 * kBandClasses));
  PlanDev& P = job->pd;
  P.n_loci = n_loci; P.n_haps = n_haps; P.n_reads = n_reads; P.raw_total = job->raw_bytes;
  P.cut = cut; P.kmax = kmax; P.band = job->band;
  P.lhb = job->lhb.as<uint32_t>(); P.lrb = job->lrb.as<uint32_t>(); P.hap_off = job->hap_off.as<uint32_t>();
  P.read_off = job->raw_off_d.as<uint32_t>();
  P.read_bytes = job->raw_bytes_d.as<uint8_t>() + kStringPad;
  P.packed = (ctx->read_encoding == 1) ? job->raw_packed_d.as<uint8_t>() : nullptr;
  {
    char* s = job->plan_scratch.as<char>();
    P.rhash = reinterpret_cast<unsigned long long*>(s); s += nr * 8;
    P.scan_partial = reinterpret_cast<unsigned long long*>(s); s += n_partial * 8;
    P.rlen = reinterpret_cast<uint32_t*>(s); s += nr * 4;
    P.rep = reinterpret_cast<uint32_t*>(s); s += nr * 4;
    P.rank_of = reinterpret_cast<uint32_t*>(s); s += nr * 4;
    P.tmp_len = reinterpret_cast<uint32_t*>(s); s += nr * 4;
    P.tmp_rep = reinterpret_cast<uint32_t*>(s); s += nr * 4;
    P.local_u = reinterpret_cast<uint32_t*>(s); s += nr * 4;
    P.ucount = reinterpret_cast<uint32_t*>(s); s += nl * 4;
    P.ubytes = reinterpret_cast<uint32_t*>(s); s += nl * 4;
    P.ubyte_off = reinterpret_cast<uint32_t*>(s); s += nl * 4;
    P.band_task_pos = reinterpret_cast<uint32_t*>(s); s += nl * 4 * kBandClasses;  // [class][locus], n_loci <= nl
    P.band_pair_pos = reinterpret_cast<uint32_t*>(s); s += nl * 4 * kBandClasses;
  }
  P.read_locus = job->rlocus.as<uint32_t>(); P.hap_locus = job->hap_locus.as<uint32_t>(); P.lub = job->lub.as<uint32_t>();
  P.ull_off = job->ull_off.as<unsigned long long>(); P.uread_off = job->read_off.as<uint32_t>();
  P.uread_bytes = job->read_bytes.as<uint8_t>() + kStringPad; P.r2u = job->r2u.as<uint32_t>();
  P.ctl = job->plan_ctl.as<uint32_t>(); P.stat = job->plan_stat.as<unsigned long long>();
  for (int k = 0; k <= kPlanMaxK; ++k) {
    P.st_tasks[k] = nullptr;
    P.st_cap[k] = 0;
    P.st_ntasks[k] = nullptr;
  }
  for (int k = 1; k <= kmax; ++k) {
    if (cap_k[k] == 0) continue;
    LTR_TRY(setup_class(ctx, job, k, 0u, cap_k[k], cap_k[k], false, multi_k[k], maxq_k[k]));
    ClassState& cs = job->classes.back();
    P.st_tasks[k] = cs.tasks.as<Task>();
    P.st_cap[k] = cs.task_cap;
    P.st_ntasks[k] = job->cls_ctrl.as<uint32_t>() + 4 * k + 1;
  }
  P.band_tasks = nullptr; P.band_pairs = nullptr; P.band_cap = 0;
  if (job->band.on && any_band && job->n_ll) {
    job->band_cap = (uint32_t)job->n_ll;
    LTR_CUDA(ctx, job->band_tasks.alloc((size_t)job->band_cap * sizeof(BandTask)));
    LTR_CUDA(ctx, job->band_pairs.alloc((size_t)job->band_cap * sizeof(uint2)));
    LTR_CUDA(ctx, job->band_ctrl.alloc(kBandCtrlAlloc));
    if (band_retry_rho() > 0) {
      job->band_retry_cap = job->band_cap;
      LTR_CUDA(ctx, job->band_retry.alloc((size_t)job->band_cap * sizeof(uint2)));
    }
    P.band_tasks = job->band_tasks.as<BandTask>();
    P.band_pairs = reinterpret_cast<PlanPair*>(job->band_pairs.p);
    P.band_cap = job->band_cap;
    for (int c = 0; c < kBandClasses; ++c) {
      ltr_job::BandLaunch bl;
      bl.cls = c;
      bl.grid = (uint32_t)(ctx->sm_count * ctx->band_blocks_per_sm[c]);
      job->band_launches.push_back(bl);
    }
    job->band_info_dev = P.ctl + PLAN_CTL_BAND_INFO;
    job->n_band_tasks_dev = P.ctl + PLAN_CTL_N_BAND_TASKS;
  } else {
    P.band.on = false;
  }
  return LTR_OK;
}

// Everything of a job that happens before its kernels: validation the host can afford, allocations, uploads (h2d stream
// of the job's lane; ev_h2d marks their end).
int job_setup(ltr_ctx* ctx, const ltr_params* params, const ltr_viterbi_batch* b, const ltr_posterior_batch* post,
              ltr_job* job) {
  static const uint32_t kZero[2] = {0, 0};
  ltr_viterbi_batch bb = *b;
  if (bb.n_loci == 0) bb.locus_hap_begin = bb.locus_read_begin = bb.hap_off = bb.read_off = kZero;
  if (!plan_offsets_valid(bb)) return LTR_ERR_INVALID;
  job->params = *params;
  job->n_loci = bb.n_loci;
  job->n_haps = bb.locus_hap_begin[bb.n_loci];
  job->n_reads = bb.locus_read_begin[bb.n_loci];
  job->raw_bytes = job->n_reads ? bb.read_off[job->n_reads] : 0u;
  if ((job->n_haps && bb.hap_off[job->n_haps] && !bb.hap_bytes) || (job->raw_bytes && !bb.read_bytes)) return LTR_ERR_INVALID;
  if (ctx->read_encoding == 1 && !job->device_plan && job->raw_bytes) {
    // small batch, host plan: the 4-bit stream is expanded here (the device plan does it in its first kernel)
    job->host_unpacked.resize(job->raw_bytes);
    for (uint32_t i = 0; i < job->raw_bytes; ++i) {
      const uint8_t byte = bb.read_bytes[i >> 1];
      job->host_unpacked[i] = (uint8_t)"=ACMGRSVTWYHKDBN"[(i & 1u) ? (byte & 15u) : (byte >> 4)];
    }
    bb.read_bytes = job->host_unpacked.data();
  }
  cudaStream_t st = job->st_h2d;
  job->drained = false;
  cudaEvent_t* evs[] = {&job->ev_start, &job->ev_plan, &job->ev_vit, &job->ev_end};
  for (cudaEvent_t* e : evs) LTR_CUDA(ctx, take_event(ctx, true, e));
  LTR_CUDA(ctx, take_event(ctx, false, &job->ev_h2d));
  LTR_CUDA(ctx, take_event(ctx, false, &job->ev_stats));
  LTR_CUDA(ctx, take_event(ctx, false, &job->ev_done));
  job->res = take_result_block(ctx);
  if (!job->res) return LTR_ERR_OOM;
  std::memset(job->res, 0, sizeof(JobResult));
  uint64_t* h2d = &job->stats.h2d_bytes;
  // Uploads that do not depend on the plan are enqueued first: with pinned caller buffers they overlap make_plan.
  LTR_TRY(upload(ctx, st, job->hap_bytes, bb.hap_bytes, (size_t)bb.hap_off[job->n_haps], kStringPad, h2d, &job->deferred_zero));
  LTR_TRY(upload(ctx, st, job->hap_off, bb.hap_off, (size_t)job->n_haps + 1, 0, h2d));
  LTR_TRY(upload(ctx, st, job->lhb, bb.locus_hap_begin, (size_t)job->n_loci + 1, 0, h2d));
  LTR_TRY(upload(ctx, st, job->lrb, bb.locus_read_begin, (size_t)job->n_loci + 1, 0, h2d));
  LTR_CUDA(ctx, job->cls_ctrl.alloc((kPlanMaxK + 1) * 4 * sizeof(uint32_t)));
  if (job->device_plan) LTR_TRY(setup_device_plan(ctx, job, bb, st));
  else LTR_TRY(setup_host_plan(ctx, job, bb, st));
  LTR_TRY(upload(ctx, st, job->tabI, job->hc.tabI.data(), job->hc.tabI.size(), 0, h2d));
  LTR_TRY(upload(ctx, st, job->tabD, job->hc.tabD.data(), job->hc.tabD.size(), 0, h2d));
  LTR_CUDA(ctx, job->out_ll.alloc(job->n_ll * sizeof(double)));
  job->hc.C.tabI = job->tabI.as<double>();
  job->hc.C.tabD = job->tabD.as<double>();
  if (post) LTR_TRY(setup_posteriors(ctx, job, bb, post, st, !job->device_plan));
  LTR_CUDA(ctx, cudaEventRecord(job->ev_h2d, st));
  // pads and scratch lines are cleared on the same stream, behind the copies (never between them: a memset executes on
  // the SMs and would queue behind another job's persistent kernels)
  for (const std::pair<void*, size_t>& z : job->deferred_zero) LTR_CUDA(ctx, cudaMemsetAsync(z.first, 0, z.second, st));
  job->deferred_zero.clear();
  if (!job->device_plan && job->band_cap)
    LTR_CUDA(ctx, launch_band_expand(job->band_tasks.as<BandTask>(), job->band_cum.as<uint32_t>(), job->band_cap,
                                     job->band_pairs.as<uint2>(), st));
  return LTR_OK;
}

// the [class][locus] band counters of the device plan (band_task_pos and band_pair_pos are adjacent in the scratch)
size_t plan_band_counter_bytes(const ltr_job* job) {
  const size_t nl = ((size_t)job->n_loci + 2 + 1) & ~(size_t)1;
  return nl * 4 * 2 * kBandClasses;
}

DevBatch job_dev_batch(const ltr_job* job) {
  DevBatch B;
  B.hap_bytes = job->hap_bytes.as<uint8_t>() + kStringPad;
  B.hap_off = job->hap_off.as<uint32_t>();
  B.hap_locus = job->hap_locus.as<uint32_t>();
  B.read_bytes = job->read_bytes.as<uint8_t>() + kStringPad;
  B.read_off = job->read_off.as<uint32_t>();
  B.locus_hap_begin = job->lhb.as<uint32_t>();
  B.locus_read_begin = job->lub.as<uint32_t>();            // distinct reads
  B.ll_off = job->ull_off.as<unsigned long long>();
  B.out_ll = job->uniq_ll.as<double>();
  return B;
}

// Launches the stream kernels of every row class on its stream (control words were initialised on the main stream).
int run_classes(ltr_ctx* ctx, ltr_job* job, JobLane& L) {
  const DevBatch B = job_dev_batch(job);
  int si = 0;
  for (size_t ci = 0; ci < job->classes.size(); ++ci) {
    ClassState& cs = job->classes[ci];
    static const bool serial = getenv("LTR_SERIAL_CLASSES") != nullptr;  // diagnostics: one row class at a time
    cudaStream_t st = L.cls[serial ? 0 : (si % kNumStreams)];
    ++si;
    uint32_t* ctrl = job->cls_ctrl.as<uint32_t>() + 4 * cs.k;
    FailSink sink;
    sink.items = cs.fails.as<Task>();
    sink.count = ctrl + 2;
    sink.capacity = cs.fail_cap;
    FailSink none;
    none.items = nullptr;
    none.count = ctrl + 2;
    none.capacity = 0;
    if (cs.force_full) {
      // parameters outside the certificate's condition: exact kernel only
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
int run_band_phase(ltr_ctx* ctx, ltr_job* job, JobLane& L) {
  const DevBatch B = job_dev_batch(job);
  LTR_CUDA(ctx, cudaMemsetAsync(job->band_ctrl.p, 0, kBandCtrlAlloc, L.main));
  LTR_CUDA(ctx, cudaEventRecord(L.ev_init, L.main));
  uint32_t* bctrl = job->band_ctrl.as<uint32_t>();
  static const bool no_abandon = getenv("LTR_BAND_NO_ABANDON") != nullptr;  // diagnostics
  for (size_t i = 0; i < job->band_launches.size(); ++i) {
    const ltr_job::BandLaunch& bl = job->band_launches[i];
    cudaStream_t st = L.cls[i % kNumStreams];
    LTR_CUDA(ctx, cudaStreamWaitEvent(st, L.ev_init, 0));
    BandArgs A;
    A.pairs = job->band_pairs.as<uint2>();
    A.info = job->band_info_dev + 2 * bl.cls;
    A.cursor = bctrl + bl.cls;
    A.counters = bctrl + 12;
    A.cells_evaluated = reinterpret_cast<unsigned long long*>(job->band_ctrl.as<char>() + kBandStatOff + 16);
    A.gap = job->band.gap;
    A.abandon_after = no_abandon ? 0u : 4096u;
    LTR_CUDA(ctx, launch_band(bl.cls, (int)bl.grid, st, job->hc.C, B, A));
    job->stats.n_launches += 1;
    LTR_CUDA(ctx, cudaEventRecord(L.ev_cls[i % kNumStreams], st));
    LTR_CUDA(ctx, cudaStreamWaitEvent(L.main, L.ev_cls[i % kNumStreams], 0));
  }
  BandCollect S;
  for (int k = 0; k < 17; ++k) {
    S.tasks[k] = nullptr;
    S.count[k] = bctrl + 15;  // never read: capacity 0
    S.cap[k] = 0;
  }
  for (ClassState& cs : job->classes) {
    S.tasks[cs.k] = cs.tasks.as<Task>();
    S.count[cs.k] = job->cls_ctrl.as<uint32_t>() + 4 * cs.k + 1;
    S.cap[cs.k] = cs.task_cap;
  }
  S.kmax = viterbi_max_rows_per_lane();
  S.n_uncertified = reinterpret_cast<unsigned long long*>(job->band_ctrl.as<char>() + kBandStatOff);
  S.cells_uncertified = reinterpret_cast<unsigned long long*>(job->band_ctrl.as<char>() + kBandStatOff + 8);
  S.bucket_count = reinterpret_cast<uint32_t*>(job->band_ctrl.as<char>() + kBandBucketOff);
  S.bucket_base = S.bucket_count + kBandBucketWords;
  S.bucket_fill = S.bucket_base + kBandBucketWords;
  uint32_t* rctrl = reinterpret_cast<uint32_t*>(job->band_ctrl.as<char>() + kBandRetryOff);
  S.retry_pairs = job->band_retry_cap ? job->band_retry.as<uint2>() : nullptr;
  S.retry_cap = job->band_retry_cap;
  S.retry_count = rctrl;
  S.retry_fill = rctrl + 16;
  S.retry_info = rctrl + 32;
  S.n_retried = reinterpret_cast<unsigned long long*>(job->band_ctrl.as<char>() + kBandStatOff + 24);
  S.n_retry_failed = reinterpret_cast<unsigned long long*>(job->band_ctrl.as<char>() + kBandStatOff + 32);
  S.gap = job->band.gap;
  S.retry_rho_pct = band_retry_rho();
  LTR_CUDA(ctx, launch_band_collect(job->hc.C, B, job->band_tasks.as<BandTask>(), job->n_band_tasks_dev, job->band_cap, S,
                                    ctx->sm_count, L.main));
  job->stats.n_launches += 3;
  if (S.retry_pairs) {
    // second band round: the pairs band_retry_class found a wider, still worthwhile class for (certain to be certified)
    LTR_CUDA(ctx, cudaEventRecord(L.ev_init, L.main));
    int used = 0;
    for (int cls = 1; cls < kBandClasses; ++cls) {  // nothing is retried into the narrowest class
      ltr_job::BandLaunch bl;
      bl.cls = cls;
      cudaStream_t st = L.cls[used % kNumStreams];
      LTR_CUDA(ctx, cudaStreamWaitEvent(st, L.ev_init, 0));
      BandArgs A;
      A.pairs = job->band_retry.as<uint2>();
      A.info = S.retry_info + 2 * bl.cls;
      A.cursor = rctrl + 64 + bl.cls;
      A.counters = bctrl + 14;  // scratch words: the statistics of the first round stay as they are
      A.cells_evaluated = reinterpret_cast<unsigned long long*>(job->band_ctrl.as<char>() + kBandStatOff + 16);
      A.gap = job->band.gap;
      A.abandon_after = 0u;
      const uint32_t full_grid = (uint32_t)(ctx->sm_count * ctx->band_blocks_per_sm[bl.cls]);
      LTR_CUDA(ctx, launch_band(bl.cls, (int)full_grid, st, job->hc.C, B, A));
      job->stats.n_launches += 1;
      LTR_CUDA(ctx, cudaEventRecord(L.ev_cls[used % kNumStreams], st));
      LTR_CUDA(ctx, cudaStreamWaitEvent(L.main, L.ev_cls[used % kNumStreams], 0));
      ++used;
    }
    LTR_CUDA(ctx, launch_band_retry_check(job->hc.C, B, S, ctx->sm_count, L.main));
    job->stats.n_launches += 1;
  }
  LTR_CUDA(ctx, cudaEventRecord(L.ev_collect, L.main));
  return LTR_OK;
}

// Enqueues every kernel of the job and the copy of its statistics; returns without waiting.
int job_enqueue_compute(ltr_ctx* ctx, ltr_job* job) {
  JobLane& L = ctx->lanes[job->lane];
  job->stats.n_launches = 0;
  job->stats.n_fallback = 0;
  job->stats.n_band_uncertified = 0;
  if (job->st_h2d != L.main) {
    // (ev_h2d only covers the copies; the clears behind it on the h2d stream are ordered with a second event)
    LTR_CUDA(ctx, cudaEventRecord(L.ev_init, job->st_h2d));
    LTR_CUDA(ctx, cudaStreamWaitEvent(L.main, L.ev_init, 0));
  }
  job->drained = false;
  LTR_CUDA(ctx, cudaEventRecord(job->ev_start, L.main));
  // a kernel that silently skipped work must not go unnoticed: results start out as NaN
  if (job->n_upairs_cap) LTR_CUDA(ctx, cudaMemsetAsync(job->uniq_ll.p, 0xFF, job->n_upairs_cap * sizeof(double), L.main));
  if (job->n_ll) LTR_CUDA(ctx, cudaMemsetAsync(job->out_ll.p, 0xFF, job->n_ll * sizeof(double), L.main));
  {
    uint32_t* init = job->res->cls_ctrl;  // pinned; the values are consumed before the statistics overwrite them
    std::memset(init, 0, sizeof(job->res->cls_ctrl));
    for (const ClassState& cs : job->classes) init[4 * cs.k + 1] = cs.n_tasks;
    LTR_CUDA(ctx, cudaMemcpyAsync(job->cls_ctrl.p, init, sizeof(job->res->cls_ctrl), cudaMemcpyHostToDevice, L.main));
  }
  if (job->device_plan) {
    LTR_CUDA(ctx, cudaMemsetAsync(job->plan_ctl.p, 0, PLAN_CTL_WORDS * 4, L.main));
    LTR_CUDA(ctx, cudaMemsetAsync(job->plan_stat.p, 0, PLAN_STAT_WORDS * 8, L.main));
    LTR_CUDA(ctx, cudaMemsetAsync(job->pd.band_task_pos, 0, plan_band_counter_bytes(job), L.main));
    LTR_CUDA(ctx, launch_device_plan(job->pd, ctx->sm_count, L.main));
    if (job->n_loci) job->stats.n_launches += 11;
  }
  LTR_CUDA(ctx, cudaEventRecord(job->ev_plan, L.main));
  const int n_used = (int)std::min<size_t>(kNumStreams, job->classes.size());  // streams run_classes touches
  const bool band = !job->band_launches.empty();
  if (band) {
    LTR_TRY(run_band_phase(ctx, job, L));
    for (int i = 0; i < n_used; ++i) LTR_CUDA(ctx, cudaStreamWaitEvent(L.cls[i], L.ev_collect, 0));
  } else {
    LTR_CUDA(ctx, cudaEventRecord(L.ev_collect, L.main));
    for (int i = 0; i < n_used; ++i) LTR_CUDA(ctx, cudaStreamWaitEvent(L.cls[i], L.ev_collect, 0));
  }
  LTR_TRY(run_classes(ctx, job, L));
  for (int i = 0; i < n_used; ++i) {
    LTR_CUDA(ctx, cudaEventRecord(L.ev_cls[i], L.cls[i]));
    LTR_CUDA(ctx, cudaStreamWaitEvent(L.main, L.ev_cls[i], 0));
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
    E.err = job->device_plan ? job->plan_ctl.as<uint32_t>() + PLAN_CTL_ERR : nullptr;
    LTR_CUDA(ctx, launch_expand_ll(E, L.main));
    if (job->n_reads) job->stats.n_launches += 1;
  }
  LTR_CUDA(ctx, cudaEventRecord(job->ev_vit, L.main));
  if (job->has_post) {
    DevPosterior P;
    P.n_loci = job->n_loci;
    P.locus_hap_begin = job->lhb.as<uint32_t>();
    P.locus_read_begin = job->lrb.as<uint32_t>();
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
    P.err = job->device_plan ? job->plan_ctl.as<uint32_t>() + PLAN_CTL_ERR : nullptr;
    P.second_mate = job->mate.p ? job->mate.as<uint8_t>() : nullptr;
    P.read_aligned = job->aligned.p ? job->aligned.as<uint8_t>() : nullptr;
    P.kept_mask = job->prune ? job->kept_mask.as<uint8_t>() : nullptr;
    P.kept_index = job->prune ? job->kept_index.as<uint32_t>() : nullptr;
    if (job->prune && job->n_haps) LTR_CUDA(ctx, cudaMemsetAsync(job->kept_mask.p, 1, job->n_haps, L.main));
    if (job->device_plan) {
      LTR_CUDA(ctx, launch_posterior_validate(P, job->plan_ctl.as<uint32_t>() + PLAN_CTL_ERR, L.main));
      job->stats.n_launches += 1;
    }
    LTR_CUDA(ctx, launch_posteriors(P, L.main));
    job->stats.n_launches += 1;
  }
  LTR_CUDA(ctx, cudaEventRecord(job->ev_end, L.main));
  // statistics and error flags travel with the results
  LTR_CUDA(ctx, cudaMemcpyAsync(job->res->cls_ctrl, job->cls_ctrl.p, sizeof(job->res->cls_ctrl), cudaMemcpyDeviceToHost, L.main));
  if (band)
    LTR_CUDA(ctx, cudaMemcpyAsync(job->res->band_words, job->band_ctrl.p, kBandCtrlBytes, cudaMemcpyDeviceToHost, L.main));
  if (job->device_plan) {
    LTR_CUDA(ctx, cudaMemcpyAsync(job->res->plan_ctl, job->plan_ctl.p, sizeof(job->res->plan_ctl), cudaMemcpyDeviceToHost, L.main));
    LTR_CUDA(ctx, cudaMemcpyAsync(job->res->plan_stat, job->plan_stat.p, sizeof(job->res->plan_stat), cudaMemcpyDeviceToHost, L.main));
  }
  LTR_CUDA(ctx, cudaEventRecord(job->ev_stats, L.main));
  job->ran = true;
  return LTR_OK;
}

// After ev_stats (or ev_done) has been waited for: statistics, error flags.
int job_collect(ltr_ctx* ctx, ltr_job* job) {
  const JobResult& R = *job->res;
  const bool band = !job->band_launches.empty();
  if (job->device_plan) {
    if (R.plan_ctl[PLAN_CTL_ERR] != 0) return LTR_ERR_INVALID;
    job->stats.n_cells = R.plan_stat[PLAN_STAT_CELLS];
    job->stats.n_pairs_computed = R.plan_stat[PLAN_STAT_PAIRS_COMPUTED];
    job->plan_cells_computed = R.plan_stat[PLAN_STAT_CELLS_STREAM];
    job->stats.n_band_pairs = R.plan_ctl[PLAN_CTL_N_BAND_PAIRS];
  }
  job->stats.n_band_uncertified = band ? R.band_words[8] : 0;
  job->stats.n_band_retried = band ? R.band_words[11] : 0;
  if (band && R.band_words[12] != 0) ctx->last_error = "band retry: a pair was not certified by its second round (re-run over the full matrix)";
  // cells evaluated: full matrices of the stream-kernel pairs (planned + uncertified) + the bands actually evaluated
  job->stats.n_cells_computed = job->plan_cells_computed + (band ? R.band_words[9] + R.band_words[10] : 0);
  job->stats.n_fallback = 0;
  for (const ClassState& cs : job->classes)
    if (!cs.force_full) job->stats.n_fallback += R.cls_ctrl[4 * cs.k + 2];
  float ms_total = 0.f, ms_vit = 0.f, ms_plan = 0.f;
  LTR_CUDA(ctx, cudaEventElapsedTime(&ms_total, job->ev_start, job->ev_end));
  LTR_CUDA(ctx, cudaEventElapsedTime(&ms_vit, job->ev_plan, job->ev_vit));
  LTR_CUDA(ctx, cudaEventElapsedTime(&ms_plan, job->ev_start, job->ev_plan));
  job->stats.kernel_ms = ms_total;
  job->stats.viterbi_ms = ms_vit;
  job->stats.plan_ms = ms_plan;
  return LTR_OK;
}

int job_new(ltr_ctx* ctx, const ltr_params* params, const ltr_viterbi_batch* b, const ltr_posterior_batch* post,
            ltr_job** out) {
  if (!ctx || !params || !b || !out) return LTR_ERR_INVALID;
  *out = nullptr;
  if (b->n_loci && (!b->locus_hap_begin || !b->locus_read_begin || !b->hap_off || !b->read_off))
    return LTR_ERR_INVALID;
  if (params->indel_flank_len < 0 || params->indel_flank_len > 35) return LTR_ERR_INVALID;
  // below 5 the reference's substr(cut, size - 2 cut) underflows for haplotypes of 61 .. 2 (35 - flank) bases
  // (HapAligner.cpp:245-246) and aligns against hap.substr(cut): not reproduced
  if (params->indel_flank_len < 5) return LTR_ERR_UNSUPPORTED;
  LTR_CUDA(ctx, cudaSetDevice(ctx->device));
  ltr_job* job = new ltr_job();
  std::memset(&job->stats, 0, sizeof(job->stats));
  job->ctx = ctx;
  job->lane = (int)(ctx->next_lane++ % (unsigned)kLanes);
  // Large batches are planned on the device and use the lane's copy streams so that uploads and downloads overlap other
  // jobs' kernels; the small batches of the per-locus entry points are planned on the host and keep everything on the
  // lane's main stream (fewer cross-stream waits per call).
  const uint32_t n_reads_hint = (b->n_loci && b->locus_read_begin) ? b->locus_read_begin[b->n_loci] : 0u;
  job->device_plan = ctx->plan_mode == 2 || (ctx->plan_mode == 0 && n_reads_hint >= kDevicePlanMinReads);
  JobLane& lane = ctx->lanes[job->lane];
  job->st_h2d = job->device_plan ? lane.h2d : lane.main;
  job->st_d2h = job->device_plan ? lane.d2h : lane.main;
  AllocScope alloc_scope(job->st_h2d);
  const int rc = job_setup(ctx, params, b, post, job);
  if (rc != LTR_OK) {
    ltr_job_destroy(ctx, job);  // drains the lane's copies first: the caller may release its arrays
    return rc;
  }
  *out = job;
  return LTR_OK;
}

}  // namespace

extern "C" {

int ltr_job_create(ltr_ctx* ctx, const ltr_params* params, const ltr_viterbi_batch* b,
                   const ltr_posterior_batch* post, ltr_job** out) {
  static const bool timing = getenv("LTR_TIMING") != nullptr;  // diagnostics: host-side phases of job creation on stderr
  const auto t_begin = std::chrono::steady_clock::now();
  ltr_job* job = nullptr;
  const int rc = job_new(ctx, params, b, post, &job);
  if (rc != LTR_OK) return rc;
  const auto t_enq = std::chrono::steady_clock::now();
  JobLane& L = ctx->lanes[job->lane];
  int rc2 = LTR_OK;
  if (job->device_plan) {
    // the batch is validated here, once, by running the plan: a resident job is then known to be well formed
    cudaError_t e = cudaStreamWaitEvent(L.main, job->ev_h2d, 0);
    if (e == cudaSuccess) e = cudaMemsetAsync(job->plan_ctl.p, 0, PLAN_CTL_WORDS * 4, L.main);
    if (e == cudaSuccess) e = cudaMemsetAsync(job->plan_stat.p, 0, PLAN_STAT_WORDS * 8, L.main);
    if (e == cudaSuccess) e = cudaMemsetAsync(job->pd.band_task_pos, 0, plan_band_counter_bytes(job), L.main);
    if (e == cudaSuccess) e = launch_device_plan(job->pd, ctx->sm_count, L.main);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(job->res->plan_ctl, job->plan_ctl.p, sizeof(job->res->plan_ctl), cudaMemcpyDeviceToHost, L.main);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(job->res->plan_stat, job->plan_stat.p, sizeof(job->res->plan_stat), cudaMemcpyDeviceToHost, L.main);
    if (e == cudaSuccess) e = cudaStreamSynchronize(L.main);
    if (e != cudaSuccess) rc2 = fail_cuda(ctx, e, "ltr_job_create: device plan");
    else if (job->res->plan_ctl[PLAN_CTL_ERR] != 0) rc2 = LTR_ERR_INVALID;
    else {
      job->stats.n_cells = job->res->plan_stat[PLAN_STAT_CELLS];
      job->stats.n_pairs_computed = job->res->plan_stat[PLAN_STAT_PAIRS_COMPUTED];
      job->stats.n_band_pairs = job->res->plan_ctl[PLAN_CTL_N_BAND_PAIRS];
    }
  }
  // every copy from host memory has to be finished: the caller may release its arrays when we return
  if (rc2 == LTR_OK) {
    cudaError_t e = cudaStreamSynchronize(job->st_h2d);
    if (e != cudaSuccess) rc2 = fail_cuda(ctx, e, "ltr_job_create: upload");
  }
  if (rc2 != LTR_OK) {
    ltr_job_destroy(ctx, job);
    return rc2;
  }
  job->drained = true;
  if (timing) {
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point c) {
      return std::chrono::duration<double, std::milli>(c - a).count();
    };
    fprintf(stderr, "[ltr] job_create (%s plan): setup + enqueue %.1f ms, drain %.1f ms (h2d %.1f MB)\n",
            job->device_plan ? "device" : "host", ms(t_begin, t_enq), ms(t_enq, std::chrono::steady_clock::now()),
            job->stats.h2d_bytes / 1e6);
  }
  *out = job;
  return LTR_OK;
}

int ltr_job_run(ltr_ctx* ctx, ltr_job* job) {
  if (!ctx || !job) return LTR_ERR_INVALID;
  LTR_CUDA(ctx, cudaSetDevice(ctx->device));
  LTR_TRY(job_enqueue_compute(ctx, job));
  LTR_CUDA(ctx, cudaEventSynchronize(job->ev_stats));
  job->drained = true;  // the uploads were awaited by ltr_job_create, every kernel and the statistics copy by this
  return job_collect(ctx, job);
}

int ltr_job_submit_outputs(ltr_ctx* ctx, const ltr_params* params, const ltr_viterbi_batch* batch,
                           const ltr_posterior_batch* post, const ltr_job_outputs* outputs, ltr_job** out) {
  static const ltr_job_outputs kNone = {nullptr, nullptr, nullptr, nullptr};
  const ltr_job_outputs& O = outputs ? *outputs : kNone;
  ltr_job* job = nullptr;
  int rc = job_new(ctx, params, batch, post, &job);
  if (rc != LTR_OK) return rc;
  JobLane& L = ctx->lanes[job->lane];
  rc = job_enqueue_compute(ctx, job);
  if (rc == LTR_OK) {
    cudaStream_t sd = job->st_d2h;
    cudaError_t e = (sd != L.main) ? cudaStreamWaitEvent(sd, job->ev_stats, 0) : cudaSuccess;
    job->stats.d2h_bytes = 0;
    if (e == cudaSuccess && O.ll && job->n_ll) {
      e = cudaMemcpyAsync(O.ll, job->out_ll.p, job->n_ll * sizeof(double), cudaMemcpyDeviceToHost, sd);
      job->stats.d2h_bytes += job->n_ll * sizeof(double);
    }
    if (e == cudaSuccess && O.post && job->has_post && job->n_post) {
      e = cudaMemcpyAsync(O.post, job->post.p, job->n_post * sizeof(double), cudaMemcpyDeviceToHost, sd);
      job->stats.d2h_bytes += job->n_post * sizeof(double);
    }
    if (e == cudaSuccess && O.totals && job->has_post && job->n_tot) {
      e = cudaMemcpyAsync(O.totals, job->totals.p, job->n_tot * sizeof(double), cudaMemcpyDeviceToHost, sd);
      job->stats.d2h_bytes += job->n_tot * sizeof(double);
    }
    if (e == cudaSuccess && O.kept_mask && job->n_haps) {
      if (job->prune) {
        e = cudaMemcpyAsync(O.kept_mask, job->kept_mask.p, job->n_haps, cudaMemcpyDeviceToHost, sd);
        job->stats.d2h_bytes += job->n_haps;
      } else {
        std::memset(O.kept_mask, 1, job->n_haps);
      }
    }
    if (e == cudaSuccess) e = cudaEventRecord(job->ev_done, sd);
    if (e != cudaSuccess) rc = fail_cuda(ctx, e, "ltr_job_submit: download");
  }
  if (rc == LTR_OK && !job->device_plan) {
    // host plan: its arrays sit in the context's staging buffers, which the next job will overwrite
    cudaError_t e = cudaEventSynchronize(job->ev_h2d);
    if (e != cudaSuccess) rc = fail_cuda(ctx, e, "ltr_job_submit: upload");
  }
  job->pending = true;
  if (rc != LTR_OK) {
    ltr_job_destroy(ctx, job);
    return rc;
  }
  *out = job;
  return LTR_OK;
}

int ltr_job_submit(ltr_ctx* ctx, const ltr_params* params, const ltr_viterbi_batch* batch,
                   const ltr_posterior_batch* post, double* out_ll, double* out_post, double* out_totals,
                   ltr_job** out) {
  ltr_job_outputs O;
  O.ll = out_ll;
  O.post = out_post;
  O.totals = out_totals;
  O.kept_mask = nullptr;
  return ltr_job_submit_outputs(ctx, params, batch, post, &O, out);
}

int ltr_job_wait(ltr_ctx* ctx, ltr_job* job) {
  if (!ctx || !job) return LTR_ERR_INVALID;
  if (!job->pending) return LTR_ERR_INVALID;
  LTR_CUDA(ctx, cudaSetDevice(ctx->device));
  LTR_CUDA(ctx, cudaEventSynchronize(job->ev_done));
  job->pending = false;
  job->drained = true;
  return job_collect(ctx, job);
}

int ltr_job_poll(ltr_ctx* ctx, ltr_job* job) {
  if (!ctx || !job || !job->pending) return LTR_ERR_INVALID;
  const cudaError_t e = cudaEventQuery(job->ev_done);
  if (e == cudaSuccess) return 1;
  if (e == cudaErrorNotReady) return 0;
  return fail_cuda(ctx, e, "ltr_job_poll");
}

void ltr_job_sizes(const ltr_job* job, uint64_t* n_ll, uint64_t* n_post, uint64_t* n_totals) {
  if (n_ll) *n_ll = job ? job->n_ll : 0;
  if (n_post) *n_post = job ? job->n_post : 0;
  if (n_totals) *n_totals = job ? job->n_tot : 0;
}

int ltr_job_download(ltr_ctx* ctx, ltr_job* job, double* out_ll, double* out_post, double* out_totals) {
  if (!ctx || !job) return LTR_ERR_INVALID;
  LTR_CUDA(ctx, cudaSetDevice(ctx->device));
  JobLane& L = ctx->lanes[job->lane];
  job->stats.d2h_bytes = 0;
  if (out_ll && job->n_ll) {
    LTR_CUDA(ctx, cudaMemcpyAsync(out_ll, job->out_ll.p, job->n_ll * sizeof(double), cudaMemcpyDeviceToHost, L.main));
    job->stats.d2h_bytes += job->n_ll * sizeof(double);
  }
  if (out_post && job->has_post && job->n_post) {
    LTR_CUDA(ctx, cudaMemcpyAsync(out_post, job->post.p, job->n_post * sizeof(double), cudaMemcpyDeviceToHost, L.main));
    job->stats.d2h_bytes += job->n_post * sizeof(double);
  }
  if (out_totals && job->has_post && job->n_tot) {
    LTR_CUDA(ctx, cudaMemcpyAsync(out_totals, job->totals.p, job->n_tot * sizeof(double), cudaMemcpyDeviceToHost, L.main));
    job->stats.d2h_bytes += job->n_tot * sizeof(double);
  }
  LTR_CUDA(ctx, cudaStreamSynchronize(L.main));
  return LTR_OK;
}

int ltr_job_download_kept(ltr_ctx* ctx, ltr_job* job, uint8_t* out_kept_mask) {
  if (!ctx || !job || !out_kept_mask) return LTR_ERR_INVALID;
  LTR_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!job->prune) {
    std::memset(out_kept_mask, 1, job->n_haps);
    return LTR_OK;
  }
  JobLane& L = ctx->lanes[job->lane];
  if (job->n_haps) LTR_CUDA(ctx, cudaMemcpyAsync(out_kept_mask, job->kept_mask.p, job->n_haps, cudaMemcpyDeviceToHost, L.main));
  LTR_CUDA(ctx, cudaStreamSynchronize(L.main));
  return LTR_OK;
}

// Posterior stage on LL matrices the caller already holds (e.g. from ltr_stutter_ll): one upload, one launch, one download.
int ltr_posteriors_batch(ltr_ctx* ctx, uint32_t n_loci, const uint32_t* locus_hap_begin, const uint32_t* locus_read_begin,
                         const double* ll, const ltr_posterior_batch* post, double* out_post, double* out_totals,
                         uint8_t* out_kept_mask) {
  if (!ctx || !post || !out_post || !out_totals) return LTR_ERR_INVALID;
  if (n_loci == 0) return LTR_OK;
  if (!locus_hap_begin || !locus_read_begin || !ll) return LTR_ERR_INVALID;
  LTR_CUDA(ctx, cudaSetDevice(ctx->device));
  ltr_job* job = new ltr_job();
  std::memset(&job->stats, 0, sizeof(job->stats));
  job->ctx = ctx;
  job->lane = (int)(ctx->next_lane++ % (unsigned)kLanes);
  JobLane& L = ctx->lanes[job->lane];
  AllocScope alloc_scope(L.h2d);
  ltr_viterbi_batch bb;
  std::memset(&bb, 0, sizeof(bb));
  bb.n_loci = n_loci;
  bb.locus_hap_begin = locus_hap_begin;
  bb.locus_read_begin = locus_read_begin;
  int rc = LTR_OK;
  std::vector<unsigned long long> ll_off((size_t)n_loci + 1, 0);
  for (uint32_t l = 0; l < n_loci && rc == LTR_OK; ++l) {
    if (locus_hap_begin[l + 1] < locus_hap_begin[l] || locus_read_begin[l + 1] < locus_read_begin[l]) rc = LTR_ERR_INVALID;
    ll_off[l + 1] = ll_off[l] + (unsigned long long)(locus_hap_begin[l + 1] - locus_hap_begin[l]) *
                                    (locus_read_begin[l + 1] - locus_read_begin[l]);
  }
  job->n_loci = n_loci;
  job->n_haps = locus_hap_begin[n_loci];
  job->n_ll = ll_off[n_loci];
  uint64_t* h2d = &job->stats.h2d_bytes;
  if (rc == LTR_OK) rc = upload(ctx, L.h2d, job->lhb, locus_hap_begin, (size_t)n_loci + 1, 0, h2d);
  if (rc == LTR_OK) rc = upload(ctx, L.h2d, job->lrb, locus_read_begin, (size_t)n_loci + 1, 0, h2d);
  if (rc == LTR_OK) rc = upload(ctx, L.h2d, job->ll_off, ll_off.data(), ll_off.size(), 0, h2d);
  if (rc == LTR_OK) rc = upload(ctx, L.h2d, job->out_ll, ll, (size_t)job->n_ll, 0, h2d);
  if (rc == LTR_OK) rc = setup_posteriors(ctx, job, bb, post, L.h2d, true);
  if (rc == LTR_OK) {
    DevPosterior P;
    P.n_loci = n_loci;
    P.locus_hap_begin = job->lhb.as<uint32_t>();
    P.locus_read_begin = job->lrb.as<uint32_t>();
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
    P.err = nullptr;
    P.second_mate = job->mate.p ? job->mate.as<uint8_t>() : nullptr;
    P.read_aligned = job->aligned.p ? job->aligned.as<uint8_t>() : nullptr;
    P.kept_mask = job->prune ? job->kept_mask.as<uint8_t>() : nullptr;
    P.kept_index = job->prune ? job->kept_index.as<uint32_t>() : nullptr;
    cudaError_t e = cudaSuccess;
    if (job->prune && job->n_haps) e = cudaMemsetAsync(job->kept_mask.p, 1, job->n_haps, L.h2d);
    if (e == cudaSuccess) e = launch_posteriors(P, L.h2d);
    if (e == cudaSuccess && job->n_post) e = cudaMemcpyAsync(out_post, job->post.p, job->n_post * 8, cudaMemcpyDeviceToHost, L.h2d);
    if (e == cudaSuccess && job->n_tot) e = cudaMemcpyAsync(out_totals, job->totals.p, job->n_tot * 8, cudaMemcpyDeviceToHost, L.h2d);
    if (e == cudaSuccess && out_kept_mask && job->n_haps) {
      if (job->prune) e = cudaMemcpyAsync(out_kept_mask, job->kept_mask.p, job->n_haps, cudaMemcpyDeviceToHost, L.h2d);
      else std::memset(out_kept_mask, 1, job->n_haps);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(L.h2d);
    if (e != cudaSuccess) rc = fail_cuda(ctx, e, "ltr_posteriors_batch");
  }
  ltr_job_destroy(ctx, job);
  return rc;
}

void ltr_job_get_stats(const ltr_job* job, ltr_job_stats* stats) {
  if (job && stats) *stats = job->stats;
}

int ltr_viterbi_ll(ltr_ctx* ctx, const ltr_params* params, const ltr_viterbi_batch* batch, double* out_ll,
                   ltr_job_stats* stats) {
  ltr_job* job = nullptr;
  int rc = ltr_job_submit(ctx, params, batch, nullptr, out_ll, nullptr, nullptr, &job);
  if (rc != LTR_OK) return rc;
  rc = ltr_job_wait(ctx, job);
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
  DeviceBuffer d_ll, d_p1, d_p2, d_lab, d_pool, d_lhb, d_lrb, d_lsb, d_ns, d_hap, d_off, d_post, d_tot, d_logs;
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
  up(d_pool, pool.data(), R * 4); up(d_lhb, lhb, 8); up(d_lrb, lsb, 8); up(d_lsb, lsb, 8); up(d_ns, ns, 4); up(d_hap, hp, 1);
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
    P.locus_read_begin = d_lrb.as<uint32_t>();
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
    P.err = nullptr;
    P.second_mate = nullptr;
    P.read_aligned = nullptr;
    P.kept_mask = nullptr;
    P.kept_index = nullptr;
    cudaError_t e = launch_posteriors(P, ctx->main_stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(post, d_post.p, S * H * H * 8, cudaMemcpyDeviceToHost, ctx->main_stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(totals, d_tot.p, S * 8, cudaMemcpyDeviceToHost, ctx->main_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->main_stream);
    if (e != cudaSuccess) rc = fail_cuda(ctx, e, "ltr_posteriors run");
  }
  DeviceBuffer* all[] = {&d_ll, &d_p1, &d_p2, &d_lab, &d_pool, &d_lhb, &d_lrb, &d_lsb, &d_ns, &d_hap, &d_off, &d_post,
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
