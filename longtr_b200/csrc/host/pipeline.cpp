// pipeline.cpp -- ltr_pipeline_*: many loci in flight for a host that produces loci one at a time.
//
// LongTR's region loop (src/genotyper_bam_processor.cpp:227-351, src/bam_processor.cpp:536-628) decodes one region,
// builds its candidate haplotypes and calls HapAligner::process_reads before it looks at the next region.  The
// pipeline lets that loop hand each locus over (deep copy of the flat locus, include/longtr_b200_locus.h) and carry
// on: loci are collected into batches of `batch_loci` and run through ltr_process_reads_flat_batch on `slots` worker
// threads, each with its own ltr_ctx, so that decoding the next regions overlaps the plan / upload / kernels of the
// previous ones.  Results are handed back by tag, in completion order.  No CPU fallback: a worker whose context
// cannot be created fails its batches with LTR_ERR_NO_DEVICE.
#include <stdlib.h>
#include <string.h>

#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "longtr_b200.h"

namespace {

struct OwnedLocus {
  ltr_flat_locus L;
  uint64_t tag = 0;
  int status = LTR_OK;
  std::string lflank, rflank, motif;
  std::vector<std::string> alleles, seq, qual, cigar;
  std::vector<const char*> allele_ptrs;
  std::vector<ltr_flat_read> reads;
  std::vector<uint8_t> realign_hap, realign_read;
  std::vector<double> ll;
  std::vector<int32_t> seeds;
};

// Deep copy; returns NULL for a locus the flat API would reject for missing pointers (it validates the rest).
OwnedLocus* copy_locus(const ltr_flat_locus* src, uint64_t tag, const int32_t* in_seeds, double fill) {
  if (!src || !src->lflank || !src->rflank || !src->motif || !src->alleles || src->n_alleles < 1 || src->n_reads < 0) return nullptr;
  if (src->n_reads > 0 && !src->reads) return nullptr;
  std::unique_ptr<OwnedLocus> o(new OwnedLocus());
  o->tag = tag;
  o->L = *src;
  o->lflank = src->lflank;
  o->rflank = src->rflank;
  o->motif = src->motif;
  o->alleles.reserve((size_t)src->n_alleles);
  for (int a = 0; a < src->n_alleles; ++a) {
    if (!src->alleles[a]) return nullptr;
    o->alleles.push_back(src->alleles[a]);
  }
  const size_t nr = (size_t)src->n_reads;
  o->seq.reserve(nr);
  o->qual.reserve(nr);
  o->cigar.reserve(nr);
  for (size_t r = 0; r < nr; ++r) {
    const ltr_flat_read& fr = src->reads[r];
    if (!fr.seq || !fr.qual || !fr.cigar) return nullptr;
    o->seq.push_back(fr.seq);
    o->qual.push_back(fr.qual);
    o->cigar.push_back(fr.cigar);
  }
  // all strings are in place: take the pointers
  for (int a = 0; a < src->n_alleles; ++a) o->allele_ptrs.push_back(o->alleles[(size_t)a].c_str());
  o->reads.resize(nr);
  for (size_t r = 0; r < nr; ++r) {
    o->reads[r].start = src->reads[r].start;
    o->reads[r].stop = src->reads[r].stop;
    o->reads[r].seq = o->seq[r].c_str();
    o->reads[r].qual = o->qual[r].c_str();
    o->reads[r].cigar = o->cigar[r].c_str();
  }
  if (src->realign_to_hap) o->realign_hap.assign(src->realign_to_hap, src->realign_to_hap + src->n_alleles);
  if (src->realign_read) o->realign_read.assign(src->realign_read, src->realign_read + src->n_reads);
  o->L.lflank = o->lflank.c_str();
  o->L.rflank = o->rflank.c_str();
  o->L.motif = o->motif.c_str();
  o->L.alleles = o->allele_ptrs.data();
  o->L.reads = o->reads.data();
  o->L.realign_to_hap = src->realign_to_hap ? o->realign_hap.data() : nullptr;
  o->L.realign_read = src->realign_read ? o->realign_read.data() : nullptr;
  o->ll.assign(nr * (size_t)src->n_alleles, fill);
  o->seeds.assign(nr, -1);
  if (in_seeds)
    for (size_t r = 0; r < nr; ++r) o->seeds[r] = in_seeds[r];
  return o.release();
}

typedef std::vector<std::unique_ptr<OwnedLocus> > Batch;

}  // namespace

struct ltr_pipeline {
  int device = 0;
  size_t batch_loci = 1, max_ready = 2;
  std::mutex mu;
  std::condition_variable cv_work, cv_done, cv_space;
  Batch filling;
  std::deque<Batch> ready;
  std::deque<std::unique_ptr<OwnedLocus> > done;
  std::unique_ptr<OwnedLocus> handed_out;  // storage behind the pointers returned by the last ltr_pipeline_next
  size_t pending = 0;                      // submitted and not yet in `done`
  bool stop = false;
  std::vector<std::thread> workers;
};

namespace {

void worker_main(ltr_pipeline* p) {
  ltr_ctx* ctx = nullptr;
  const int ctx_rc = ltr_ctx_create(p->device, &ctx);
  for (;;) {
    Batch batch;
    {
      std::unique_lock<std::mutex> lk(p->mu);
      p->cv_work.wait(lk, [&] { return p->stop || !p->ready.empty(); });
      if (p->ready.empty()) break;  // stop requested and nothing left to do
      batch.swap(p->ready.front());
      p->ready.pop_front();
      p->cv_space.notify_all();
    }
    int rc = ctx_rc;
    if (rc == LTR_OK) {
      std::vector<ltr_flat_locus> loci(batch.size());
      std::vector<double*> lls(batch.size());
      std::vector<int32_t*> seeds(batch.size());
      for (size_t i = 0; i < batch.size(); ++i) {
        loci[i] = batch[i]->L;
        lls[i] = batch[i]->ll.data();
        seeds[i] = batch[i]->seeds.data();
      }
      rc = ltr_process_reads_flat_batch(ctx, (int32_t)batch.size(), loci.data(), lls.data(), seeds.data());
      if ((rc == LTR_ERR_INVALID || rc == LTR_ERR_UNSUPPORTED) && batch.size() > 1) {
        // one malformed locus must not fail its neighbours: fall back to one call per locus for this batch (device
        // errors -- LTR_ERR_CUDA, LTR_ERR_OOM -- go to every locus of the batch as they are: retrying would repeat them)
        for (size_t i = 0; i < batch.size(); ++i)
          batch[i]->status = ltr_process_reads_flat(ctx, &loci[i], lls[i], seeds[i]);
        rc = LTR_OK;
      } else {
        for (size_t i = 0; i < batch.size(); ++i) batch[i]->status = rc;
      }
    } else {
      for (size_t i = 0; i < batch.size(); ++i) batch[i]->status = rc;
    }
    {
      std::lock_guard<std::mutex> lk(p->mu);
      for (size_t i = 0; i < batch.size(); ++i) p->done.push_back(std::move(batch[i]));
      p->pending -= batch.size();
    }
    p->cv_done.notify_all();
  }
  if (ctx) ltr_ctx_destroy(ctx);
}

void queue_filling_locked(ltr_pipeline* p) {
  if (p->filling.empty()) return;
  p->ready.emplace_back();
  p->ready.back().swap(p->filling);
  p->cv_work.notify_one();
}

}  // namespace

extern "C" {

int ltr_pipeline_create(int device, int32_t batch_loci, int32_t slots, ltr_pipeline** out) {
  if (!out || batch_loci < 1 || slots < 1 || slots > 16) return LTR_ERR_INVALID;
  *out = nullptr;
  ltr_ctx* probe = nullptr;  // fail here, not in a worker, when there is no usable device
  const int rc = ltr_ctx_create(device, &probe);
  if (rc != LTR_OK) return rc;
  ltr_ctx_destroy(probe);
  ltr_pipeline* p = new ltr_pipeline();
  p->device = device;
  p->batch_loci = (size_t)batch_loci;
  p->max_ready = 2 * (size_t)slots;
  for (int s = 0; s < slots; ++s) p->workers.emplace_back(worker_main, p);
  *out = p;
  return LTR_OK;
}

int ltr_pipeline_submit(ltr_pipeline* p, const ltr_flat_locus* locus, uint64_t tag, const int32_t* in_seeds, double fill) {
  if (!p || !locus) return LTR_ERR_INVALID;
  OwnedLocus* o = copy_locus(locus, tag, in_seeds, fill);
  if (!o) return LTR_ERR_INVALID;
  std::unique_lock<std::mutex> lk(p->mu);
  p->cv_space.wait(lk, [&] { return p->ready.size() < p->max_ready; });  // back-pressure on the producer
  p->filling.push_back(std::unique_ptr<OwnedLocus>(o));
  p->pending += 1;
  if (p->filling.size() >= p->batch_loci) queue_filling_locked(p);
  return LTR_OK;
}

int ltr_pipeline_flush(ltr_pipeline* p) {
  if (!p) return LTR_ERR_INVALID;
  std::lock_guard<std::mutex> lk(p->mu);
  queue_filling_locked(p);
  return LTR_OK;
}

int ltr_pipeline_next(ltr_pipeline* p, int wait, uint64_t* tag, int32_t* n_reads, int32_t* n_alleles, const double** ll,
                      const int32_t** seeds, int* status) {
  if (!p) return LTR_ERR_INVALID;
  std::unique_lock<std::mutex> lk(p->mu);
  if (wait) {
    // a partial batch is only pushed out when nothing else can produce a result (every pending locus sits in it):
    // flushing while workers still hold batches would cut the GPU jobs short
    if (p->done.empty() && !p->filling.empty() && p->pending == p->filling.size()) queue_filling_locked(p);
    p->cv_done.wait(lk, [&] { return !p->done.empty() || p->pending == 0; });
  }
  if (p->done.empty()) return 0;
  p->handed_out = std::move(p->done.front());
  p->done.pop_front();
  const OwnedLocus& o = *p->handed_out;
  if (tag) *tag = o.tag;
  if (n_reads) *n_reads = o.L.n_reads;
  if (n_alleles) *n_alleles = o.L.n_alleles;
  if (ll) *ll = o.ll.data();
  if (seeds) *seeds = o.seeds.data();
  if (status) *status = o.status;
  return 1;
}

void ltr_pipeline_destroy(ltr_pipeline* p) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> lk(p->mu);
    queue_filling_locked(p);  // pending loci are still processed; their results are dropped with the pipeline
    p->stop = true;
  }
  p->cv_work.notify_all();
  for (std::thread& t : p->workers) t.join();
  delete p;
}

}  // extern "C"
