// tir_batcher.cpp -- concurrent front-end of the query path.
//
// In the reference every dialplan Tiresias() call runs fp_search_fingerprint_info() on its own PBX
// thread (src/application_handler.c:180); the threads serialise on SQLite's mutex (one connection,
// src/fp_handler.c:45).  Here the callers block in tir_search_one(); a dispatcher thread owned by
// the context gathers the waiting requests that share (coefs, tolerance, freq_ignore_*) into ONE
// tir_search batch -- one H2D copy, one extraction launch, one match pass -- and hands every caller
// its own tir_hit.  A request never waits longer than max_wait_us for company.
//
// tir_stream_* replaces the WAV-file round trip of the application (record_voice writes
// /tmp/tiresias-<uuid>.wav, src/application_handler.c:153-155,248-312, which
// create_audio_fingerprints reopens): the channel's frames are appended to a host buffer as
// ast_read delivers them and tir_stream_finish submits the recording to the batcher.
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <memory>
#include <new>
#include <thread>

#include "tir_internal.h"

namespace {

struct Request {
  const int16_t *pcm;
  uint64_t n;
  int coefs, ign_lo, ign_hi;
  double tol;
  tir_hit *hit;
  int rc = TIR_OK;
  bool done = false;
  std::string err;
  bool same_params(const Request &o) const {
    return coefs == o.coefs && ign_lo == o.ign_lo && ign_hi == o.ign_hi && std::memcmp(&tol, &o.tol, sizeof tol) == 0;
  }
};

} // namespace

struct TirBatcher {
  tir_ctx *ctx = nullptr;
  uint32_t max_batch = 1024, max_wait_us = 200;
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  std::deque<Request *> queue;
  bool stop = false;
  std::thread worker;
  uint64_t n_requests = 0, n_batches = 0, max_seen = 0;
  // staging (pinned, grows only)
  int16_t *h_pcm = nullptr;
  size_t h_cap = 0;
  std::vector<uint64_t> off;
  std::vector<tir_hit> hits;

  void run();
};

void TirBatcher::run() {
  std::unique_lock<std::mutex> lk(mu);
  std::vector<Request *> batch;
  for (;;) {
    cv_work.wait(lk, [&] { return stop || !queue.empty(); });
    if (queue.empty()) {
      if (stop) return;
      continue;
    }
    // company for the first request: until max_batch are waiting or max_wait_us have passed
    if (!stop && queue.size() < max_batch && max_wait_us)
      cv_work.wait_for(lk, std::chrono::microseconds(max_wait_us), [&] { return stop || queue.size() >= max_batch; });
    batch.clear();
    Request *head = queue.front();
    for (auto it = queue.begin(); it != queue.end() && batch.size() < max_batch;) {
      if ((*it)->same_params(*head)) {
        batch.push_back(*it);
        it = queue.erase(it);
      } else {
        ++it;
      }
    }
    lk.unlock();
    // ---- one batched search (takes the context lock inside tir_search)
    uint64_t total = 0;
    off.assign(batch.size() + 1, 0);
    for (size_t i = 0; i < batch.size(); i++) total += batch[i]->n, off[i + 1] = total;
    int rc = TIR_OK;
    std::string err;
    if (total > h_cap) {
      if (h_pcm) cudaFreeHost(h_pcm);
      h_pcm = nullptr;
      h_cap = total + total / 4 + 4096;
      if (cudaMallocHost((void **)&h_pcm, h_cap * sizeof(int16_t)) != cudaSuccess) {
        h_pcm = nullptr, h_cap = 0;
        rc = TIR_ERR_NOMEM, err = "cudaMallocHost failed for the batch staging buffer";
      }
    }
    if (rc == TIR_OK) {
      for (size_t i = 0; i < batch.size(); i++)
        if (batch[i]->n) std::memcpy(h_pcm + off[i], batch[i]->pcm, batch[i]->n * sizeof(int16_t));
      hits.resize(batch.size());
      rc = tir_search(ctx, h_pcm, off.data(), (uint32_t)batch.size(), head->coefs, head->tol, head->ign_lo, head->ign_hi,
                      hits.data());
      if (rc != TIR_OK) err = tir_last_error(ctx);
    }
    lk.lock();
    n_batches++, n_requests += batch.size();
    if (batch.size() > max_seen) max_seen = batch.size();
    for (size_t i = 0; i < batch.size(); i++) {
      if (rc == TIR_OK) *batch[i]->hit = hits[i];
      batch[i]->rc = rc, batch[i]->err = err, batch[i]->done = true;
    }
    cv_done.notify_all();
  }
}

void tir_batcher_destroy(TirBatcher *b) {
  if (!b) return;
  {
    std::lock_guard<std::mutex> lk(b->mu);
    b->stop = true;
  }
  b->cv_work.notify_all();
  if (b->worker.joinable()) b->worker.join();
  if (b->h_pcm) cudaFreeHost(b->h_pcm);
  delete b;
}

struct tir_stream {
  tir_ctx *ctx;
  std::vector<int16_t> pcm;
};

extern "C" {

int tir_batcher_start(tir_ctx *ctx, uint32_t max_batch, uint32_t max_wait_us) {
  if (!ctx || max_batch == 0) return tir_fail(ctx, TIR_ERR_ARG, "bad batcher arguments");
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (ctx->batcher) return tir_fail(ctx, TIR_ERR_STATE, "batcher already running");
  TirBatcher *b = new (std::nothrow) TirBatcher();
  if (!b) return tir_fail(ctx, TIR_ERR_NOMEM, "out of memory");
  b->ctx = ctx, b->max_batch = max_batch, b->max_wait_us = max_wait_us;
  b->worker = std::thread([b] {
    cudaSetDevice(b->ctx->cfg.device);
    b->run();
  });
  ctx->batcher = b;
  return TIR_OK;
}

int tir_batcher_stop(tir_ctx *ctx) {
  if (!ctx) return TIR_ERR_ARG;
  TirBatcher *b;
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    b = ctx->batcher, ctx->batcher = nullptr;
  }
  tir_batcher_destroy(b); // pending requests are served first (the worker drains the queue)
  return TIR_OK;
}

int tir_search_one(tir_ctx *ctx, const int16_t *pcm, uint64_t n_samples, int coefs, double tolerance,
                   int freq_ignore_low, int freq_ignore_high, tir_hit *hit) {
  if (!ctx || !hit || (!pcm && n_samples)) return tir_fail(ctx, TIR_ERR_ARG, "null argument");
  // argument checks come first in the reference too (src/fp_handler.c:247)
  if (coefs < 1 || coefs > TIR_N_COEFS) return tir_fail(ctx, TIR_ERR_ARG, "Wrong coefs count. max[%d], coefs[%d]", TIR_N_COEFS, coefs);
  TirBatcher *b = ctx->batcher;
  if (!b) { // no dispatcher: a batch of one
    const uint64_t off[2] = {0, n_samples};
    return tir_search(ctx, pcm, off, 1, coefs, tolerance, freq_ignore_low, freq_ignore_high, hit);
  }
  Request r;
  r.pcm = pcm, r.n = n_samples, r.coefs = coefs, r.tol = tolerance, r.ign_lo = freq_ignore_low, r.ign_hi = freq_ignore_high;
  r.hit = hit;
  std::unique_lock<std::mutex> lk(b->mu);
  if (b->stop) return tir_fail(ctx, TIR_ERR_STATE, "batcher is stopping");
  b->queue.push_back(&r);
  b->cv_work.notify_one();
  b->cv_done.wait(lk, [&] { return r.done; });
  return r.rc;
}

int tir_batcher_stats(tir_ctx *ctx, uint64_t *n_requests, uint64_t *n_batches, uint64_t *max_batch_seen) {
  if (!ctx) return TIR_ERR_ARG;
  TirBatcher *b = ctx->batcher;
  uint64_t r = 0, n = 0, m = 0;
  if (b) {
    std::lock_guard<std::mutex> lk(b->mu);
    r = b->n_requests, n = b->n_batches, m = b->max_seen;
  }
  if (n_requests) *n_requests = r;
  if (n_batches) *n_batches = n;
  if (max_batch_seen) *max_batch_seen = m;
  return TIR_OK;
}

int tir_stream_open(tir_ctx *ctx, tir_stream **out) {
  if (!ctx || !out) return TIR_ERR_ARG;
  tir_stream *s = new (std::nothrow) tir_stream();
  if (!s) return tir_fail(ctx, TIR_ERR_NOMEM, "out of memory");
  s->ctx = ctx;
  *out = s;
  return TIR_OK;
}

int tir_stream_feed(tir_stream *s, const int16_t *pcm, uint32_t n_samples) {
  if (!s || (!pcm && n_samples)) return TIR_ERR_ARG;
  try {
    s->pcm.insert(s->pcm.end(), pcm, pcm + n_samples);
  } catch (const std::bad_alloc &) {
    return tir_fail(s->ctx, TIR_ERR_NOMEM, "out of memory");
  }
  return TIR_OK;
}

uint64_t tir_stream_samples(const tir_stream *s) { return s ? s->pcm.size() : 0; }

int tir_stream_finish(tir_stream *s, int coefs, double tolerance, int freq_ignore_low, int freq_ignore_high, tir_hit *hit) {
  if (!s) return TIR_ERR_ARG;
  return tir_search_one(s->ctx, s->pcm.data(), s->pcm.size(), coefs, tolerance, freq_ignore_low, freq_ignore_high, hit);
}

void tir_stream_close(tir_stream *s) { delete s; }

} // extern "C"
