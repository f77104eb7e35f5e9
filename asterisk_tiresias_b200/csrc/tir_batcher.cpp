// tir_batcher.cpp -- concurrent front-end of the query path.
//
// In the reference every dialplan Tiresias() call runs fp_search_fingerprint_info() on its own PBX
// thread (src/application_handler.c:180); the threads serialise on SQLite's mutex (one connection,
// src/fp_handler.c:45).  Here the callers block in tir_search_one(); a dispatcher thread owned by
// the context gathers the waiting requests that share (coefs, tolerance, freq_ignore_*) into ONE
// tir_search batch -- one H2D copy, one extraction launch, one match pass -- and hands every caller
// its own tir_hit.  A request never waits longer than max_wait_us for company.
//
// (The streaming entry points tir_stream_* live in tir_stream.cu.)
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <new>
#include <thread>

#include <climits>
#include <linux/futex.h>
#include <sys/syscall.h>
#include <unistd.h>

#include "tir_internal.h"

namespace {

struct Request {
  uint64_t n = 0, off = 0;
  tir_hit *hit = nullptr;
  int rc = TIR_OK;
};

// Completion of a batch is announced through a per-slot generation word that the members sleep on with futex(2):
// a thousand callers woken by one condition variable would re-acquire the dispatcher's mutex one after the other
// (and contend with the callers joining the next batch) -- the tail of the latency distribution.  The members'
// results are written before the release-increment, read after the acquire-load: no lock on the way out.
static void gen_wait(std::atomic<uint32_t> &g, uint32_t seen) {
  while (g.load(std::memory_order_acquire) == seen)
    syscall(SYS_futex, reinterpret_cast<uint32_t *>(&g), FUTEX_WAIT_PRIVATE, seen, nullptr, nullptr, 0);
}
static void gen_bump(std::atomic<uint32_t> &g) {
  g.fetch_add(1, std::memory_order_release);
  syscall(SYS_futex, reinterpret_cast<uint32_t *>(&g), FUTEX_WAKE_PRIVATE, INT_MAX, nullptr, nullptr, 0);
}

struct Params {
  int coefs = 1, ign_lo = -1, ign_hi = -1;
  double tol = 0.001;
  bool operator==(const Params &o) const {
    return coefs == o.coefs && ign_lo == o.ign_lo && ign_hi == o.ign_hi && std::memcmp(&tol, &o.tol, sizeof tol) == 0;
  }
};

// One batch being assembled / in flight.  Callers copy their own recording into the pinned staging
// buffer (in parallel, outside the lock); the dispatcher only launches.
struct Slot {
  int16_t *h_pcm = nullptr; // pinned
  uint64_t cap = 0, used = 0;
  std::vector<Request *> members;
  Params params;
  uint32_t copying = 0; // members that have not finished their memcpy yet
  bool open = false;    // accepts members
  bool sealed = false;  // somebody could not join: dispatch without waiting for max_wait
  std::chrono::steady_clock::time_point t_first;
  std::atomic<uint32_t> done_gen{0}; // + 1 per finished batch of this slot (its members sleep on it)
};

} // namespace

struct TirBatcher {
  // what one sealed batch runs: tir_search of a context, or tir_group_search of a group of devices
  std::function<int(const int16_t *, const uint64_t *, uint32_t, int, double, int, int, tir_hit *)> search;
  std::function<std::string()> last_error;
  int device = 0; // where the pinned staging buffers are registered
  uint32_t max_batch = 1024, max_wait_us = 200;
  std::mutex mu;
  std::condition_variable cv_work, cv_space, cv_done;
  Slot slot[2];
  int fill = 0;
  bool stop = false;
  std::thread worker;
  uint64_t n_requests = 0, n_batches = 0, max_seen = 0;
  std::atomic<int> callers{0}; // threads inside tir_search_one / tir_batcher_stats
  std::vector<uint64_t> off;
  std::vector<tir_hit> hits;
  std::string last_err;

  void run();
};

void TirBatcher::run() {
  std::unique_lock<std::mutex> lk(mu);
  for (;;) {
    Slot &s = slot[fill];
    cv_work.wait(lk, [&] { return stop || !s.members.empty(); });
    if (s.members.empty()) {
      if (stop) return;
      continue;
    }
    // company for the first request: until the batch is full / sealed or max_wait_us have passed
    const auto deadline = s.t_first + std::chrono::microseconds(max_wait_us);
    cv_work.wait_until(lk, deadline, [&] { return stop || s.sealed || s.members.size() >= max_batch; });
    s.open = false;
    const int cur = fill;
    fill ^= 1; // the other slot is idle: this thread is the only one that puts slots in flight
    slot[fill].open = true, slot[fill].sealed = false, slot[fill].used = 0;
    cv_space.notify_all();
    cv_work.wait(lk, [&] { return s.copying == 0; });
    lk.unlock();
    // ---- one batched search (takes the context lock inside tir_search)
    const size_t nb = s.members.size();
    off.assign(nb + 1, 0);
    for (size_t i = 0; i < nb; i++) off[i] = s.members[i]->off;
    off[nb] = s.used; // members were appended in offset order
    hits.resize(nb);
    const int rc = search(s.h_pcm, off.data(), (uint32_t)nb, s.params.coefs, s.params.tol, s.params.ign_lo, s.params.ign_hi, hits.data());
    lk.lock();
    if (rc != TIR_OK) last_err = last_error();
    n_batches++, n_requests += nb;
    if (nb > max_seen) max_seen = nb;
    for (size_t i = 0; i < nb; i++) {
      if (rc == TIR_OK) *s.members[i]->hit = hits[i];
      s.members[i]->rc = rc;
    }
    s.members.clear(), s.used = 0, s.sealed = false; // (the Requests live on their callers' stacks: not touched after the bump)
    (void)cur;
    gen_bump(s.done_gen);
  }
}

void tir_batcher_destroy(TirBatcher *b) {
  if (!b) return;
  {
    std::lock_guard<std::mutex> lk(b->mu);
    b->stop = true;
  }
  b->cv_work.notify_all();
  b->cv_space.notify_all();
  if (b->worker.joinable()) b->worker.join();
  while (b->callers.load() != 0) std::this_thread::yield(); // callers on their way out
  for (Slot &s : b->slot)
    if (s.h_pcm) cudaFreeHost(s.h_pcm);
  delete b;
}

extern "C" {

} // extern "C"

// a dispatcher over any batched search function (tir_internal.h)
TirBatcher *tir_batcher_create(int device, uint32_t max_batch, uint32_t max_wait_us,
                               std::function<int(const int16_t *, const uint64_t *, uint32_t, int, double, int, int, tir_hit *)> search,
                               std::function<std::string()> last_error) {
  TirBatcher *b = new (std::nothrow) TirBatcher();
  if (!b) return nullptr;
  b->search = std::move(search), b->last_error = std::move(last_error), b->device = device;
  b->max_batch = max_batch, b->max_wait_us = max_wait_us;
  // staging: room for max_batch recordings of 8 s at 8 kHz (4 s at 16 kHz) each, at least 4 M samples; a recording
  // that does not fit an empty buffer is searched directly by its caller
  cudaSetDevice(device);
  const uint64_t cap = std::max<uint64_t>((uint64_t)max_batch * 65536ull, 4ull << 20);
  for (Slot &s : b->slot) {
    if (cudaMallocHost((void **)&s.h_pcm, cap * sizeof(int16_t)) != cudaSuccess) {
      tir_batcher_destroy(b);
      return nullptr;
    }
    s.cap = cap;
    s.members.reserve(max_batch);
  }
  b->slot[0].open = true;
  b->worker = std::thread([b] {
    cudaSetDevice(b->device);
    b->run();
  });
  return b;
}

// one caller's recording through the dispatcher; rc of the batch it rode in (message in *err)
int tir_batcher_submit(TirBatcher *b, const int16_t *pcm, uint64_t n_samples, int coefs, double tolerance, int freq_ignore_low,
                       int freq_ignore_high, tir_hit *hit, std::string *err) {
  Params p;
  p.coefs = coefs, p.tol = tolerance, p.ign_lo = freq_ignore_low, p.ign_hi = freq_ignore_high;
  Request r;
  r.n = n_samples, r.hit = hit;
  std::unique_lock<std::mutex> lk(b->mu);
  Slot *s = nullptr;
  for (;;) {
    if (b->stop) {
      if (err) *err = "batcher is stopping";
      return TIR_ERR_STATE;
    }
    s = &b->slot[b->fill];
    if (s->open && (s->members.empty() || s->params == p) && s->members.size() < b->max_batch && s->used + n_samples <= s->cap) break;
    if (s->open && !s->members.empty()) { // cannot join this batch: send it off now, take the next one
      s->sealed = true;
      b->cv_work.notify_all();
    }
    b->cv_space.wait(lk);
  }
  if (s->members.empty()) s->params = p, s->t_first = std::chrono::steady_clock::now();
  r.off = s->used, s->used += n_samples;
  s->members.push_back(&r);
  s->copying++;
  const uint32_t gen0 = s->done_gen.load(std::memory_order_relaxed); // this slot's batches finished so far
  if (s->members.size() == 1 || s->members.size() >= b->max_batch) b->cv_work.notify_all();
  int16_t *dst = s->h_pcm + r.off;
  lk.unlock();
  if (n_samples) std::memcpy(dst, pcm, n_samples * sizeof(int16_t)); // in parallel with the other callers
  lk.lock();
  if (--s->copying == 0) b->cv_work.notify_all();
  lk.unlock();
  gen_wait(s->done_gen, gen0); // the batch this request rode in has been served
  if (r.rc != TIR_OK && err) {
    lk.lock();
    *err = b->last_err;
  }
  return r.rc;
}

bool tir_batcher_fits(TirBatcher *b, uint64_t n_samples) { return n_samples <= b->slot[0].cap; }
void tir_batcher_enter(TirBatcher *b) { b->callers++; }
void tir_batcher_leave(TirBatcher *b) { b->callers--; }
void tir_batcher_counters(TirBatcher *b, uint64_t *n_requests, uint64_t *n_batches, uint64_t *max_batch_seen) {
  std::lock_guard<std::mutex> lk(b->mu);
  *n_requests = b->n_requests, *n_batches = b->n_batches, *max_batch_seen = b->max_seen;
}

extern "C" {

int tir_batcher_start(tir_ctx *ctx, uint32_t max_batch, uint32_t max_wait_us) {
  if (!ctx || max_batch == 0) return tir_fail(ctx, TIR_ERR_ARG, "bad batcher arguments");
  std::lock_guard<std::mutex> lk(ctx->batcher_mu);
  if (ctx->batcher) return tir_fail(ctx, TIR_ERR_STATE, "batcher already running");
  TirBatcher *b = tir_batcher_create(
      ctx->cfg.device, max_batch, max_wait_us,
      [ctx](const int16_t *pcm, const uint64_t *off, uint32_t n, int coefs, double tol, int lo, int hi, tir_hit *hits) {
        return tir_search(ctx, pcm, off, n, coefs, tol, lo, hi, hits);
      },
      [ctx] { return std::string(tir_last_error(ctx)); });
  if (!b) return tir_fail(ctx, TIR_ERR_NOMEM, "could not start the batcher (pinned staging buffers)");
  ctx->batcher = b;
  return TIR_OK;
}

int tir_batcher_stop(tir_ctx *ctx) {
  if (!ctx) return TIR_ERR_ARG;
  TirBatcher *b;
  {
    std::lock_guard<std::mutex> lk(ctx->batcher_mu);
    b = ctx->batcher, ctx->batcher = nullptr; // no new caller can reach it
  }
  tir_batcher_destroy(b); // requests already admitted to a batch are served first
  return TIR_OK;
}

int tir_search_one(tir_ctx *ctx, const int16_t *pcm, uint64_t n_samples, int coefs, double tolerance,
                   int freq_ignore_low, int freq_ignore_high, tir_hit *hit) {
  if (!ctx || !hit || (!pcm && n_samples)) return tir_fail(ctx, TIR_ERR_ARG, "null argument");
  // argument checks come first in the reference too (src/fp_handler.c:247)
  if (coefs < 1 || coefs > TIR_N_COEFS) return tir_fail(ctx, TIR_ERR_ARG, "Wrong coefs count. max[%d], coefs[%d]", TIR_N_COEFS, coefs);
  TirBatcher *b;
  {
    std::lock_guard<std::mutex> g(ctx->batcher_mu);
    b = ctx->batcher;
    if (b) b->callers++;
  }
  struct Leave {
    TirBatcher *b;
    ~Leave() { if (b) b->callers--; }
  } leave{b};
  const uint64_t direct_off[2] = {0, n_samples};
  if (!b || n_samples > b->slot[0].cap) // no dispatcher (or an oversized recording): a batch of one
    return tir_search(ctx, pcm, direct_off, 1, coefs, tolerance, freq_ignore_low, freq_ignore_high, hit);
  std::string err;
  const int rc = tir_batcher_submit(b, pcm, n_samples, coefs, tolerance, freq_ignore_low, freq_ignore_high, hit, &err);
  if (rc != TIR_OK) return tir_fail(ctx, rc, "%s", err.c_str());
  return TIR_OK;
}

int tir_batcher_stats(tir_ctx *ctx, uint64_t *n_requests, uint64_t *n_batches, uint64_t *max_batch_seen) {
  if (!ctx) return TIR_ERR_ARG;
  uint64_t r = 0, n = 0, m = 0;
  {
    std::lock_guard<std::mutex> g(ctx->batcher_mu);
    if (TirBatcher *b = ctx->batcher) {
      std::lock_guard<std::mutex> lk(b->mu);
      r = b->n_requests, n = b->n_batches, m = b->max_seen;
    }
  }
  if (n_requests) *n_requests = r;
  if (n_batches) *n_batches = n;
  if (max_batch_seen) *max_batch_seen = m;
  return TIR_OK;
}

} // extern "C"
