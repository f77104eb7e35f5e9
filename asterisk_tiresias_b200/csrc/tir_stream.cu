// tir_stream.cu -- streaming front-end of the dialplan application.
//
// The reference records the caller to /tmp/tiresias-<uuid>.wav for `duration` ms (record_voice,
// src/application_handler.c:248-312: one ast_read() of a 20 ms slinear frame after the other), closes the
// file, and only then runs the whole hop loop over it (fp_search_fingerprint_info ->
// create_audio_fingerprints, src/fp_handler.c:577-671) followed by the match.  Here the hop loop runs WHILE
// the call is being recorded: tir_stream_feed() appends the channel's frames to the stream's pending buffer;
// a pump thread owned by the context gathers, every few milliseconds, the COMPLETED hops of ALL live streams
// into one batched launch of the extraction kernel and scatters the new coefficients to each stream's
// device-resident coefficient array.  A frame depends only on its hop and the one before it (aubio_pvoc keeps
// win - hop samples of history), so a stream's state between batches is one hop of PCM:
//      segment = [last hop already processed (history) | new complete hops]  -> extracted as a pseudo-clip,
//      its first frame ([zeros | history]) is discarded, the others are the stream's next frames, bit for bit
//      the frames a one-shot extraction of the whole recording produces.
// tir_stream_finish() hands the final partial hop (zero padded, as aubio_source_do pads the last block) to the
// pump and waits: the pump batches the finishing streams of equal parameters into ONE match over their
// device-resident coefficients.  Time from the last feed to the result: one pump period + the extraction of
// at most one or two hops + one match chain -- independent of the length of the recording.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <new>
#include <thread>

#include <climits>
#include <linux/futex.h>
#include <sys/syscall.h>
#include <unistd.h>

#include "tir_internal.h"

namespace {

struct Params {
  int coefs = 1, ign_lo = -1, ign_hi = -1;
  double tol = 0.001;
  bool operator==(const Params &o) const {
    return coefs == o.coefs && ign_lo == o.ign_lo && ign_hi == o.ign_hi && std::memcmp(&tol, &o.tol, sizeof tol) == 0;
  }
};

// one segment of a batch: frames [skip, n_frames) of pseudo-clip `clip` go to dst[0 ..]
struct SegMeta {
  float *dst;
  uint64_t src_frame; // first frame of the pseudo-clip in the batch's coefficient buffer
  uint32_t skip, n_frames;
};

} // namespace

__global__ void tir_stream_scatter_kernel(const SegMeta *__restrict__ seg, uint32_t n_seg, const float *__restrict__ src) {
  const uint32_t s = blockIdx.x;
  if (s >= n_seg) return;
  const SegMeta m = seg[s];
  const float2 *in = reinterpret_cast<const float2 *>(src) + m.src_frame + m.skip;
  float2 *out = reinterpret_cast<float2 *>(m.dst);
  for (uint32_t f = threadIdx.x; f + m.skip < m.n_frames; f += blockDim.x) out[f] = in[f];
}
// the other way round for the match batch: a stream's frames [0, n_frames) -> dst (contiguous batch buffer)
__global__ void tir_stream_gather_kernel(const SegMeta *__restrict__ seg, uint32_t n_seg, float *__restrict__ dst) {
  const uint32_t s = blockIdx.x;
  if (s >= n_seg) return;
  const SegMeta m = seg[s];
  const float2 *in = reinterpret_cast<const float2 *>(m.dst);
  float2 *out = reinterpret_cast<float2 *>(dst) + m.src_frame;
  for (uint32_t f = threadIdx.x; f < m.n_frames; f += blockDim.x) out[f] = in[f];
}

struct TirStreamHub;

struct tir_stream {
  tir_ctx *ctx = nullptr;
  TirStreamHub *hub = nullptr;
  std::mutex mu;                // pending / counters: feed (caller thread) against the pump
  std::vector<int16_t> pending; // [history hop, once a hop has been taken] + samples not yet handed to the pump
  uint64_t n_samples = 0;       // fed so far
  uint64_t hops_taken = 0;      // complete hops handed to the pump
  bool finishing = false, flushed = false, failed = false;
  // owned by the pump thread
  float *d_coef = nullptr;
  uint64_t cap_frames = 0, frames = 0; // frames whose extraction has been enqueued (stream order makes them visible to the match)
  std::atomic<uint64_t> frames_pub{0};
  // finish rendez-vous (guarded by hub->mu)
  Params params;
  tir_hit hit{};
  int rc = TIR_OK;
  bool closing = false, busy = false;
  // set (release) after rc / hit are in place; tir_stream_finish sleeps on it with futex(2) -- a thousand finishing
  // callers woken through one condition variable would queue up on the hub's mutex
  std::atomic<uint32_t> done{0};
};

static void stream_mark_done(tir_stream *s) {
  s->done.store(1, std::memory_order_release);
  syscall(SYS_futex, reinterpret_cast<uint32_t *>(&s->done), FUTEX_WAKE_PRIVATE, INT_MAX, nullptr, nullptr, 0);
}
static void stream_wait_done(tir_stream *s) {
  while (s->done.load(std::memory_order_acquire) == 0)
    syscall(SYS_futex, reinterpret_cast<uint32_t *>(&s->done), FUTEX_WAIT_PRIVATE, 0, nullptr, nullptr, 0);
}

struct TirStreamHub {
  tir_ctx *ctx = nullptr;
  std::mutex mu; // live list, finish / close rendez-vous
  std::condition_variable cv_pump, cv_done;
  std::vector<tir_stream *> live;
  bool stop = false, kick = false;
  uint32_t period_us = 2000;
  std::thread worker;
  std::string last_err;
  // pump-owned scratch
  int16_t *h_pcm = nullptr; // pinned
  size_t h_cap = 0;
  SegMeta *h_seg = nullptr; // pinned
  size_t seg_cap = 0;
  DevBuf d_pcm, d_tmp, d_seg, d_qcoef, d_qhits;
  tir_hit *h_hits = nullptr;
  size_t hits_cap = 0;
  uint64_t n_batches = 0, n_hops = 0, n_match_batches = 0;
  // coefficient arrays of closed streams, reused by the next ones: a cudaMalloc per new stream and a device-wide
  // synchronising cudaFree per closed one were most of the front-end's latency at a thousand channels.  Everything
  // that touches these arrays is ordered on the context's stream, so a reused array needs no synchronisation.
  std::mutex pool_mu;
  std::vector<std::pair<float *, uint64_t>> pool; // (array, capacity in frames)
  float *pool_take(uint64_t cap_frames, uint64_t *got_cap) {
    std::lock_guard<std::mutex> lk(pool_mu);
    for (size_t i = 0; i < pool.size(); i++)
      if (pool[i].second >= cap_frames) {
        float *p = pool[i].first;
        *got_cap = pool[i].second;
        pool[i] = pool.back(), pool.pop_back();
        return p;
      }
    return nullptr;
  }
  bool pool_give(float *p, uint64_t cap_frames) { // false: the pool is full, the caller frees
    std::lock_guard<std::mutex> lk(pool_mu);
    if (pool.size() >= 8192) return false;
    pool.emplace_back(p, cap_frames);
    return true;
  }

  void run();
  int extract_batch(std::vector<tir_stream *> &streams);
  void match_batch(std::vector<tir_stream *> &group);
};

int TirStreamHub_grow_pcm(TirStreamHub *h, size_t samples);

static int hub_reserve_pinned(void **p, size_t *cap, size_t bytes) {
  if (bytes <= *cap) return TIR_OK;
  if (*p) cudaFreeHost(*p);
  *p = nullptr, *cap = 0;
  const size_t c = bytes + bytes / 2 + 4096;
  if (cudaMallocHost(p, c) != cudaSuccess) return TIR_ERR_NOMEM;
  *cap = c;
  return TIR_OK;
}

// take what the pump may process from a stream: all complete hops (+ the final partial one when finishing)
struct Taken {
  tir_stream *s;
  size_t off, len; // in the batch's PCM buffer
  uint32_t skip, n_frames;
};

int TirStreamHub::extract_batch(std::vector<tir_stream *> &streams) {
  const int hop = ctx->cfg.hop;
  std::vector<Taken> taken;
  size_t total = 0;
  // ---- collect (per-stream lock only)
  for (tir_stream *s : streams) {
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->failed || s->flushed) continue; // (a flushed stream has handed in its last hop: only its match is left)
    const size_t hist = s->hops_taken ? (size_t)hop : 0;
    if (s->pending.size() < hist) continue;
    const size_t avail = s->pending.size() - hist;
    size_t nh = avail / (size_t)hop, rem = avail % (size_t)hop;
    bool last = false;
    if (s->finishing && !s->flushed) {
      s->flushed = true; // the final, zero-padded hop of aubio_source_do (src/fp_handler.c:633-636)
      last = rem > 0;
    } else {
      rem = 0;
    }
    if (nh == 0 && !last) continue;
    const size_t len = hist + nh * (size_t)hop + (last ? rem : 0);
    if (TirStreamHub_grow_pcm(this, total + len)) return TIR_ERR_NOMEM;
    std::memcpy(h_pcm + total, s->pending.data(), len * sizeof(int16_t)); // the pseudo-clips lie back to back
    Taken t{s, total, len, (uint32_t)(hist ? 1 : 0), (uint32_t)((len + hop - 1) / hop)};
    taken.push_back(t);
    total += len;
    // keep the last complete hop as the next segment's history, and whatever was not taken
    const size_t consumed = hist + nh * (size_t)hop; // (a final partial hop ends the stream)
    if (nh > 0) {
      s->pending.erase(s->pending.begin(), s->pending.begin() + (consumed - hop));
      s->hops_taken += nh;
    }
    if (last) s->pending.clear();
  }
  if (taken.empty()) return TIR_OK;
  // ---- one H2D, one extraction launch, one scatter
  std::lock_guard<std::mutex> lk(ctx->mu);
  TIR_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
  const uint32_t n = (uint32_t)taken.size();
  std::vector<uint64_t> clip_off((size_t)n + 1);
  uint64_t F = 0;
  if (n > seg_cap) {
    void *p = h_seg;
    size_t cap = seg_cap * sizeof(SegMeta);
    if (hub_reserve_pinned(&p, &cap, (size_t)n * sizeof(SegMeta))) return TIR_ERR_NOMEM;
    h_seg = (SegMeta *)p, seg_cap = cap / sizeof(SegMeta);
  }
  int rc;
  for (uint32_t i = 0; i < n; i++) {
    Taken &t = taken[i];
    tir_stream *s = t.s;
    clip_off[i] = t.off;
    const uint64_t add = t.n_frames - t.skip;
    if (s->frames + add > s->cap_frames) { // grow the stream's coefficient array (4096 frames = 131 s at 8 kHz / hop 256 to start with)
      uint64_t cap = std::max<uint64_t>(4096, s->cap_frames * 2);
      while (cap < s->frames + add) cap *= 2;
      float *np = pool_take(cap, &cap);
      if (!np) TIR_CUDA(ctx, cudaMalloc(&np, cap * TIR_N_COEFS * sizeof(float)));
      if (s->frames) TIR_CUDA(ctx, cudaMemcpyAsync(np, s->d_coef, s->frames * TIR_N_COEFS * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
      if (s->d_coef && !pool_give(s->d_coef, s->cap_frames)) {
        TIR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(s->d_coef);
      }
      s->d_coef = np, s->cap_frames = cap;
    }
    h_seg[i] = SegMeta{s->d_coef + s->frames * TIR_N_COEFS, F, t.skip, t.n_frames};
    F += t.n_frames;
  }
  clip_off[n] = total;
  if ((rc = tir_reserve(ctx, d_pcm, total * sizeof(int16_t) + 16))) return rc;
  if ((rc = tir_reserve(ctx, d_tmp, F * TIR_N_COEFS * sizeof(float) + 16))) return rc;
  if ((rc = tir_reserve(ctx, d_seg, (size_t)n * sizeof(SegMeta)))) return rc;
  TIR_CUDA(ctx, cudaMemcpyAsync(d_pcm.p, h_pcm, total * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
  TIR_CUDA(ctx, cudaMemcpyAsync(d_seg.p, h_seg, (size_t)n * sizeof(SegMeta), cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = tir_extract_launch(ctx, (const int16_t *)d_pcm.p, total, clip_off.data(), n, (float *)d_tmp.p, nullptr, nullptr))) return rc;
  tir_stream_scatter_kernel<<<n, 128, 0, ctx->stream>>>((const SegMeta *)d_seg.p, n, (const float *)d_tmp.p);
  TIR_CUDA(ctx, cudaGetLastError());
  ctx->launches++;
  // the pinned buffers are reused by the next batch: wait for the copies (the kernels behind them are short)
  TIR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (uint32_t i = 0; i < n; i++) {
    taken[i].s->frames += taken[i].n_frames - taken[i].skip;
    taken[i].s->frames_pub.store(taken[i].s->frames, std::memory_order_release);
  }
  n_batches++, n_hops += F;
  return TIR_OK;
}

// finishing streams of equal parameters: ONE match over their device-resident coefficients
void TirStreamHub::match_batch(std::vector<tir_stream *> &group) {
  const uint32_t Q = (uint32_t)group.size();
  std::vector<uint64_t> foff((size_t)Q + 1, 0);
  for (uint32_t i = 0; i < Q; i++) foff[i + 1] = foff[i] + group[i]->frames;
  const uint64_t F = foff[Q];
  int rc = TIR_OK;
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    rc = [&]() -> int {
      TIR_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
      int r;
      if (Q > seg_cap) {
        void *p = h_seg;
        size_t cap = seg_cap * sizeof(SegMeta);
        if (hub_reserve_pinned(&p, &cap, (size_t)Q * sizeof(SegMeta))) return TIR_ERR_NOMEM;
        h_seg = (SegMeta *)p, seg_cap = cap / sizeof(SegMeta);
      }
      if (Q > hits_cap) {
        void *p = h_hits;
        size_t cap = hits_cap * sizeof(tir_hit);
        if (hub_reserve_pinned(&p, &cap, (size_t)Q * sizeof(tir_hit))) return TIR_ERR_NOMEM;
        h_hits = (tir_hit *)p, hits_cap = cap / sizeof(tir_hit);
      }
      for (uint32_t i = 0; i < Q; i++) h_seg[i] = SegMeta{group[i]->d_coef, foff[i], 0, (uint32_t)group[i]->frames};
      if ((r = tir_reserve(ctx, d_seg, (size_t)Q * sizeof(SegMeta)))) return r;
      if ((r = tir_reserve(ctx, d_qcoef, std::max<uint64_t>(F, 1) * TIR_N_COEFS * sizeof(float)))) return r;
      if ((r = tir_reserve(ctx, d_qhits, (size_t)Q * sizeof(tir_hit)))) return r;
      TIR_CUDA(ctx, cudaMemcpyAsync(d_seg.p, h_seg, (size_t)Q * sizeof(SegMeta), cudaMemcpyHostToDevice, ctx->stream));
      tir_stream_gather_kernel<<<Q, 128, 0, ctx->stream>>>((const SegMeta *)d_seg.p, Q, (float *)d_qcoef.p);
      TIR_CUDA(ctx, cudaGetLastError());
      ctx->launches++;
      return TIR_OK;
    }();
  }
  const Params &p = group[0]->params;
  if (rc == TIR_OK) rc = tir_match_dev(ctx, (const float *)d_qcoef.p, foff.data(), Q, p.coefs, p.tol, p.ign_lo, p.ign_hi, (tir_hit *)d_qhits.p);
  if (rc == TIR_OK) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaSetDevice(ctx->cfg.device);
    if (cudaMemcpyAsync(h_hits, d_qhits.p, (size_t)Q * sizeof(tir_hit), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess)
      rc = tir_fail(ctx, TIR_ERR_CUDA, "tir_stream: reading the hits failed");
  }
  if (rc != TIR_OK) last_err = tir_last_error(ctx);
  n_match_batches++;
  std::lock_guard<std::mutex> lk(mu);
  for (uint32_t i = 0; i < Q; i++) {
    group[i]->rc = rc;
    if (rc == TIR_OK) group[i]->hit = h_hits[i];
    stream_mark_done(group[i]);
  }
}

void TirStreamHub::run() {
  cudaSetDevice(ctx->cfg.device);
  std::unique_lock<std::mutex> lk(mu);
  for (;;) {
    cv_pump.wait_for(lk, std::chrono::microseconds(period_us), [&] { return stop || kick; });
    if (stop) return;
    kick = false;
    std::vector<tir_stream *> cur;
    for (tir_stream *s : live)
      if (!s->closing) s->busy = true, cur.push_back(s);
    lk.unlock();
    const int rc = extract_batch(cur);
    if (rc != TIR_OK) { // the streams of this batch cannot be completed consistently: fail them
      last_err = tir_last_error(ctx);
      for (tir_stream *s : cur) {
        std::lock_guard<std::mutex> l2(s->mu);
        s->failed = true;
      }
    }
    // finishing streams whose last hop is in: group by parameters, one match per group
    std::vector<tir_stream *> fin;
    for (tir_stream *s : cur) {
      std::lock_guard<std::mutex> l2(s->mu);
      if (s->finishing && (s->flushed || s->failed)) fin.push_back(s);
    }
    lk.lock();
    fin.erase(std::remove_if(fin.begin(), fin.end(), [](tir_stream *s) { return s->done.load(std::memory_order_acquire) != 0; }), fin.end());
    lk.unlock();
    while (!fin.empty()) {
      std::vector<tir_stream *> group, rest;
      for (tir_stream *s : fin) {
        bool failed;
        {
          std::lock_guard<std::mutex> l2(s->mu);
          failed = s->failed;
        }
        if (failed) {
          std::lock_guard<std::mutex> l3(mu);
          s->rc = TIR_ERR_CUDA;
          stream_mark_done(s);
        } else if (group.empty() || s->params == group[0]->params) {
          group.push_back(s);
        } else {
          rest.push_back(s);
        }
      }
      if (!group.empty()) match_batch(group);
      fin.swap(rest);
    }
    lk.lock();
    for (tir_stream *s : cur) s->busy = false;
    cv_done.notify_all();
  }
}

// (helper used above: grows the hub's pinned PCM buffer keeping its contents)
int TirStreamHub_grow_pcm(TirStreamHub *h, size_t samples) {
  if (samples <= h->h_cap) return TIR_OK;
  const size_t cap = samples + samples / 2 + 65536;
  int16_t *np = nullptr;
  if (cudaMallocHost((void **)&np, cap * sizeof(int16_t)) != cudaSuccess) return TIR_ERR_NOMEM;
  if (h->h_pcm) {
    std::memcpy(np, h->h_pcm, h->h_cap * sizeof(int16_t));
    cudaFreeHost(h->h_pcm);
  }
  h->h_pcm = np, h->h_cap = cap;
  return TIR_OK;
}

void tir_stream_hub_destroy(TirStreamHub *h) {
  if (!h) return;
  {
    std::lock_guard<std::mutex> lk(h->mu);
    h->stop = true;
    for (tir_stream *s : h->live) // callers blocked in finish: released with an error
      if (s->finishing && !s->done.load()) {
        s->rc = TIR_ERR_STATE;
        stream_mark_done(s);
      }
  }
  h->cv_pump.notify_all();
  h->cv_done.notify_all();
  if (h->worker.joinable()) h->worker.join();
  cudaSetDevice(h->ctx->cfg.device);
  if (h->h_pcm) cudaFreeHost(h->h_pcm);
  if (h->h_seg) cudaFreeHost(h->h_seg);
  if (h->h_hits) cudaFreeHost(h->h_hits);
  for (DevBuf *b : {&h->d_pcm, &h->d_tmp, &h->d_seg, &h->d_qcoef, &h->d_qhits})
    if (b->p) cudaFree(b->p);
  for (auto &e : h->pool) cudaFree(e.first);
  for (tir_stream *s : h->live) s->hub = nullptr; // the owners still close them
  delete h;
}

static TirStreamHub *hub_of(tir_ctx *ctx) {
  std::lock_guard<std::mutex> lk(ctx->batcher_mu);
  if (!ctx->stream_hub) {
    TirStreamHub *h = new (std::nothrow) TirStreamHub();
    if (!h) return nullptr;
    h->ctx = ctx;
    if (const char *e = getenv("TIR_STREAM_PERIOD_US")) h->period_us = (uint32_t)std::max(50, atoi(e));
    h->worker = std::thread([h] { h->run(); });
    ctx->stream_hub = h;
  }
  return ctx->stream_hub;
}

extern "C" {

int tir_stream_open(tir_ctx *ctx, tir_stream **out) {
  if (!ctx || !out) return TIR_ERR_ARG;
  *out = nullptr;
  TirStreamHub *h = hub_of(ctx);
  tir_stream *s = new (std::nothrow) tir_stream();
  if (!s || !h) {
    delete s;
    return tir_fail(ctx, TIR_ERR_NOMEM, "out of memory");
  }
  s->ctx = ctx, s->hub = h;
  {
    std::lock_guard<std::mutex> lk(h->mu);
    h->live.push_back(s);
  }
  *out = s;
  return TIR_OK;
}

int tir_stream_feed(tir_stream *s, const int16_t *pcm, uint32_t n_samples) {
  if (!s || (!pcm && n_samples)) return TIR_ERR_ARG;
  std::lock_guard<std::mutex> lk(s->mu);
  if (s->finishing) return tir_fail(s->ctx, TIR_ERR_STATE, "tir_stream_feed after tir_stream_finish");
  if (s->failed) return tir_fail(s->ctx, TIR_ERR_CUDA, "the stream failed on the device");
  try {
    s->pending.insert(s->pending.end(), pcm, pcm + n_samples);
  } catch (const std::bad_alloc &) {
    return tir_fail(s->ctx, TIR_ERR_NOMEM, "out of memory");
  }
  s->n_samples += n_samples;
  return TIR_OK;
}

uint64_t tir_stream_samples(const tir_stream *s) { return s ? s->n_samples : 0; }
uint64_t tir_stream_frames_done(const tir_stream *s) { return s ? s->frames_pub.load(std::memory_order_acquire) : 0; }

int tir_stream_finish(tir_stream *s, int coefs, double tolerance, int freq_ignore_low, int freq_ignore_high, tir_hit *hit) {
  if (!s || !hit) return TIR_ERR_ARG;
  tir_ctx *ctx = s->ctx;
  if (coefs < 1 || coefs > TIR_N_COEFS) return tir_fail(ctx, TIR_ERR_ARG, "Wrong coefs count. max[%d], coefs[%d]", TIR_N_COEFS, coefs); // src/fp_handler.c:247
  TirStreamHub *h = s->hub;
  if (!h) return tir_fail(ctx, TIR_ERR_STATE, "the context is closing");
  {
    std::lock_guard<std::mutex> l2(s->mu);
    if (s->finishing) return tir_fail(ctx, TIR_ERR_STATE, "tir_stream_finish called twice");
    s->params.coefs = coefs, s->params.tol = tolerance, s->params.ign_lo = freq_ignore_low, s->params.ign_hi = freq_ignore_high;
    s->finishing = true;
  }
  {
    std::lock_guard<std::mutex> lk(h->mu);
    h->kick = true;
  }
  h->cv_pump.notify_all(); // do not wait for the period
  stream_wait_done(s);
  if (s->rc != TIR_OK) {
    std::lock_guard<std::mutex> lk(h->mu);
    return tir_fail(ctx, s->rc, "tir_stream_finish: %s", h->last_err.c_str());
  }
  *hit = s->hit;
  return TIR_OK;
}

void tir_stream_close(tir_stream *s) {
  if (!s) return;
  if (TirStreamHub *h = s->hub) {
    std::unique_lock<std::mutex> lk(h->mu);
    s->closing = true;
    h->cv_done.wait(lk, [&] { return !s->busy; }); // the pump may be in the middle of a batch that holds this stream
    h->live.erase(std::remove(h->live.begin(), h->live.end(), s), h->live.end());
  }
  if (s->d_coef && !(s->hub && s->hub->pool_give(s->d_coef, s->cap_frames))) {
    std::lock_guard<std::mutex> lk(s->ctx->mu);
    cudaSetDevice(s->ctx->cfg.device);
    cudaStreamSynchronize(s->ctx->stream);
    cudaFree(s->d_coef);
  }
  delete s;
}

int tir_stream_stats(tir_ctx *ctx, uint64_t *n_extract_batches, uint64_t *n_frames, uint64_t *n_match_batches) {
  if (!ctx) return TIR_ERR_ARG;
  uint64_t a = 0, b = 0, c = 0;
  {
    std::lock_guard<std::mutex> g(ctx->batcher_mu);
    if (TirStreamHub *h = ctx->stream_hub) {
      std::lock_guard<std::mutex> lk(h->mu);
      a = h->n_batches, b = h->n_hops, c = h->n_match_batches;
    }
  }
  if (n_extract_batches) *n_extract_batches = a;
  if (n_frames) *n_frames = b;
  if (n_match_batches) *n_match_batches = c;
  return TIR_OK;
}

} // extern "C"
