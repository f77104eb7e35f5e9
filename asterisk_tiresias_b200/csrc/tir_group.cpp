// tir_group.cpp -- several GPUs inside ONE process.
//
// bench.py scales with one process per GPU (torch.distributed), but an Asterisk module is a single
// process: to use the 8 GPUs of a box it needs the same scheme without a launcher.  A tir_group owns one
// tir_ctx per device; table audio_fingerprint is sharded by uuid (tir_shard_of) over the contexts.
// A search is the sharded search of tir_p2p_search driven from this one process (regions connected
// with tir_p2p_connect_local): the batch's clips are cut into one slice per device, every device uploads
// and extracts its slice, the coefficients cross NVLink inside the extraction kernel, every device matches
// all queries against its shard and the winners are exchanged and folded inside the match kernels.  One
// host thread enqueues all devices, so every buffer is sized BEFORE the first enqueue (tir_p2p_reserve).
// Devices without peer access (or a batch beyond the exchange buffers) take the copy-based path: extract
// on the first device, cudaMemcpyPeerAsync of coefficients and winners, tir_merge_hits_dev.
// tir_group_batcher_start / tir_group_search_one: the concurrent front-end of config[4] (1 000 dialplan
// channels) on top of the group.  Results equal those of one context holding the whole table
// (tests/test_gpu_group.py).
#include <cstring>
#include <new>

#include "tir_internal.h"

struct tir_group {
  std::vector<tir_ctx *> ctx;     // ctx[0] extracts and merges
  std::vector<float *> d_coef;    // per device: coefficients of the current batch
  std::vector<size_t> coef_cap;
  std::vector<tir_hit *> d_hits;  // per device: its shard's winners
  std::vector<size_t> hits_cap;
  tir_hit *d_gather = nullptr;    // device 0: [n_shards][n_queries]
  tir_hit *d_out = nullptr;
  size_t gather_cap = 0, out_cap = 0;
  std::vector<cudaEvent_t> ev;    // per device: "my match is done"
  cudaEvent_t ev_coef = nullptr;  // device 0: coefficients ready
  std::mutex mu;
  std::string err;
  // fused path
  std::vector<tir_p2p *> p2p;
  bool p2p_tried = false;
  uint32_t p2p_queries = 0;
  uint64_t p2p_frames = 0, p2p_samples = 0;
  std::vector<tir_hit *> d_final; // per device
  std::vector<size_t> final_cap;
  uint64_t n_fused = 0, n_copy = 0;
  TirBatcher *batcher = nullptr;
  std::mutex batcher_mu;
};

static int gfail(tir_group *g, int code, const char *msg) {
  if (g) g->err = msg ? msg : "";
  return code;
}

#define TIRG_CUDA(g, expr)                                                            \
  do {                                                                                \
    cudaError_t e_ = (expr);                                                          \
    if (e_ != cudaSuccess) return gfail((g), TIR_ERR_CUDA, cudaGetErrorString(e_));   \
  } while (0)

static int grow(tir_group *g, void **p, size_t *cap, size_t bytes, int device) {
  if (bytes <= *cap) return TIR_OK;
  TIRG_CUDA(g, cudaSetDevice(device));
  TIRG_CUDA(g, cudaDeviceSynchronize());
  if (*p) cudaFree(*p);
  *p = nullptr, *cap = 0;
  const size_t c = bytes + bytes / 4 + 256;
  TIRG_CUDA(g, cudaMalloc(p, c));
  *cap = c;
  return TIR_OK;
}

extern "C" {

static void group_drop_p2p(tir_group *g) {
  for (size_t i = 0; i < g->ctx.size(); i++)
    if (g->ctx[i]) cudaSetDevice(g->ctx[i]->cfg.device), cudaDeviceSynchronize();
  for (tir_p2p *p : g->p2p) tir_p2p_destroy(p);
  g->p2p.clear();
}

void tir_group_close(tir_group *g) {
  if (!g) return;
  if (g->batcher) tir_batcher_destroy(g->batcher), g->batcher = nullptr;
  group_drop_p2p(g);
  for (size_t i = 0; i < g->ctx.size(); i++) {
    if (g->ctx[i] && i < g->d_final.size() && g->d_final[i]) cudaSetDevice(g->ctx[i]->cfg.device), cudaFree(g->d_final[i]);
    if (g->ctx[i]) cudaSetDevice(g->ctx[i]->cfg.device), cudaDeviceSynchronize();
    if (i < g->d_coef.size() && g->d_coef[i]) cudaFree(g->d_coef[i]);
    if (i < g->d_hits.size() && g->d_hits[i]) cudaFree(g->d_hits[i]);
    if (i < g->ev.size() && g->ev[i]) cudaEventDestroy(g->ev[i]);
  }
  if (!g->ctx.empty() && g->ctx[0]) {
    cudaSetDevice(g->ctx[0]->cfg.device);
    if (g->d_gather) cudaFree(g->d_gather);
    if (g->d_out) cudaFree(g->d_out);
    if (g->ev_coef) cudaEventDestroy(g->ev_coef);
  }
  for (tir_ctx *c : g->ctx) tir_close(c);
  delete g;
}

int tir_group_open(const tir_cfg *cfg, const int *devices, int n_devices, tir_group **out) {
  if (!cfg || !devices || n_devices <= 0 || !out) return TIR_ERR_ARG;
  *out = nullptr;
  tir_group *g = new (std::nothrow) tir_group();
  if (!g) return TIR_ERR_NOMEM;
  *out = g; // handed back even on failure so that tir_group_last_error() can be read
  g->d_coef.assign(n_devices, nullptr), g->coef_cap.assign(n_devices, 0);
  g->d_hits.assign(n_devices, nullptr), g->hits_cap.assign(n_devices, 0);
  g->ev.assign(n_devices, nullptr);
  g->d_final.assign(n_devices, nullptr), g->final_cap.assign(n_devices, 0);
  for (int i = 0; i < n_devices; i++) {
    tir_cfg c = *cfg;
    c.device = devices[i], c.stream = nullptr; // every context runs on a stream of its own
    tir_ctx *ctx = nullptr;
    const int rc = tir_open(&c, &ctx);
    g->ctx.push_back(ctx);
    if (rc != TIR_OK) return gfail(g, rc, tir_last_error(ctx));
    TIRG_CUDA(g, cudaSetDevice(devices[i]));
    TIRG_CUDA(g, cudaEventCreateWithFlags(&g->ev[i], cudaEventDisableTiming));
    if (i > 0) { // the coefficients and the winners travel device to device
      int can = 0;
      if (devices[i] != devices[0]) cudaDeviceCanAccessPeer(&can, devices[i], devices[0]);
      if (can) {
        cudaError_t e = cudaDeviceEnablePeerAccess(devices[0], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return gfail(g, TIR_ERR_CUDA, cudaGetErrorString(e));
        cudaGetLastError();
        cudaSetDevice(devices[0]);
        e = cudaDeviceEnablePeerAccess(devices[i], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return gfail(g, TIR_ERR_CUDA, cudaGetErrorString(e));
        cudaGetLastError();
      } // without peer access cudaMemcpyPeerAsync stages through the host: slower, still correct
    }
  }
  TIRG_CUDA(g, cudaSetDevice(devices[0]));
  TIRG_CUDA(g, cudaEventCreateWithFlags(&g->ev_coef, cudaEventDisableTiming));
  return TIR_OK;
}

const char *tir_group_last_error(tir_group *g) { return g ? g->err.c_str() : "null group"; }
int tir_group_size(tir_group *g) { return g ? (int)g->ctx.size() : 0; }
tir_ctx *tir_group_ctx(tir_group *g, int i) { return (g && i >= 0 && i < (int)g->ctx.size()) ? g->ctx[i] : nullptr; }

int tir_group_db_load(tir_group *g, uint32_t n_audio, const uint8_t (*uuid)[16], const uint64_t *row_off, const int32_t *v1,
                      const int32_t *v2) {
  if (!g || (n_audio && (!uuid || !row_off))) return gfail(g, TIR_ERR_ARG, "null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  const uint32_t S = (uint32_t)g->ctx.size();
  std::vector<std::vector<uint8_t>> uu(S);
  std::vector<std::vector<uint64_t>> off(S, std::vector<uint64_t>(1, 0));
  std::vector<std::vector<int32_t>> a1(S), a2(S);
  for (uint32_t a = 0; a < n_audio; a++) {
    const uint32_t s = tir_shard_of(uuid[a], S);
    uu[s].insert(uu[s].end(), uuid[a], uuid[a] + 16);
    a1[s].insert(a1[s].end(), v1 + row_off[a], v1 + row_off[a + 1]);
    a2[s].insert(a2[s].end(), v2 + row_off[a], v2 + row_off[a + 1]);
    off[s].push_back(a1[s].size());
  }
  for (uint32_t s = 0; s < S; s++) {
    const int rc = tir_db_load(g->ctx[s], (uint32_t)(uu[s].size() / 16), (const uint8_t(*)[16])uu[s].data(), off[s].data(),
                               a1[s].data(), a2[s].data());
    if (rc != TIR_OK) return gfail(g, rc, tir_last_error(g->ctx[s]));
  }
  return TIR_OK;
}

int tir_group_db_add(tir_group *g, const uint8_t uuid[16], const int32_t *v1, const int32_t *v2, uint32_t n_rows) {
  if (!g || !uuid) return gfail(g, TIR_ERR_ARG, "null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  tir_ctx *c = g->ctx[tir_shard_of(uuid, (uint32_t)g->ctx.size())];
  const int rc = tir_db_add(c, uuid, v1, v2, n_rows);
  return rc == TIR_OK ? rc : gfail(g, rc, tir_last_error(c));
}

int tir_group_db_remove(tir_group *g, const uint8_t uuid[16]) {
  if (!g || !uuid) return gfail(g, TIR_ERR_ARG, "null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  tir_ctx *c = g->ctx[tir_shard_of(uuid, (uint32_t)g->ctx.size())];
  const int rc = tir_db_remove(c, uuid);
  return rc == TIR_OK ? rc : gfail(g, rc, tir_last_error(c));
}

int tir_group_db_stats(tir_group *g, uint64_t *n_audio, uint64_t *n_rows) {
  if (!g) return TIR_ERR_ARG;
  uint64_t a = 0, r = 0;
  for (tir_ctx *c : g->ctx) {
    uint64_t x = 0, y = 0;
    tir_db_stats(c, &x, &y);
    a += x, r += y;
  }
  if (n_audio) *n_audio = a;
  if (n_rows) *n_rows = r;
  return TIR_OK;
}

// (re)create the exchange objects for batches of up to n_queries / n_frames / n_samples; false: no peer access
static bool group_ensure_p2p(tir_group *g, uint32_t n_queries, uint64_t n_frames, uint64_t n_samples) {
  const int S = (int)g->ctx.size();
  if (S < 2) return false;
  // Kernels of different shards wait for one another (flags over peer memory): that needs one DEVICE per shard.
  // Several shards on one device (tests; a box with fewer GPUs than configured) would rely on the streams of one
  // device running concurrently, which nothing guarantees (shared hardware queues): they take the copy path.
  for (int a = 0; a < S; a++)
    for (int b = a + 1; b < S; b++)
      if (g->ctx[a]->cfg.device == g->ctx[b]->cfg.device) return false;
  if (!g->p2p.empty() && n_queries <= g->p2p_queries && n_frames <= g->p2p_frames && n_samples <= g->p2p_samples) return true;
  if (g->p2p_tried && g->p2p.empty()) return false; // tried before: these devices cannot reach each other
  g->p2p_tried = true;
  group_drop_p2p(g);
  const uint32_t q = std::max<uint32_t>(4096, n_queries * 2);
  const uint64_t f = std::max<uint64_t>(1u << 20, n_frames * 2), smp = std::max<uint64_t>(64u << 20, n_samples * 2);
  std::vector<tir_p2p *> ps(S, nullptr);
  bool ok = true;
  for (int s = 0; s < S && ok; s++) ok = tir_p2p_create2(g->ctx[s], s, S, q, f, &ps[s]) == TIR_OK;
  for (int s = 0; s < S && ok; s++) ok = tir_p2p_connect_local(ps[s], ps.data()) == TIR_OK;
  for (int s = 0; s < S && ok; s++) ok = tir_p2p_reserve(ps[s], smp) == TIR_OK;
  if (!ok) {
    for (tir_p2p *p : ps) tir_p2p_destroy(p);
    return false;
  }
  g->p2p = ps, g->p2p_queries = q, g->p2p_frames = f, g->p2p_samples = smp;
  return true;
}

static int group_search_copy_path(tir_group *g, const int16_t *pcm, const uint64_t *clip_off, uint32_t n_clips, int coefs,
                                  double tolerance, int freq_ignore_low, int freq_ignore_high, tir_hit *hits);

int tir_group_search(tir_group *g, const int16_t *pcm, const uint64_t *clip_off, uint32_t n_clips, int coefs,
                     double tolerance, int freq_ignore_low, int freq_ignore_high, tir_hit *hits) {
  if (!g || !clip_off || !hits || (!pcm && n_clips && clip_off[n_clips] > clip_off[0])) return gfail(g, TIR_ERR_ARG, "null argument");
  if (coefs < 1 || coefs > TIR_N_COEFS) return gfail(g, TIR_ERR_ARG, "Wrong coefs count."); // src/fp_handler.c:247
  if (n_clips == 0) return TIR_OK;
  std::lock_guard<std::mutex> lk(g->mu);
  const int S = (int)g->ctx.size();
  const int hop = g->ctx[0]->cfg.hop;
  std::vector<uint64_t> foff((size_t)n_clips + 1, 0);
  for (uint32_t c = 0; c < n_clips; c++) {
    if (clip_off[c + 1] < clip_off[c]) return gfail(g, TIR_ERR_ARG, "clip_off must be non-decreasing");
    foff[c + 1] = foff[c] + tir_n_frames(clip_off[c + 1] - clip_off[c], hop);
  }
  const uint64_t total = clip_off[n_clips] - clip_off[0];
  if (!group_ensure_p2p(g, n_clips, foff[n_clips], total)) {
    g->n_copy++;
    return group_search_copy_path(g, pcm, clip_off, n_clips, coefs, tolerance, freq_ignore_low, freq_ignore_high, hits);
  }
  // ---- fused path: one slice of the clips per device (balanced by samples), all devices enqueued by this thread
  std::vector<uint32_t> cut((size_t)S + 1, n_clips);
  cut[0] = 0;
  for (int s = 1; s < S; s++) {
    const uint64_t want = clip_off[0] + total * (uint64_t)s / (uint64_t)S;
    uint32_t c = cut[s - 1];
    while (c < n_clips && clip_off[c] < want) c++;
    cut[s] = c;
  }
  int rc;
  const size_t hit_bytes = (size_t)n_clips * sizeof(tir_hit);
  for (int s = 0; s < S; s++) {
    if ((rc = grow(g, (void **)&g->d_final[s], &g->final_cap[s], hit_bytes, g->ctx[s]->cfg.device))) return rc;
    if (tir_db_ensure_index_public(g->ctx[s]) != TIR_OK) return gfail(g, TIR_ERR_CUDA, tir_last_error(g->ctx[s])); // no sort between the enqueues
  }
  for (int s = 0; s < S; s++) {
    const uint32_t a = cut[s], b = cut[s + 1];
    rc = tir_p2p_search(g->p2p[s], pcm, clip_off + a, b - a, a, foff.data(), n_clips, coefs, tolerance, freq_ignore_low, freq_ignore_high,
                        nullptr, g->d_final[s]);
    if (rc != TIR_OK) { // (the ranks already enqueued complete through the time-out of their waits)
      gfail(g, rc, tir_last_error(g->ctx[s]));
      group_drop_p2p(g), g->p2p_tried = false;
      return rc;
    }
  }
  tir_ctx *c0 = g->ctx[0];
  TIRG_CUDA(g, cudaSetDevice(c0->cfg.device));
  TIRG_CUDA(g, cudaMemcpyAsync(hits, g->d_final[0], hit_bytes, cudaMemcpyDeviceToHost, c0->stream));
  TIRG_CUDA(g, cudaStreamSynchronize(c0->stream));
  uint32_t bad = 0;
  tir_p2p_error(g->p2p[0], &bad);
  if (bad) {
    if (getenv("TIR_DEBUG")) {
      for (int s = 0; s < S; s++) {
        uint32_t e = 0;
        tir_p2p_error(g->p2p[s], &e);
        fprintf(stderr, "[tir_group] rank %d: clips [%u, %u), error word %u\n", s, cut[s], cut[s + 1], e);
      }
    }
    return gfail(g, TIR_ERR_CUDA, "a device did not answer in time (tir_p2p_error)");
  }
  g->n_fused++;
  return TIR_OK;
}

static int group_search_copy_path(tir_group *g, const int16_t *pcm, const uint64_t *clip_off, uint32_t n_clips, int coefs,
                                  double tolerance, int freq_ignore_low, int freq_ignore_high, tir_hit *hits) {
  const int S = (int)g->ctx.size();
  tir_ctx *c0 = g->ctx[0];
  const int dev0 = c0->cfg.device;
  const uint64_t base = clip_off[0], total = clip_off[n_clips] - base;
  std::vector<uint64_t> rel((size_t)n_clips + 1), foff((size_t)n_clips + 1, 0);
  for (uint32_t c = 0; c <= n_clips; c++) rel[c] = clip_off[c] - base;
  for (uint32_t c = 0; c < n_clips; c++) foff[c + 1] = foff[c] + tir_n_frames(rel[c + 1] - rel[c], c0->cfg.hop);
  const uint64_t F = foff[n_clips];
  const size_t coef_bytes = (size_t)(F ? F : 1) * TIR_N_COEFS * sizeof(float), hit_bytes = (size_t)n_clips * sizeof(tir_hit);
  int rc;
  for (int s = 0; s < S; s++) {
    if ((rc = grow(g, (void **)&g->d_coef[s], &g->coef_cap[s], coef_bytes, g->ctx[s]->cfg.device))) return rc;
    if ((rc = grow(g, (void **)&g->d_hits[s], &g->hits_cap[s], hit_bytes, g->ctx[s]->cfg.device))) return rc;
  }
  if ((rc = grow(g, (void **)&g->d_gather, &g->gather_cap, hit_bytes * S, dev0))) return rc;
  if ((rc = grow(g, (void **)&g->d_out, &g->out_cap, hit_bytes, dev0))) return rc;
  // ---- device 0: PCM in, extraction
  {
    std::lock_guard<std::mutex> l0(c0->mu);
    TIRG_CUDA(g, cudaSetDevice(dev0));
    if ((rc = tir_reserve(c0, c0->d_pcm, total * sizeof(int16_t) + 16))) return gfail(g, rc, tir_last_error(c0));
    if (total) TIRG_CUDA(g, cudaMemcpyAsync(c0->d_pcm.p, pcm + base, total * sizeof(int16_t), cudaMemcpyHostToDevice, c0->stream));
    if ((rc = tir_extract_launch(c0, (const int16_t *)c0->d_pcm.p, total, rel.data(), n_clips, g->d_coef[0], nullptr, nullptr)))
      return gfail(g, rc, tir_last_error(c0));
    TIRG_CUDA(g, cudaEventRecord(g->ev_coef, c0->stream));
  }
  // ---- every device: coefficients over NVLink, match against its shard (all queued without waiting)
  for (int s = 0; s < S; s++) {
    tir_ctx *c = g->ctx[s];
    TIRG_CUDA(g, cudaSetDevice(c->cfg.device));
    if (s > 0) {
      TIRG_CUDA(g, cudaStreamWaitEvent(c->stream, g->ev_coef, 0));
      TIRG_CUDA(g, cudaMemcpyPeerAsync(g->d_coef[s], c->cfg.device, g->d_coef[0], dev0, coef_bytes, c->stream));
    }
    uint64_t na = 0, nr = 0;
    tir_db_stats(c, &na, &nr);
    if (na == 0) { // an empty shard has no winners
      TIRG_CUDA(g, cudaMemsetAsync(g->d_hits[s], 0, hit_bytes, c->stream));
    } else if ((rc = tir_match_dev(c, g->d_coef[s], foff.data(), n_clips, coefs, tolerance, freq_ignore_low, freq_ignore_high,
                                   g->d_hits[s]))) {
      return gfail(g, rc, tir_last_error(c));
    }
    TIRG_CUDA(g, cudaMemcpyPeerAsync(g->d_gather + (size_t)s * n_clips, dev0, g->d_hits[s], c->cfg.device, hit_bytes, c->stream));
    TIRG_CUDA(g, cudaEventRecord(g->ev[s], c->stream));
  }
  // ---- device 0: fold the S winners of every query, hand them to the caller
  TIRG_CUDA(g, cudaSetDevice(dev0));
  for (int s = 1; s < S; s++) TIRG_CUDA(g, cudaStreamWaitEvent(c0->stream, g->ev[s], 0));
  if ((rc = tir_merge_hits_dev(c0, g->d_gather, (uint32_t)S, n_clips, g->d_out))) return gfail(g, rc, tir_last_error(c0));
  TIRG_CUDA(g, cudaMemcpyAsync(hits, g->d_out, hit_bytes, cudaMemcpyDeviceToHost, c0->stream));
  TIRG_CUDA(g, cudaStreamSynchronize(c0->stream));
  for (uint32_t q = 0; q < n_clips; q++) hits[q].frame_count = (int32_t)(foff[q + 1] - foff[q]); // also when no shard had a row
  return TIR_OK;
}

int tir_group_stats(tir_group *g, uint64_t *n_fused, uint64_t *n_copy_path) {
  if (!g) return TIR_ERR_ARG;
  std::lock_guard<std::mutex> lk(g->mu);
  if (n_fused) *n_fused = g->n_fused;
  if (n_copy_path) *n_copy_path = g->n_copy;
  return TIR_OK;
}

int tir_group_batcher_start(tir_group *g, uint32_t max_batch, uint32_t max_wait_us) {
  if (!g || max_batch == 0) return gfail(g, TIR_ERR_ARG, "bad batcher arguments");
  std::lock_guard<std::mutex> lk(g->batcher_mu);
  if (g->batcher) return gfail(g, TIR_ERR_STATE, "batcher already running");
  g->batcher = tir_batcher_create(
      g->ctx[0]->cfg.device, max_batch, max_wait_us,
      [g](const int16_t *pcm, const uint64_t *off, uint32_t n, int coefs, double tol, int lo, int hi, tir_hit *hits) {
        return tir_group_search(g, pcm, off, n, coefs, tol, lo, hi, hits);
      },
      [g] { return std::string(tir_group_last_error(g)); });
  return g->batcher ? TIR_OK : gfail(g, TIR_ERR_NOMEM, "could not start the batcher");
}

int tir_group_batcher_stop(tir_group *g) {
  if (!g) return TIR_ERR_ARG;
  TirBatcher *b;
  {
    std::lock_guard<std::mutex> lk(g->batcher_mu);
    b = g->batcher, g->batcher = nullptr;
  }
  tir_batcher_destroy(b);
  return TIR_OK;
}

int tir_group_search_one(tir_group *g, const int16_t *pcm, uint64_t n_samples, int coefs, double tolerance, int freq_ignore_low,
                         int freq_ignore_high, tir_hit *hit) {
  if (!g || !hit || (!pcm && n_samples)) return gfail(g, TIR_ERR_ARG, "null argument");
  if (coefs < 1 || coefs > TIR_N_COEFS) return gfail(g, TIR_ERR_ARG, "Wrong coefs count."); // src/fp_handler.c:247
  TirBatcher *b;
  {
    std::lock_guard<std::mutex> lk(g->batcher_mu);
    b = g->batcher;
    if (b) tir_batcher_enter(b);
  }
  const uint64_t off[2] = {0, n_samples};
  if (!b || !tir_batcher_fits(b, n_samples)) {
    if (b) tir_batcher_leave(b);
    return tir_group_search(g, pcm, off, 1, coefs, tolerance, freq_ignore_low, freq_ignore_high, hit);
  }
  std::string err;
  const int rc = tir_batcher_submit(b, pcm, n_samples, coefs, tolerance, freq_ignore_low, freq_ignore_high, hit, &err);
  tir_batcher_leave(b);
  return rc == TIR_OK ? rc : gfail(g, rc, err.c_str());
}

int tir_group_batcher_stats(tir_group *g, uint64_t *n_requests, uint64_t *n_batches, uint64_t *max_batch_seen) {
  if (!g) return TIR_ERR_ARG;
  uint64_t r = 0, n = 0, m = 0;
  {
    std::lock_guard<std::mutex> lk(g->batcher_mu);
    if (g->batcher) tir_batcher_counters(g->batcher, &r, &n, &m);
  }
  if (n_requests) *n_requests = r;
  if (n_batches) *n_batches = n;
  if (max_batch_seen) *max_batch_seen = m;
  return TIR_OK;
}

} // extern "C"
