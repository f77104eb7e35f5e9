// tir_group.cpp -- several GPUs inside ONE process.
//
// bench.py scales with one process per GPU and NCCL (torch.distributed), but an Asterisk module is a
// single process: to use the 8 GPUs of a box it needs the same scheme without a launcher.  A
// tir_group owns one tir_ctx per device; table audio_fingerprint is sharded by uuid
// (tir_shard_of) over the contexts; a search extracts the query recordings on the first device,
// hands the coefficients to every device over NVLink (cudaMemcpyPeerAsync), lets every device match
// against its shard concurrently, pulls the per-shard winners (24 bytes per query and shard) back
// and folds them with tir_merge_hits_dev -- the same data flow as the NCCL path: only the per-query
// top-1 of every shard crosses the links.  Results equal those of one context holding the whole
// table (tests/test_gpu_group.py).
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

void tir_group_close(tir_group *g) {
  if (!g) return;
  for (size_t i = 0; i < g->ctx.size(); i++) {
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

int tir_group_search(tir_group *g, const int16_t *pcm, const uint64_t *clip_off, uint32_t n_clips, int coefs,
                     double tolerance, int freq_ignore_low, int freq_ignore_high, tir_hit *hits) {
  if (!g || !clip_off || !hits || (!pcm && n_clips && clip_off[n_clips] > clip_off[0])) return gfail(g, TIR_ERR_ARG, "null argument");
  if (coefs < 1 || coefs > TIR_N_COEFS) return gfail(g, TIR_ERR_ARG, "Wrong coefs count."); // src/fp_handler.c:247
  if (n_clips == 0) return TIR_OK;
  std::lock_guard<std::mutex> lk(g->mu);
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

} // extern "C"
