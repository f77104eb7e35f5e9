// tir_api.cu -- C ABI entry points of libtiresias_gpu.so (context, extraction); the device DB and
// the match entry points live in tir_match.cu.  See include/tiresias_gpu.h.
#include <cuda_runtime.h>

#include <cstring>
#include <new>

#include "tir_internal.h"

int tir_fail(tir_ctx *ctx, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) {
    std::lock_guard<std::mutex> lk(ctx->err_mu);
    ctx->err = buf;
  }
  return code;
}

int tir_reserve(tir_ctx *ctx, DevBuf &b, size_t bytes) {
  if (bytes <= b.cap) return TIR_OK;
  if (b.p) {
    TIR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    TIR_CUDA(ctx, cudaFree(b.p));
    b.p = nullptr, b.cap = 0;
  }
  size_t cap = bytes + bytes / 8 + 256;
  cudaError_t e = cudaMalloc(&b.p, cap);
  if (e != cudaSuccess) {
    b.p = nullptr;
    return tir_fail(ctx, e == cudaErrorMemoryAllocation ? TIR_ERR_NOMEM : TIR_ERR_CUDA, "cudaMalloc(%zu): %s", cap,
                    cudaGetErrorString(e));
  }
  b.cap = cap;
  return TIR_OK;
}

int tir_reserve_host(tir_ctx *ctx, DevBuf &b, size_t bytes) {
  if (bytes <= b.cap) return TIR_OK;
  if (b.p) {
    TIR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    TIR_CUDA(ctx, cudaFreeHost(b.p));
    b.p = nullptr, b.cap = 0;
  }
  size_t cap = bytes + bytes / 8 + 256;
  cudaError_t e = cudaMallocHost(&b.p, cap);
  if (e != cudaSuccess) {
    b.p = nullptr;
    return tir_fail(ctx, TIR_ERR_NOMEM, "cudaMallocHost(%zu): %s", cap, cudaGetErrorString(e));
  }
  b.cap = cap;
  return TIR_OK;
}

int tir_stage_acquire(tir_ctx *ctx, size_t bytes, void **p, int *slot) {
  const int k = ctx->h_stage_next;
  ctx->h_stage_next = (k + 1) % tir_ctx::kStageSlots;
  if (!ctx->h_stage_ev[k]) TIR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->h_stage_ev[k], cudaEventDisableTiming));
  if (ctx->h_stage_used[k]) TIR_CUDA(ctx, cudaEventSynchronize(ctx->h_stage_ev[k])); // four calls ago: normally long done
  int rc;
  if ((rc = tir_reserve_host(ctx, ctx->h_stage[k], bytes))) return rc;
  *p = ctx->h_stage[k].p, *slot = k;
  return TIR_OK;
}

int tir_stage_release(tir_ctx *ctx, int slot) {
  TIR_CUDA(ctx, cudaEventRecord(ctx->h_stage_ev[slot], ctx->stream));
  ctx->h_stage_used[slot] = true;
  return TIR_OK;
}

static void free_dev(DevBuf &b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr, b.cap = 0;
}

extern "C" {

int tir_abi_version(void) { return TIR_ABI_VERSION; }

void tir_cfg_default(tir_cfg *cfg) {
  if (!cfg) return;
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->device = 0;
  cfg->win = 512, cfg->hop = 256, cfg->n_filters = 40, cfg->samplerate = 8000;
  cfg->stream = nullptr;
}

uint64_t tir_n_frames(uint64_t n_samples, int hop) {
  if (hop <= 0) return 0;
  return (n_samples + (uint64_t)hop - 1) / (uint64_t)hop;
}

static int upload(tir_ctx *ctx, float4 **dst, const std::vector<float4> &src) {
  TIR_CUDA(ctx, cudaMalloc((void **)dst, src.size() * sizeof(float4)));
  TIR_CUDA(ctx, cudaMemcpy(*dst, src.data(), src.size() * sizeof(float4), cudaMemcpyHostToDevice));
  return TIR_OK;
}

int tir_open(const tir_cfg *cfg, tir_ctx **out) {
  if (!cfg || !out) return TIR_ERR_ARG;
  *out = nullptr;
  tir_ctx *ctx = new (std::nothrow) tir_ctx();
  if (!ctx) return TIR_ERR_NOMEM;
  ctx->cfg = *cfg;
  *out = ctx; // handed back even on failure so that tir_last_error() can be read; tir_close() frees it
  if (!tir_build_tables(cfg->win, cfg->hop, cfg->n_filters, TIR_N_COEFS, cfg->samplerate, ctx->tab))
    return tir_fail(ctx, TIR_ERR_ARG, "unsupported plan win=%d hop=%d filters=%d rate=%d", cfg->win, cfg->hop,
                    cfg->n_filters, cfg->samplerate);
  int ndev = 0;
  TIR_CUDA(ctx, cudaGetDeviceCount(&ndev));
  if (ndev <= 0) return tir_fail(ctx, TIR_ERR_CUDA, "no CUDA device: libtiresias_gpu has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return tir_fail(ctx, TIR_ERR_ARG, "device %d of %d", cfg->device, ndev);
  TIR_CUDA(ctx, cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  TIR_CUDA(ctx, cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10)
    return tir_fail(ctx, TIR_ERR_CUDA, "device %s is sm_%d%d; this library carries sm_100a code only", prop.name,
                    prop.major, prop.minor);
  ctx->num_sms = prop.multiProcessorCount;
  if (cfg->stream) {
    ctx->stream = (cudaStream_t)cfg->stream;
  } else {
    TIR_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
  }
  int rc;
  if ((rc = upload(ctx, &ctx->d_win4, ctx->tab.win4))) return rc;
  if ((rc = upload(ctx, &ctx->d_twp4, ctx->tab.twp4))) return rc;
  if ((rc = upload(ctx, &ctx->d_twu4, ctx->tab.twu4))) return rc;
  for (int w = 0; w < 2; w++)
    for (int e = 0; e < 2; e++) TIR_CUDA(ctx, cudaEventCreate(&ctx->ev[w][e]));
  return TIR_OK;
}

void tir_close(tir_ctx *ctx) {
  if (!ctx) return;
  if (ctx->stream_hub) tir_stream_hub_destroy(ctx->stream_hub), ctx->stream_hub = nullptr;
  if (ctx->batcher) tir_batcher_destroy(ctx->batcher), ctx->batcher = nullptr; // serves what is queued, then joins
  if (ctx->num_sms) {
    cudaSetDevice(ctx->cfg.device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  }
  if (ctx->db) tir_db_destroy(ctx->db);
  cudaFree(ctx->d_win4), cudaFree(ctx->d_twp4), cudaFree(ctx->d_twu4);
  free_dev(ctx->d_clipmeta), free_dev(ctx->d_tilemeta), free_dev(ctx->d_pcm), free_dev(ctx->d_coef);
  free_dev(ctx->d_vq), free_dev(ctx->d_qmeta), free_dev(ctx->d_qmeta2), free_dev(ctx->d_hits), free_dev(ctx->d_hits2), free_dev(ctx->d_y), free_dev(ctx->d_counter), free_dev(ctx->d_ulaw), free_dev(ctx->d_mix), free_dev(ctx->d_items);
  for (int k = 0; k < tir_ctx::kStageSlots; k++) {
    if (ctx->h_stage[k].p) cudaFreeHost(ctx->h_stage[k].p);
    if (ctx->h_stage_ev[k]) cudaEventDestroy(ctx->h_stage_ev[k]);
  }
  for (int w = 0; w < 2; w++)
    for (int e = 0; e < 2; e++)
      if (ctx->ev[w][e]) cudaEventDestroy(ctx->ev[w][e]);
  if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
  if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

// The text is copied into a buffer owned by the calling thread: it stays valid until the same thread
// asks again, whatever the other threads of the context do meanwhile.
const char *tir_last_error(tir_ctx *ctx) {
  if (!ctx) return "null context";
  static thread_local std::string mine;
  std::lock_guard<std::mutex> lk(ctx->err_mu);
  mine = ctx->err;
  return mine.c_str();
}

uint64_t tir_launch_count(tir_ctx *ctx) { return ctx ? ctx->launches : 0; }

void tir_set_profiling(tir_ctx *ctx, int on) {
  if (ctx) ctx->profiling = on != 0;
}

float tir_last_kernel_ms(tir_ctx *ctx, int which) {
  if (!ctx || which < 0 || which > 1 || !ctx->ev_valid[which]) return -1.f;
  std::lock_guard<std::mutex> lk(ctx->mu);
  cudaSetDevice(ctx->cfg.device);
  float ms = -1.f;
  if (cudaEventSynchronize(ctx->ev[which][1]) != cudaSuccess) return -1.f;
  if (cudaEventElapsedTime(&ms, ctx->ev[which][0], ctx->ev[which][1]) != cudaSuccess) return -1.f;
  return ms;
}

int tir_selftest(tir_ctx *ctx, uint64_t *sqrt_mismatches, uint32_t first_bits, uint32_t step, uint32_t count, float *log10f_out) {
  if (!ctx) return TIR_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  TIR_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
  return tir_selftest_launch(ctx, sqrt_mismatches, first_bits, step, count, log10f_out);
}

int tir_get_tables(tir_ctx *ctx, float *window, float *filters, float *dct) {
  if (!ctx) return TIR_ERR_ARG;
  if (window) std::memcpy(window, ctx->tab.window.data(), ctx->tab.window.size() * sizeof(float));
  if (filters) std::memcpy(filters, ctx->tab.filters.data(), ctx->tab.filters.size() * sizeof(float));
  if (dct) std::memcpy(dct, ctx->tab.dct.data(), ctx->tab.dct.size() * sizeof(float));
  return TIR_OK;
}

int tir_extract_dev(tir_ctx *ctx, const int16_t *d_pcm, const uint64_t *clip_off, uint32_t n_clips, float *d_coef,
                    int32_t *d_vq, uint64_t *n_frames) {
  if (!ctx || !clip_off || (!d_pcm && n_clips && clip_off[n_clips] > clip_off[0])) return tir_fail(ctx, TIR_ERR_ARG, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  TIR_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
  return tir_extract_launch(ctx, d_pcm, clip_off[n_clips], clip_off, n_clips, d_coef, d_vq, n_frames);
}

// host buffers in, host buffers out; `src` is PCM16 (ulaw == false) or G.711 mu-law bytes
static int extract_host(tir_ctx *ctx, const void *src, bool ulaw, const uint64_t *clip_off, uint32_t n_clips, float *coef,
                        int32_t *vq, uint64_t *n_frames) {
  if (!ctx || !clip_off || (!src && n_clips && clip_off[n_clips] > clip_off[0])) return tir_fail(ctx, TIR_ERR_ARG, "null argument");
  const int16_t *pcm = ulaw ? nullptr : (const int16_t *)src;
  const uint8_t *law = ulaw ? (const uint8_t *)src : nullptr;
  std::lock_guard<std::mutex> lk(ctx->mu);
  TIR_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
  const uint64_t base = clip_off[0], total = clip_off[n_clips] - base;
  uint64_t F = 0;
  for (uint32_t c = 0; c < n_clips; c++) {
    if (clip_off[c + 1] < clip_off[c]) return tir_fail(ctx, TIR_ERR_ARG, "clip_off must be non-decreasing");
    F += tir_n_frames(clip_off[c + 1] - clip_off[c], ctx->cfg.hop);
  }
  if (n_frames) *n_frames = F;
  if (F == 0) return TIR_OK;
  int rc;
  if ((rc = tir_reserve(ctx, ctx->d_pcm, total * sizeof(int16_t) + 16))) return rc;
  if ((rc = tir_reserve(ctx, ctx->d_coef, F * TIR_N_COEFS * sizeof(float)))) return rc;
  if ((rc = tir_reserve(ctx, ctx->d_vq, F * TIR_N_COEFS * sizeof(int32_t)))) return rc;
  if (ulaw && (rc = tir_reserve(ctx, ctx->d_ulaw, total + 16))) return rc;
  uint8_t *d_law = (uint8_t *)ctx->d_ulaw.p;
  // The batch is cut into chunks of whole clips (~96 MB of PCM) that flow through three queues:
  // copy-in stream -> ctx's stream (kernel) -> copy-out stream, so the PCIe transfers of one chunk
  // overlap the kernel of another.
  if (!ctx->s_in) {
    TIR_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
    TIR_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
  }
  struct Chunk { uint32_t c0, c1; uint64_t f0, f1; cudaEvent_t in, done; };
  std::vector<Chunk> chunks;
  const uint64_t target = 48ull << 20; // samples
  for (uint32_t c = 0; c < n_clips;) {
    Chunk k{c, c, 0, 0, nullptr, nullptr};
    while (k.c1 < n_clips && (k.c1 == k.c0 || clip_off[k.c1 + 1] - clip_off[k.c0] <= target)) k.c1++;
    chunks.push_back(k);
    c = k.c1;
  }
  uint64_t f = 0;
  for (Chunk &k : chunks) {
    k.f0 = f;
    for (uint32_t c = k.c0; c < k.c1; c++) f += tir_n_frames(clip_off[c + 1] - clip_off[c], ctx->cfg.hop);
    k.f1 = f;
    TIR_CUDA(ctx, cudaEventCreateWithFlags(&k.in, cudaEventDisableTiming));
    TIR_CUDA(ctx, cudaEventCreateWithFlags(&k.done, cudaEventDisableTiming));
  }
  auto cleanup = [&] {
    for (Chunk &k : chunks) {
      if (k.in) cudaEventDestroy(k.in);
      if (k.done) cudaEventDestroy(k.done);
    }
  };
  int16_t *d_pcm = (int16_t *)ctx->d_pcm.p;
  float *d_coef = (float *)ctx->d_coef.p;
  int32_t *d_vq = (int32_t *)ctx->d_vq.p;
  cudaError_t e = cudaSuccess;
  // every host->device copy is queued up front: the copy engine never waits for the host
  for (Chunk &k : chunks) {
    const uint64_t s0 = clip_off[k.c0] - base, s1 = clip_off[k.c1] - base;
    if (s1 > s0 && e == cudaSuccess)
      e = ulaw ? cudaMemcpyAsync(d_law + s0, law + base + s0, s1 - s0, cudaMemcpyHostToDevice, ctx->s_in)
               : cudaMemcpyAsync(d_pcm + s0, pcm + base + s0, (s1 - s0) * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->s_in);
    if (e == cudaSuccess) e = cudaEventRecord(k.in, ctx->s_in);
  }
  std::vector<uint64_t> rel;
  for (Chunk &k : chunks) {
    if (e != cudaSuccess) break;
    e = cudaStreamWaitEvent(ctx->stream, k.in, 0);
    if (e != cudaSuccess) break;
    if (ulaw) { // bytes -> PCM16 on the device: half the PCIe traffic of the PCM16 entry point
      const uint64_t s0 = clip_off[k.c0] - base, s1 = clip_off[k.c1] - base;
      if ((rc = tir_ulaw_decode_launch(ctx, d_law + s0, d_pcm + s0, s1 - s0))) {
        cudaDeviceSynchronize();
        cleanup();
        return rc;
      }
    }
    // offsets stay relative to the whole staged batch: a clip keeps the alignment it has in the caller's buffer
    rel.assign((size_t)(k.c1 - k.c0) + 1, 0);
    for (uint32_t c = k.c0; c <= k.c1; c++) rel[c - k.c0] = clip_off[c] - base;
    if ((rc = tir_extract_launch(ctx, d_pcm, total, rel.data(), k.c1 - k.c0, d_coef + k.f0 * TIR_N_COEFS,
                                 d_vq + k.f0 * TIR_N_COEFS, nullptr))) {
      cudaDeviceSynchronize();
      cleanup();
      return rc;
    }
    e = cudaEventRecord(k.done, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->s_out, k.done, 0);
    const size_t n = (size_t)(k.f1 - k.f0) * TIR_N_COEFS;
    if (coef && n && e == cudaSuccess)
      e = cudaMemcpyAsync(coef + k.f0 * TIR_N_COEFS, d_coef + k.f0 * TIR_N_COEFS, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->s_out);
    if (vq && n && e == cudaSuccess)
      e = cudaMemcpyAsync(vq + k.f0 * TIR_N_COEFS, d_vq + k.f0 * TIR_N_COEFS, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->s_out);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->s_out);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) cudaDeviceSynchronize();
  cleanup();
  if (e != cudaSuccess) return tir_fail(ctx, TIR_ERR_CUDA, "tir_extract: %s", cudaGetErrorString(e));
  return TIR_OK;
}

int tir_extract(tir_ctx *ctx, const int16_t *pcm, const uint64_t *clip_off, uint32_t n_clips, float *coef, int32_t *vq,
                uint64_t *n_frames) {
  return extract_host(ctx, pcm, false, clip_off, n_clips, coef, vq, n_frames);
}

int tir_extract_ulaw(tir_ctx *ctx, const uint8_t *ulaw, const uint64_t *clip_off, uint32_t n_clips, float *coef, int32_t *vq,
                     uint64_t *n_frames) {
  return extract_host(ctx, ulaw, true, clip_off, n_clips, coef, vq, n_frames);
}

// Multi-channel files are an ingest-time case (a directory scan, the CLI): one copy in, the down-mix, one launch, one
// copy out -- no chunk pipeline.
int tir_extract_interleaved(tir_ctx *ctx, const int16_t *pcm, int channels, const uint64_t *clip_off, uint32_t n_clips,
                            float *coef, int32_t *vq, uint64_t *n_frames) {
  if (channels == 1) return tir_extract(ctx, pcm, clip_off, n_clips, coef, vq, n_frames);
  if (!ctx || !clip_off || (!pcm && n_clips && clip_off[n_clips] > clip_off[0])) return tir_fail(ctx, TIR_ERR_ARG, "null argument");
  if (channels < 1 || channels > 64) return tir_fail(ctx, TIR_ERR_ARG, "channels must be 1..64");
  std::lock_guard<std::mutex> lk(ctx->mu);
  TIR_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
  const uint64_t base = clip_off[0], total = clip_off[n_clips] - base;
  uint64_t F = 0;
  std::vector<uint64_t> rel((size_t)n_clips + 1, 0);
  for (uint32_t c = 0; c < n_clips; c++) {
    if (clip_off[c + 1] < clip_off[c]) return tir_fail(ctx, TIR_ERR_ARG, "clip_off must be non-decreasing");
    F += tir_n_frames(clip_off[c + 1] - clip_off[c], ctx->cfg.hop);
    rel[c + 1] = clip_off[c + 1] - base;
  }
  if (n_frames) *n_frames = F;
  if (F == 0) return TIR_OK;
  int rc;
  if ((rc = tir_reserve(ctx, ctx->d_pcm, total * (uint64_t)channels * sizeof(int16_t) + 16))) return rc;
  if ((rc = tir_reserve(ctx, ctx->d_mix, total * sizeof(float) + 16))) return rc;
  if ((rc = tir_reserve(ctx, ctx->d_coef, F * TIR_N_COEFS * sizeof(float)))) return rc;
  if ((rc = tir_reserve(ctx, ctx->d_vq, F * TIR_N_COEFS * sizeof(int32_t)))) return rc;
  int16_t *d_in = (int16_t *)ctx->d_pcm.p;
  float *d_mix = (float *)ctx->d_mix.p;
  TIR_CUDA(ctx, cudaMemcpyAsync(d_in, pcm + base * (uint64_t)channels, total * (uint64_t)channels * sizeof(int16_t),
                                cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = tir_downmix_launch(ctx, d_in, total, channels, d_mix))) return rc;
  if ((rc = tir_extract_launch_f32(ctx, d_mix, rel.data(), n_clips, (float *)ctx->d_coef.p, (int32_t *)ctx->d_vq.p, nullptr))) {
    cudaStreamSynchronize(ctx->stream);
    return rc;
  }
  const size_t n = (size_t)F * TIR_N_COEFS;
  cudaError_t e = cudaSuccess;
  if (coef) e = cudaMemcpyAsync(coef, ctx->d_coef.p, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
  if (vq && e == cudaSuccess) e = cudaMemcpyAsync(vq, ctx->d_vq.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
  const cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) e = e2;
  if (e != cudaSuccess) return tir_fail(ctx, TIR_ERR_CUDA, "tir_extract_interleaved: %s", cudaGetErrorString(e));
  return TIR_OK;
}

} // extern "C"
