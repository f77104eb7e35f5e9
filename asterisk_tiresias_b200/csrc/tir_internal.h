// tir_internal.h -- context object shared by the translation units of libtiresias_gpu.so
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <functional>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/tiresias_gpu.h"
#include "tir_tables.h"

struct TirDb;      // tir_match.cu
struct TirBatcher; // tir_batcher.cpp
struct TirStreamHub; // tir_stream.cu

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
};

struct tir_ctx {
  tir_cfg cfg{};
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t s_in = nullptr, s_out = nullptr; // copy-in / copy-out queues of tir_extract's pipeline
  std::mutex mu;
  std::mutex err_mu; // guards `err` alone: tir_fail is called with and without `mu` held, from any thread
  std::string err;
  uint64_t launches = 0;
  bool profiling = false;
  bool smem_attr_set = false;
  bool match_smem_attr_set = false;
  cudaEvent_t ev[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}}; // [which][begin/end]
  bool ev_valid[2] = {false, false};
  TirHostTables tab;
  // device copies of the kernel-layout tables
  float4 *d_win4 = nullptr, *d_twp4 = nullptr, *d_twu4 = nullptr;
  // reusable device scratch
  DevBuf d_clipmeta, d_tilemeta, d_pcm, d_coef, d_vq, d_qmeta, d_qmeta2, d_hits, d_hits2, d_y, d_counter, d_ulaw, d_mix, d_items;
  // pinned staging for small metadata: a ring, so that a call does not have to wait for the previous
  // call's copy (each slot is guarded by an event recorded after the copy that reads it)
  static constexpr int kStageSlots = 4;
  DevBuf h_stage[kStageSlots];
  cudaEvent_t h_stage_ev[kStageSlots] = {nullptr, nullptr, nullptr, nullptr};
  bool h_stage_used[kStageSlots] = {false, false, false, false};
  int h_stage_next = 0;
  TirDb *db = nullptr;
  TirBatcher *batcher = nullptr;
  TirStreamHub *stream_hub = nullptr; // streaming front-end (pump thread), created by the first tir_stream_open
  std::mutex batcher_mu; // guards `batcher` / `stream_hub` themselves (never held across GPU work)
};

int tir_fail(tir_ctx *ctx, int code, const char *fmt, ...);
#define TIR_CUDA(ctx, expr)                                                                     \
  do {                                                                                          \
    cudaError_t e_ = (expr);                                                                    \
    if (e_ != cudaSuccess)                                                                      \
      return tir_fail((ctx), TIR_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                      __FILE__, __LINE__);                                                      \
  } while (0)

int tir_reserve(tir_ctx *ctx, DevBuf &b, size_t bytes);      // device scratch, grows only
int tir_reserve_host(tir_ctx *ctx, DevBuf &b, size_t bytes); // pinned host scratch
// next staging slot, at least `bytes` large and no longer read by the device; after enqueueing the
// copy that reads it on ctx->stream call tir_stage_release(ctx, slot)
int tir_stage_acquire(tir_ctx *ctx, size_t bytes, void **p, int *slot);
int tir_stage_release(tir_ctx *ctx, int slot);

// tir_extract.cu
// Sharded search: the coefficients of a rank's own query clips are stored, as P4 computes them, into
// every rank's coefficient buffer over NVLink peer memory (frame index = frame_base + local frame), and
// the kernel's last CTA releases flag[TIR_P2P_COEF_FLAG0 + rank] = epoch on every peer.  peer == nullptr:
// no exchange.
struct TirCoefX {
  unsigned char *const *peer; // device table: base of every rank's region
  unsigned long long off;     // byte offset of this batch's coefficient buffer inside a region
  int rank, world;
  uint32_t epoch;
  uint32_t *done;             // CTA counter (left at 0)
  unsigned long long frame_base;
};
int tir_extract_launch(tir_ctx *ctx, const int16_t *d_pcm, uint64_t total_samples, const uint64_t *clip_off,
                       uint32_t n_clips, float *d_coef, int32_t *d_vq, uint64_t *n_frames, const TirCoefX *cx = nullptr);
// multi-channel input: interleaved PCM16 -> the float mean aubio's source makes of it (* 2^15), then the same kernel on floats
int tir_downmix_launch(tir_ctx *ctx, const int16_t *d_in, uint64_t n_frames, int channels, float *d_out);
int tir_extract_launch_f32(tir_ctx *ctx, const float *d_pcm, const uint64_t *clip_off, uint32_t n_clips, float *d_coef,
                           int32_t *d_vq, uint64_t *n_frames);
size_t tir_extract_smem_bytes(int win);
int tir_selftest_launch(tir_ctx *ctx, uint64_t *sqrt_mismatches, uint32_t first, uint32_t step, uint32_t count, float *log10f_out);
int tir_ulaw_decode_launch(tir_ctx *ctx, const uint8_t *d_in, int16_t *d_out, uint64_t n);

// tir_match.cu
void tir_db_destroy(TirDb *db);
// The cross-GPU exchange of the winners, fused into the kernels that produce them (tir_p2p.cu owns
// the buffers): every hit is also stored into row `rank` of every peer's gather buffer, and the last
// CTA of the producing kernel releases flag[rank] = epoch on every peer.  peer == nullptr: no exchange.
// With `final_out` set the last CTA goes on to acquire the flags of all ranks and folds the candidates
// into final_out itself (tir_p2p_dev.cuh): no separate merge launch.
struct TirP2PArgs {
  unsigned char *const *peer; // device table: base of every rank's region
  int rank, world;
  uint32_t max_queries, epoch;
  uint32_t *done; // CTA counter (left at 0)
  unsigned char *local = nullptr; // this rank's own region
  tir_hit *final_out = nullptr;   // [n_queries] global winners
  // When set, the kernels read the batch number from this device word instead of `epoch`: the match chain copies it in
  // together with the query offsets, so the kernels' arguments do not change from batch to batch and a steady caller's
  // exchange chain replays as a CUDA graph too.
  const uint32_t *epoch_dev = nullptr;
  bool may_alloc = false; // the caller drives this rank alone (one process per GPU): scratch may grow inside the call
};
#define TIR_P2P_MAX_RANKS 16
#define TIR_P2P_HDR 256 // bytes of flags before the two gather buffers of a region:
                        // u32[0..15] winners flags | u32[16] error word | u32[32..47] coefficient flags
#define TIR_P2P_ERR_WORD 16
#define TIR_P2P_COEF_FLAG0 32
int tir_match_dev_exchange(tir_ctx *ctx, const float *d_coef, const uint64_t *frame_off, uint32_t n_queries, int coefs,
                           double tolerance, int freq_ignore_low, int freq_ignore_high, tir_hit *d_hits, const TirP2PArgs *p2p);
// tir_p2p.cu: publish finished local hits (used when the shard is empty and no match kernel runs)
int tir_p2p_publish_launch(tir_ctx *ctx, const tir_hit *d_hits, uint32_t n_queries, const TirP2PArgs &a);
// ... and fold the ranks' candidates into a.final_out in a launch of its own (the match kernels do it in their last CTA)
int tir_p2p_merge_launch(tir_ctx *ctx, const TirP2PArgs &a, uint32_t n_queries);

// tir_match.cu: pre-size the scratch of a search (no allocation / free inside the calls afterwards) and
// rebuild a dirty index now rather than inside the next match
int tir_search_reserve(tir_ctx *ctx, uint32_t n_queries, uint64_t F, uint64_t n_samples);
int tir_db_ensure_index(tir_ctx *ctx);        // caller holds ctx->mu
int tir_db_ensure_index_public(tir_ctx *ctx); // takes it

// tir_batcher.cpp: the batching dispatcher, over any batched search function (a context's tir_search, a group's
// tir_group_search)
TirBatcher *tir_batcher_create(int device, uint32_t max_batch, uint32_t max_wait_us,
                               std::function<int(const int16_t *, const uint64_t *, uint32_t, int, double, int, int, tir_hit *)> search,
                               std::function<std::string()> last_error);
int tir_batcher_submit(TirBatcher *b, const int16_t *pcm, uint64_t n_samples, int coefs, double tolerance, int freq_ignore_low,
                       int freq_ignore_high, tir_hit *hit, std::string *err);
bool tir_batcher_fits(TirBatcher *b, uint64_t n_samples);
void tir_batcher_enter(TirBatcher *b);
void tir_batcher_leave(TirBatcher *b);
void tir_batcher_counters(TirBatcher *b, uint64_t *n_requests, uint64_t *n_batches, uint64_t *max_batch_seen);
void tir_batcher_destroy(TirBatcher *b);
// tir_stream.cu
void tir_stream_hub_destroy(TirStreamHub *h);
