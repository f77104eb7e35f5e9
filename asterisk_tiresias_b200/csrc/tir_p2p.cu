// tir_p2p.cu -- the one cross-GPU step of the sharded match, over NVLink peer memory.
//
// Every rank holds the fingerprints whose uuid hashes to it (tir_shard_of) and computes its local
// winner per query; the global winner is the greatest (match_count, uuid bytes) over the ranks
// (src/fp_handler.c:354-377: ORDER BY count(*) DESC, ties as SQLite resolves them).  With NCCL that
// is an all_gather of Q x 24 B plus a merge kernel -- about 25 us of collective latency for 24 KB.
// Here the exchange is part of the match itself: after the local chain, tir_p2p_publish_kernel
// STORES this rank's winners straight into every peer's gather buffer (peer pointers opened with
// CUDA IPC; NVLink P2P writes), the last CTA releases a per-peer flag with the batch number, and
// tir_p2p_merge_kernel acquires the flags of all ranks and folds the N candidates per query.
// (That describes the exchange; in the match path the publishing half is FUSED into the kernels that
// produce the winners -- tir_pattern_resolve_kernel / tir_match_kernel store every hit into the peers'
// buffers as they compute it and their last CTA releases the flags, see TirP2PArgs in tir_internal.h --
// so a batch costs one extra launch, the merge.  tir_p2p_publish_kernel serves shards with no rows.)
// No collective call, no host synchronisation; 24 bytes per query and rank cross the links.
//
// One process per GPU (tir_p2p_create / _handle / _connect with handles exchanged by the launcher,
// e.g. torch.distributed.all_gather), or several contexts of one process (tir_p2p_connect_local).
// All ranks must call tir_p2p_match_dev the same number of times (SPMD), like any collective.
// Two gather buffers alternate by batch parity: a rank can only be one batch ahead of a peer (its
// merge of batch e waits for the peer's flag e, which the peer raises after it has merged e - 1).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

#include "tir_internal.h"
#include "tir_p2p_dev.cuh"

struct tir_p2p {
  tir_ctx *ctx = nullptr;
  int rank = 0, world = 1;
  uint32_t max_queries = 0, epoch = 0;
  uint64_t max_frames = 0;                   // > 0: the region also holds two coefficient buffers (sharded search)
  unsigned char *local = nullptr;            // [flags, TIR_P2P_HDR bytes][gather 0][gather 1][coef 0][coef 1]
  std::vector<unsigned char *> peer;         // base of every rank's region (peer[rank] == local)
  std::vector<bool> opened;                  // peer regions opened with cudaIpcOpenMemHandle
  unsigned char **d_peer = nullptr;          // the same table on the device
  uint32_t *d_done = nullptr;                // CTA counters: [0] winners (publish / resolve / match kernels), [1] extraction
  tir_hit *d_local_hits = nullptr;           // this rank's winners
};

static size_t p2p_gather_bytes(const tir_p2p *p) { return (size_t)p->world * p->max_queries * sizeof(tir_hit); }
static size_t p2p_coef_bytes(const tir_p2p *p) { return (size_t)p->max_frames * TIR_N_COEFS * sizeof(float); }
static size_t p2p_coef_off(const tir_p2p *p, uint32_t epoch) { return TIR_P2P_HDR + 2 * p2p_gather_bytes(p) + (size_t)(epoch & 1u) * p2p_coef_bytes(p); }
static size_t p2p_region_bytes(const tir_p2p *p) { return TIR_P2P_HDR + 2 * p2p_gather_bytes(p) + 2 * p2p_coef_bytes(p); }

// this rank's winners -> row `rank` of every rank's gather buffer; then flag[rank] = epoch everywhere
__global__ void __launch_bounds__(256)
    tir_p2p_publish_kernel(const tir_hit *__restrict__ hits, uint32_t n_queries, unsigned char *const *__restrict__ peer, int rank,
                           int world, uint32_t max_queries, uint32_t epoch, uint32_t *__restrict__ done) {
  const size_t gather = (size_t)TIR_P2P_HDR + (size_t)(epoch & 1u) * world * max_queries * sizeof(tir_hit);
  const uint32_t words = n_queries * (uint32_t)(sizeof(tir_hit) / 8); // 3 x 8 bytes per hit
  const unsigned long long *src = reinterpret_cast<const unsigned long long *>(hits);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < words; i += gridDim.x * blockDim.x) {
    const unsigned long long v = src[i];
    for (int p = 0; p < world; p++) {
      unsigned long long *dst = reinterpret_cast<unsigned long long *>(peer[p] + gather + (size_t)rank * max_queries * sizeof(tir_hit));
      dst[i] = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  __shared__ uint32_t s_last;
  if (threadIdx.x == 0) s_last = atomicAdd(done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if (threadIdx.x < (unsigned)world) tir_st_release_sys(reinterpret_cast<uint32_t *>(peer[threadIdx.x]) + rank, epoch);
  if (threadIdx.x == 0) *done = 0; // for the next batch (stream order)
}

// wait for the flags of all ranks, then the greatest (match_count, uuid bytes) per query.  (The match
// kernels fold in their own last CTA, tir_p2p_fused_merge; this launch serves ranks whose shard is empty.)
__global__ void __launch_bounds__(256)
    tir_p2p_merge_kernel(unsigned char *local, int world, uint32_t max_queries, uint32_t n_queries, uint32_t epoch,
                         tir_hit *__restrict__ out) {
  const uint32_t first = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  if (tir_p2p_wait_flags(reinterpret_cast<uint32_t *>(local), 0, world, epoch)) tir_p2p_fold(local, world, max_queries, n_queries, epoch, out, first, stride);
  else tir_p2p_poison(out, n_queries, first, stride);
}

// sharded search: wait until every rank's coefficients of this batch have landed in this rank's buffer
__global__ void tir_p2p_wait_coef_kernel(unsigned char *local, int world, uint32_t epoch) {
  (void)tir_p2p_wait_flags(reinterpret_cast<uint32_t *>(local), TIR_P2P_COEF_FLAG0, world, epoch);
}
// a rank with no query clip of its own in the batch still raises its coefficient flag everywhere
__global__ void tir_p2p_flag_coef_kernel(unsigned char *const *__restrict__ peer, int rank, int world, uint32_t epoch) {
  if (threadIdx.x < (unsigned)world) tir_st_release_sys(reinterpret_cast<uint32_t *>(peer[threadIdx.x]) + TIR_P2P_COEF_FLAG0 + rank, epoch);
}

int tir_p2p_publish_launch(tir_ctx *ctx, const tir_hit *d_hits, uint32_t n_queries, const TirP2PArgs &a) {
  const uint32_t words = n_queries * 3;
  const uint32_t pgrid = std::min<uint32_t>((words + 255) / 256, 64);
  tir_p2p_publish_kernel<<<pgrid, 256, 0, ctx->stream>>>(d_hits, n_queries, a.peer, a.rank, a.world, a.max_queries, a.epoch, a.done);
  TIR_CUDA(ctx, cudaGetLastError());
  ctx->launches++;
  return TIR_OK;
}

int tir_p2p_merge_launch(tir_ctx *ctx, const TirP2PArgs &a, uint32_t n_queries) {
  tir_p2p_merge_kernel<<<std::min<uint32_t>((n_queries + 255) / 256, 64), 256, 0, ctx->stream>>>(a.local, a.world, a.max_queries,
                                                                                                  n_queries, a.epoch, a.final_out);
  TIR_CUDA(ctx, cudaGetLastError());
  ctx->launches++;
  return TIR_OK;
}

extern "C" {

int tir_p2p_create(tir_ctx *ctx, int rank, int world, uint32_t max_queries, tir_p2p **out) {
  return tir_p2p_create2(ctx, rank, world, max_queries, 0, out);
}

int tir_p2p_create2(tir_ctx *ctx, int rank, int world, uint32_t max_queries, uint64_t max_frames, tir_p2p **out) {
  if (!ctx || !out) return TIR_ERR_ARG;
  *out = nullptr;
  if (world < 1 || world > TIR_P2P_MAX_RANKS || rank < 0 || rank >= world || max_queries == 0)
    return tir_fail(ctx, TIR_ERR_ARG, "tir_p2p_create: rank %d of %d, max_queries %u", rank, world, max_queries);
  std::lock_guard<std::mutex> lk(ctx->mu);
  TIR_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
  tir_p2p *p = new (std::nothrow) tir_p2p();
  if (!p) return tir_fail(ctx, TIR_ERR_NOMEM, "out of memory");
  p->ctx = ctx, p->rank = rank, p->world = world, p->max_queries = max_queries, p->max_frames = max_frames;
  p->peer.assign(world, nullptr), p->opened.assign(world, false);
  cudaError_t e = cudaMalloc(&p->local, p2p_region_bytes(p));
  if (e == cudaSuccess) e = cudaMemset(p->local, 0, p2p_region_bytes(p));
  if (e == cudaSuccess) e = cudaMalloc(&p->d_peer, sizeof(unsigned char *) * world);
  if (e == cudaSuccess) e = cudaMalloc(&p->d_done, 256);
  if (e == cudaSuccess) e = cudaMemset(p->d_done, 0, 256);
  if (e == cudaSuccess) e = cudaMalloc(&p->d_local_hits, (size_t)max_queries * sizeof(tir_hit));
  if (e != cudaSuccess) {
    tir_fail(ctx, TIR_ERR_NOMEM, "tir_p2p_create: %s", cudaGetErrorString(e));
    cudaFree(p->local), cudaFree(p->d_peer), cudaFree(p->d_done), cudaFree(p->d_local_hits);
    delete p;
    return TIR_ERR_NOMEM;
  }
  p->peer[rank] = p->local;
  *out = p;
  return TIR_OK;
}

int tir_p2p_handle(tir_p2p *p, unsigned char handle[64]) {
  if (!p || !handle) return TIR_ERR_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  std::lock_guard<std::mutex> lk(p->ctx->mu);
  TIR_CUDA(p->ctx, cudaSetDevice(p->ctx->cfg.device));
  cudaIpcMemHandle_t h;
  TIR_CUDA(p->ctx, cudaIpcGetMemHandle(&h, p->local));
  std::memcpy(handle, &h, 64);
  return TIR_OK;
}

static int p2p_upload_table(tir_p2p *p) {
  TIR_CUDA(p->ctx, cudaMemcpy(p->d_peer, p->peer.data(), sizeof(unsigned char *) * p->world, cudaMemcpyHostToDevice));
  return TIR_OK;
}

int tir_p2p_connect(tir_p2p *p, const unsigned char *handles) {
  if (!p || !handles) return TIR_ERR_ARG;
  std::lock_guard<std::mutex> lk(p->ctx->mu);
  TIR_CUDA(p->ctx, cudaSetDevice(p->ctx->cfg.device));
  for (int r = 0; r < p->world; r++) {
    if (r == p->rank) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handles + (size_t)r * 64, 64);
    void *ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return tir_fail(p->ctx, TIR_ERR_CUDA, "tir_p2p_connect: rank %d: %s", r, cudaGetErrorString(e));
    p->peer[r] = (unsigned char *)ptr, p->opened[r] = true;
  }
  return p2p_upload_table(p);
}

int tir_p2p_connect_local(tir_p2p *p, tir_p2p *const *all) {
  if (!p || !all) return TIR_ERR_ARG;
  std::lock_guard<std::mutex> lk(p->ctx->mu);
  TIR_CUDA(p->ctx, cudaSetDevice(p->ctx->cfg.device));
  for (int r = 0; r < p->world; r++) {
    if (!all[r] || all[r]->world != p->world || all[r]->rank != r || all[r]->max_queries != p->max_queries)
      return tir_fail(p->ctx, TIR_ERR_ARG, "tir_p2p_connect_local: entry %d does not belong to this group", r);
    p->peer[r] = all[r]->local;
    const int peer_dev = all[r]->ctx->cfg.device;
    if (peer_dev != p->ctx->cfg.device) {
      int can = 0;
      TIR_CUDA(p->ctx, cudaDeviceCanAccessPeer(&can, p->ctx->cfg.device, peer_dev));
      if (!can) return tir_fail(p->ctx, TIR_ERR_CUDA, "device %d cannot access device %d", p->ctx->cfg.device, peer_dev);
      cudaError_t e = cudaDeviceEnablePeerAccess(peer_dev, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) TIR_CUDA(p->ctx, e);
      (void)cudaGetLastError();
    }
  }
  return p2p_upload_table(p);
}

// local failure after the batch number was taken: the peers' merges of this batch would wait for this
// rank until their time-out.  Publish "no winner" rows and the flag so that they complete.
static void p2p_publish_nothing(tir_p2p *p, uint32_t n_queries, const TirP2PArgs &a) {
  tir_ctx *ctx = p->ctx;
  std::lock_guard<std::mutex> lk(ctx->mu);
  cudaSetDevice(ctx->cfg.device);
  (void)cudaGetLastError();
  if (cudaMemsetAsync(p->d_local_hits, 0, (size_t)n_queries * sizeof(tir_hit), ctx->stream) == cudaSuccess)
    (void)tir_p2p_publish_launch(ctx, p->d_local_hits, n_queries, a);
}

int tir_p2p_match_dev(tir_p2p *p, const float *d_coef, const uint64_t *frame_off, uint32_t n_queries, int coefs,
                      double tolerance, int freq_ignore_low, int freq_ignore_high, tir_hit *d_final) {
  if (!p || !d_final) return TIR_ERR_ARG;
  tir_ctx *ctx = p->ctx;
  if (n_queries > p->max_queries) return tir_fail(ctx, TIR_ERR_ARG, "tir_p2p_match_dev: %u queries, sized for %u", n_queries, p->max_queries);
  for (int r = 0; r < p->world; r++)
    if (!p->peer[r]) return tir_fail(ctx, TIR_ERR_STATE, "tir_p2p_match_dev: not connected");
  // (SPMD: every rank counts the same batches; a batch of zero queries is counted but exchanges nothing
  // on any rank)
  const uint32_t epoch = ++p->epoch;
  if (n_queries == 0) return TIR_OK;
  TirP2PArgs a{p->d_peer, p->rank, p->world, p->max_queries, epoch, p->d_done};
  a.local = p->local, a.final_out = d_final; // the kernel that produces the winners also folds the ranks' candidates
  a.may_alloc = true; // (SPMD entry point: this rank's host thread enqueues nothing for other ranks)
  const int rc = tir_match_dev_exchange(ctx, d_coef, frame_off, n_queries, coefs, tolerance, freq_ignore_low, freq_ignore_high,
                                        p->d_local_hits, &a);
  if (rc != TIR_OK) p2p_publish_nothing(p, n_queries, a);
  return rc;
}

int tir_p2p_reserve(tir_p2p *p, uint64_t max_local_samples) {
  if (!p) return TIR_ERR_ARG;
  tir_ctx *ctx = p->ctx;
  std::lock_guard<std::mutex> lk(ctx->mu);
  TIR_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
  int rc;
  if ((rc = tir_db_ensure_index(ctx))) return rc; // (first: the reserve sizes the coefs == 2 item list from the index's blocks)
  return tir_search_reserve(ctx, p->max_queries, p->max_frames, max_local_samples);
}

// SPMD sharded search (header: tir_p2p_search).  Rank r brings the clips [first_query, first_query + n_local)
// of a batch of n_total queries: H2D of its own clips only, extraction with the coefficients stored into
// every rank's buffer as they are produced, wait for all ranks' coefficient flags, match of ALL queries
// against the local shard with the winners exchanged and folded inside the match kernels.
int tir_p2p_search(tir_p2p *p, const int16_t *pcm, const uint64_t *clip_off, uint32_t n_local, uint32_t first_query,
                   const uint64_t *all_frame_off, uint32_t n_total, int coefs, double tolerance, int freq_ignore_low,
                   int freq_ignore_high, tir_hit *hits, tir_hit *d_final) {
  if (!p || !all_frame_off || (n_local && (!clip_off || !pcm))) return TIR_ERR_ARG;
  tir_ctx *ctx = p->ctx;
  if (coefs < 1 || coefs > TIR_N_COEFS) return tir_fail(ctx, TIR_ERR_ARG, "Wrong coefs count. max[%d], coefs[%d]", TIR_N_COEFS, coefs);
  if (n_total > p->max_queries || (uint64_t)first_query + n_local > n_total)
    return tir_fail(ctx, TIR_ERR_ARG, "tir_p2p_search: queries [%u, %u) of %u, sized for %u", first_query, first_query + n_local, n_total, p->max_queries);
  if (all_frame_off[0] != 0 || all_frame_off[n_total] > p->max_frames)
    return tir_fail(ctx, TIR_ERR_ARG, "tir_p2p_search: %llu frames, sized for %llu (tir_p2p_create2)", (unsigned long long)all_frame_off[n_total], (unsigned long long)p->max_frames);
  for (uint32_t c = 0; c < n_local; c++)
    if (clip_off[c + 1] < clip_off[c] ||
        tir_n_frames(clip_off[c + 1] - clip_off[c], ctx->cfg.hop) != all_frame_off[first_query + c + 1] - all_frame_off[first_query + c])
      return tir_fail(ctx, TIR_ERR_ARG, "tir_p2p_search: clip %u does not have the frames all_frame_off gives query %u", c, first_query + c);
  for (int r = 0; r < p->world; r++)
    if (!p->peer[r]) return tir_fail(ctx, TIR_ERR_STATE, "tir_p2p_search: not connected");
  const uint32_t epoch = ++p->epoch;
  if (n_total == 0) return TIR_OK;
  TirP2PArgs a{p->d_peer, p->rank, p->world, p->max_queries, epoch, p->d_done};
  tir_hit *d_out = d_final;
  int rc = TIR_OK;
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    rc = [&]() -> int {
      TIR_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
      int rc2;
      if (!d_out) {
        if ((rc2 = tir_reserve(ctx, ctx->d_hits, (size_t)n_total * sizeof(tir_hit)))) return rc2;
        d_out = (tir_hit *)ctx->d_hits.p;
      }
      const uint64_t base = n_local ? clip_off[0] : 0, total = n_local ? clip_off[n_local] - base : 0;
      const uint64_t F_local = all_frame_off[first_query + n_local] - all_frame_off[first_query];
      if (F_local) {
        std::vector<uint64_t> rel((size_t)n_local + 1);
        for (uint32_t c = 0; c <= n_local; c++) rel[c] = clip_off[c] - base;
        if ((rc2 = tir_reserve(ctx, ctx->d_pcm, total * sizeof(int16_t) + 16))) return rc2;
        TIR_CUDA(ctx, cudaMemcpyAsync(ctx->d_pcm.p, pcm + base, total * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
        const TirCoefX cx{p->d_peer, (unsigned long long)p2p_coef_off(p, epoch), p->rank, p->world, epoch, p->d_done + 1,
                          (unsigned long long)all_frame_off[first_query]};
        if ((rc2 = tir_extract_launch(ctx, (const int16_t *)ctx->d_pcm.p, total, rel.data(), n_local, nullptr, nullptr, nullptr, &cx))) return rc2;
      } else {
        tir_p2p_flag_coef_kernel<<<1, 32, 0, ctx->stream>>>(p->d_peer, p->rank, p->world, epoch);
        TIR_CUDA(ctx, cudaGetLastError());
        ctx->launches++;
      }
      tir_p2p_wait_coef_kernel<<<1, 32, 0, ctx->stream>>>(p->local, p->world, epoch);
      TIR_CUDA(ctx, cudaGetLastError());
      ctx->launches++;
      return TIR_OK;
    }();
  }
  a.local = p->local, a.final_out = d_out;
  if (rc == TIR_OK)
    rc = tir_match_dev_exchange(ctx, (const float *)(p->local + p2p_coef_off(p, epoch)), all_frame_off, n_total, coefs, tolerance,
                                freq_ignore_low, freq_ignore_high, p->d_local_hits, &a);
  if (rc != TIR_OK) {
    // (a rank that failed before its extraction never raises its coefficient flag: the peers' waits time
    // out and poison their results -- tir_p2p_error reports it; the winners' exchange is completed here)
    p2p_publish_nothing(p, n_total, a);
    return rc;
  }
  if (hits) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    TIR_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    TIR_CUDA(ctx, cudaMemcpyAsync(hits, d_out, (size_t)n_total * sizeof(tir_hit), cudaMemcpyDeviceToHost, ctx->stream));
    TIR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return TIR_OK;
}

/* 0, or the batch number at which a merge gave up waiting for a peer (after a stream synchronize) */
int tir_p2p_error(tir_p2p *p, uint32_t *epoch_out) {
  if (!p || !epoch_out) return TIR_ERR_ARG;
  std::lock_guard<std::mutex> lk(p->ctx->mu);
  TIR_CUDA(p->ctx, cudaSetDevice(p->ctx->cfg.device));
  TIR_CUDA(p->ctx, cudaMemcpyAsync(epoch_out, p->local + TIR_P2P_ERR_WORD * 4, 4, cudaMemcpyDeviceToHost, p->ctx->stream));
  TIR_CUDA(p->ctx, cudaStreamSynchronize(p->ctx->stream));
  return TIR_OK;
}

void tir_p2p_destroy(tir_p2p *p) {
  if (!p) return;
  cudaSetDevice(p->ctx->cfg.device);
  cudaStreamSynchronize(p->ctx->stream);
  for (int r = 0; r < p->world; r++)
    if (p->opened[r]) cudaIpcCloseMemHandle(p->peer[r]);
  cudaFree(p->local), cudaFree(p->d_peer), cudaFree(p->d_done), cudaFree(p->d_local_hits);
  delete p;
}

} // extern "C"
