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

struct tir_p2p {
  tir_ctx *ctx = nullptr;
  int rank = 0, world = 1;
  uint32_t max_queries = 0, epoch = 0;
  unsigned char *local = nullptr;            // [flags: TIR_P2P_MAX_RANKS u32 | error u32 | pad][gather 0][gather 1]
  std::vector<unsigned char *> peer;         // base of every rank's region (peer[rank] == local)
  std::vector<bool> opened;                  // peer regions opened with cudaIpcOpenMemHandle
  unsigned char **d_peer = nullptr;          // the same table on the device
  uint32_t *d_done = nullptr;                // CTA counter of the publish kernel
  tir_hit *d_local_hits = nullptr;           // this rank's winners
};

static size_t p2p_gather_bytes(const tir_p2p *p) { return (size_t)p->world * p->max_queries * sizeof(tir_hit); }
static size_t p2p_region_bytes(const tir_p2p *p) { return TIR_P2P_HDR + 2 * p2p_gather_bytes(p); }

__device__ __forceinline__ void tir_st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t tir_ld_acquire_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

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

// wait for the flags of all ranks, then the greatest (match_count, uuid bytes) per query
__global__ void __launch_bounds__(256)
    tir_p2p_merge_kernel(unsigned char *local, int world, uint32_t max_queries, uint32_t n_queries, uint32_t epoch,
                         tir_hit *__restrict__ out) {
  __shared__ uint32_t s_ok;
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  uint32_t *flags = reinterpret_cast<uint32_t *>(local);
  if (threadIdx.x < (unsigned)world) {
    // batch numbers only grow; a bounded wait so that a rank that died cannot hang this GPU
    uint32_t spins = 0;
    while ((int32_t)(tir_ld_acquire_sys(flags + threadIdx.x) - epoch) < 0) {
      __nanosleep(64);
      if (++spins > (1u << 25)) { // ~ seconds
        s_ok = 0;
        flags[TIR_P2P_MAX_RANKS] = epoch; // error word: read by tir_p2p_match_dev's caller through tir_p2p_error
        break;
      }
    }
  }
  __syncthreads();
  if (!s_ok) return;
  // the peers' stores are read with ld.global.cg: never from this SM's L1 or the read-only path
  const unsigned long long *gathered = reinterpret_cast<const unsigned long long *>(local + TIR_P2P_HDR + (size_t)(epoch & 1u) * world * max_queries * sizeof(tir_hit));
  auto load_hit = [&](int s, uint32_t q) {
    const unsigned long long *w = gathered + ((size_t)s * max_queries + q) * 3;
    union { unsigned long long u[3]; tir_hit h; } v;
    v.u[0] = __ldcg(w), v.u[1] = __ldcg(w + 1), v.u[2] = __ldcg(w + 2);
    return v.h;
  };
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n_queries; q += gridDim.x * blockDim.x) {
    tir_hit bestv = load_hit(0, q);
    for (int s = 1; s < world; s++) {
      const tir_hit h = load_hit(s, q);
      bool better = h.match_count > bestv.match_count;
      if (h.match_count == bestv.match_count && h.match_count > 0) {
        int c = 0;
        for (int i = 0; i < 16 && c == 0; i++) c = (int)h.uuid[i] - (int)bestv.uuid[i];
        better = c > 0;
      }
      if (better) bestv = h;
    }
    out[q] = bestv;
  }
}

int tir_p2p_publish_launch(tir_ctx *ctx, const tir_hit *d_hits, uint32_t n_queries, const TirP2PArgs &a) {
  const uint32_t words = n_queries * 3;
  const uint32_t pgrid = std::min<uint32_t>((words + 255) / 256, 64);
  tir_p2p_publish_kernel<<<pgrid, 256, 0, ctx->stream>>>(d_hits, n_queries, a.peer, a.rank, a.world, a.max_queries, a.epoch, a.done);
  TIR_CUDA(ctx, cudaGetLastError());
  ctx->launches++;
  return TIR_OK;
}

extern "C" {

int tir_p2p_create(tir_ctx *ctx, int rank, int world, uint32_t max_queries, tir_p2p **out) {
  if (!ctx || !out) return TIR_ERR_ARG;
  *out = nullptr;
  if (world < 1 || world > TIR_P2P_MAX_RANKS || rank < 0 || rank >= world || max_queries == 0)
    return tir_fail(ctx, TIR_ERR_ARG, "tir_p2p_create: rank %d of %d, max_queries %u", rank, world, max_queries);
  std::lock_guard<std::mutex> lk(ctx->mu);
  TIR_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
  tir_p2p *p = new (std::nothrow) tir_p2p();
  if (!p) return tir_fail(ctx, TIR_ERR_NOMEM, "out of memory");
  p->ctx = ctx, p->rank = rank, p->world = world, p->max_queries = max_queries;
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

int tir_p2p_match_dev(tir_p2p *p, const float *d_coef, const uint64_t *frame_off, uint32_t n_queries, int coefs,
                      double tolerance, int freq_ignore_low, int freq_ignore_high, tir_hit *d_final) {
  if (!p || !d_final) return TIR_ERR_ARG;
  tir_ctx *ctx = p->ctx;
  if (n_queries > p->max_queries) return tir_fail(ctx, TIR_ERR_ARG, "tir_p2p_match_dev: %u queries, sized for %u", n_queries, p->max_queries);
  for (int r = 0; r < p->world; r++)
    if (!p->peer[r]) return tir_fail(ctx, TIR_ERR_STATE, "tir_p2p_match_dev: not connected");
  const uint32_t epoch = ++p->epoch; // (SPMD: every rank counts the same batches)
  if (n_queries == 0) return TIR_OK;
  const TirP2PArgs a{p->d_peer, p->rank, p->world, p->max_queries, epoch, p->d_done};
  int rc = tir_match_dev_exchange(ctx, d_coef, frame_off, n_queries, coefs, tolerance, freq_ignore_low, freq_ignore_high,
                                  p->d_local_hits, &a);
  if (rc != TIR_OK) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  TIR_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
  tir_p2p_merge_kernel<<<std::min<uint32_t>((n_queries + 255) / 256, 64), 256, 0, ctx->stream>>>(p->local, p->world, p->max_queries,
                                                                                                  n_queries, epoch, d_final);
  TIR_CUDA(ctx, cudaGetLastError());
  ctx->launches += 1;
  return TIR_OK;
}

/* 0, or the batch number at which a merge gave up waiting for a peer (after a stream synchronize) */
int tir_p2p_error(tir_p2p *p, uint32_t *epoch_out) {
  if (!p || !epoch_out) return TIR_ERR_ARG;
  std::lock_guard<std::mutex> lk(p->ctx->mu);
  TIR_CUDA(p->ctx, cudaSetDevice(p->ctx->cfg.device));
  TIR_CUDA(p->ctx, cudaMemcpyAsync(epoch_out, p->local + TIR_P2P_MAX_RANKS * 4, 4, cudaMemcpyDeviceToHost, p->ctx->stream));
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
