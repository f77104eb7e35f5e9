// tir_p2p_dev.cuh -- device side of the NVLink peer-memory exchange (shared by tir_match.cu, tir_extract.cu
// and tir_p2p.cu): release / acquire of the per-rank flags and the fold of the ranks' candidates.
#pragma once
#include "tir_internal.h"

__device__ __forceinline__ void tir_st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t tir_ld_acquire_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// All threads of a CTA: wait until flags[0..world) have all reached `epoch` (batch numbers only grow).
// Bounded, so that a rank that died cannot hang this GPU: on a timeout the error word of the region is
// set to the batch number and false is returned.
__device__ __forceinline__ bool tir_p2p_wait_flags(uint32_t *hdr, int flag0, int world, uint32_t epoch) {
  __shared__ uint32_t s_p2p_ok;
  if (threadIdx.x == 0) s_p2p_ok = 1;
  __syncthreads();
  if (threadIdx.x < (unsigned)world) {
    uint32_t spins = 0;
    while ((int32_t)(tir_ld_acquire_sys(hdr + flag0 + threadIdx.x) - epoch) < 0) {
      __nanosleep(64);
      if (++spins > (1u << 25)) { // ~ seconds
        s_p2p_ok = 0;
        hdr[TIR_P2P_ERR_WORD] = epoch; // read through tir_p2p_error
        break;
      }
    }
  }
  __syncthreads();
  return s_p2p_ok != 0;
}

// All threads of a CTA (or of a grid: `first`/`stride` in queries): the greatest (match_count, uuid bytes)
// per query over the ranks' rows of this batch's gather buffer.  The peers' stores are read with
// ld.global.cg: never from this SM's L1 or the read-only path.
__device__ __forceinline__ void tir_p2p_fold(const unsigned char *local, int world, uint32_t max_queries, uint32_t n_queries,
                                             uint32_t epoch, tir_hit *__restrict__ out, uint32_t first, uint32_t stride) {
  const unsigned long long *gathered =
      reinterpret_cast<const unsigned long long *>(local + TIR_P2P_HDR + (size_t)(epoch & 1u) * world * max_queries * sizeof(tir_hit));
  auto load_hit = [&](int s, uint32_t q) {
    const unsigned long long *w = gathered + ((size_t)s * max_queries + q) * 3;
    union { unsigned long long u[3]; tir_hit h; } v;
    v.u[0] = __ldcg(w), v.u[1] = __ldcg(w + 1), v.u[2] = __ldcg(w + 2);
    return v.h;
  };
  for (uint32_t q = first; q < n_queries; q += stride) {
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

// a merge that gave up waiting: every query reads match_count = -1 (no stale winners)
__device__ __forceinline__ void tir_p2p_poison(tir_hit *__restrict__ out, uint32_t n_queries, uint32_t first, uint32_t stride) {
  for (uint32_t q = first; q < n_queries; q += stride) {
    tir_hit h;
    for (int i = 0; i < 16; i++) h.uuid[i] = 0;
    h.match_count = -1, h.frame_count = 0;
    out[q] = h;
  }
}

// the fused tail of the kernels that produce the winners: (all threads of the LAST CTA, after
// tir_exchange_release) wait for every rank's flag, fold into x.final_out
__device__ __forceinline__ uint32_t tir_p2p_epoch(const TirP2PArgs &x) { return x.epoch_dev ? *x.epoch_dev : x.epoch; }
// (`first` / `stride` in queries: one CTA folds everything, or every CTA of a grid its share -- each CTA waits for
// the flags itself)
__device__ __forceinline__ void tir_p2p_fused_merge(const TirP2PArgs &x, uint32_t n_queries, uint32_t first, uint32_t stride) {
  if (!x.final_out) return;
  const uint32_t epoch = tir_p2p_epoch(x);
  if (tir_p2p_wait_flags(reinterpret_cast<uint32_t *>(x.local), 0, x.world, epoch))
    tir_p2p_fold(x.local, x.world, x.max_queries, n_queries, epoch, x.final_out, first, stride);
  else
    tir_p2p_poison(x.final_out, n_queries, first, stride);
}
__device__ __forceinline__ void tir_p2p_fused_merge(const TirP2PArgs &x, uint32_t n_queries) {
  tir_p2p_fused_merge(x, n_queries, threadIdx.x, blockDim.x);
}
