// tir_extract.cu -- the fused extraction kernel (sm_100a) and its launcher.
//
// Replaces the hop loop of create_audio_fingerprints(), src/fp_handler.c:632-661, for a whole
// batch of clips in one launch.  See tir_extract_core.cuh for the phases; this file adds the
// tile bookkeeping, the PCM staging (P0) and the persistent-CTA loop.
//
// HBM traffic per frame: hop*2 bytes of PCM16 in (each sample read once; the 50 % overlap lives
// in shared memory) + 16 bytes out (2 x f32 coefficient, 2 x i32 micro-units) = 528 B at hop 256.
#include <cuda_runtime.h>

#include <cstring>

#include "tir_internal.h"
#include "tir_p2p_dev.cuh"

// one tile = T consecutive frames of one clip
struct __align__(16) TirTile {
  uint64_t c0;    // first sample of the clip in the PCM buffer
  int64_t nsamp;  // samples in the clip
  uint64_t out0;  // output frame index of the tile's first frame
  int32_t f0;     // first frame of the tile within the clip
  int32_t nvalid; // frames of the tile that exist (1..T)
};

struct TirExtractArgs {
  const void *pcm;        // int16_t samples, or float samples (F32IN kernels: the down-mixed multi-channel input * 2^15)
  const TirTile *tiles; // [n_tiles]
  const float4 *win4, *twp4, *twu4;
  float *coef;
  int32_t *vq;
  uint32_t n_tiles;
  uint32_t pcm_aligned8; // base pointer is aligned to one 4-sample unit (8 bytes of s16, 16 bytes of float)
  float2 neg_zero;       // (-0, -0): see tir_pmulx
  uint32_t *tile_counter; // zeroed before the launch: tiles beyond the first two of every CTA are claimed dynamically
  TirCoefX cx;            // sharded search: coefficients also go to every rank's buffer (cx.peer == nullptr: off)
};

__device__ const double2 k_logf_tab[16] = TIR_LOGF_TAB_INIT;
// Build with -DTIR_TRACE to get clock64 timestamps of the phases of one tile per warp, and the loop
// cycles of every CTA, printed by the launcher (how the imbalances described in DESIGN.md 2.3 were
// found; __syncthreads does not block at issue, so a phase's time includes the wait for the barrier
// before it).  Not compiled into the product.
#ifdef TIR_TRACE
__device__ long long g_trace[TIR_MAX_WARPS][8];
__device__ long long g_trace_cta[2048][3];
#define TIR_TRACE_DECL() long long tr_[6] = {0, 0, 0, 0, 0, 0}; int tr_iter = 0; const long long tr_loop0 = clock64()
#define TIR_TRACE_MARK(i) tr_[i] = clock64()
#define TIR_TRACE_TILE()                                                                     \
  do {                                                                                       \
    if (blockIdx.x == 7 && lane == 0 && tr_iter == 100)                                      \
      for (int k_ = 0; k_ < 6; k_++) g_trace[warp][k_] = tr_[k_];                            \
    tr_iter++;                                                                               \
  } while (0)
#define TIR_TRACE_END()                                                                      \
  do {                                                                                       \
    if (tid == 0 && blockIdx.x < 2048) {                                                     \
      uint32_t smid_;                                                                        \
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid_));                                     \
      g_trace_cta[blockIdx.x][0] = clock64() - tr_loop0, g_trace_cta[blockIdx.x][1] = tr_iter, g_trace_cta[blockIdx.x][2] = smid_; \
    }                                                                                        \
  } while (0)
#else
#define TIR_TRACE_DECL() (void)0
#define TIR_TRACE_MARK(i) (void)0
#define TIR_TRACE_TILE() (void)0
#define TIR_TRACE_END() (void)0
#endif

// tile descriptors from the per-clip prefix arrays (one thread per tile, binary search for its clip)
template <int T, int HOP>
__global__ void tir_build_tiles_kernel(const uint64_t *__restrict__ clip_off, const uint64_t *__restrict__ frame_off,
                                       const uint32_t *__restrict__ tile_off, uint32_t n_clips, uint32_t n_tiles,
                                       TirTile *__restrict__ tiles) {
  const uint32_t tile = blockIdx.x * blockDim.x + threadIdx.x;
  if (tile >= n_tiles) return;
  uint32_t lo = 0, hi = n_clips; // largest c with tile_off[c] <= tile
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (tile_off[mid] <= tile) lo = mid; else hi = mid;
  }
  TirTile td;
  td.c0 = clip_off[lo];
  td.nsamp = (int64_t)(clip_off[lo + 1] - clip_off[lo]);
  td.f0 = (int32_t)((tile - tile_off[lo]) * T);
  const int64_t nframes = (td.nsamp + HOP - 1) / HOP;
  td.nvalid = (int32_t)min((int64_t)T, nframes - td.f0);
  td.out0 = frame_off[lo] + (uint64_t)td.f0;
  tiles[tile] = td;
}

__device__ __forceinline__ TirTile tir_load_tile_desc(const TirTile *p) {
  const uint4 a = __ldg(reinterpret_cast<const uint4 *>(p));
  const uint4 b = __ldg(reinterpret_cast<const uint4 *>(p) + 1);
  TirTile t;
  t.c0 = (uint64_t)a.x | ((uint64_t)a.y << 32);
  t.nsamp = (int64_t)((uint64_t)a.z | ((uint64_t)a.w << 32));
  t.out0 = (uint64_t)b.x | ((uint64_t)b.y << 32);
  t.f0 = (int32_t)b.z, t.nvalid = (int32_t)b.w;
  return t;
}

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_unit(uint2 *dst, const int16_t *src) { cp_async8(dst, src); }
__device__ __forceinline__ void cp_async_unit(float4 *dst, const float *src) { cp_async16(dst, src); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// four samples s .. s+3 of a clip, zeros outside it, as one unit of the tile
__device__ __forceinline__ uint2 tir_edge_unit(const int16_t *clip, int64_t s, int64_t nsamp, uint2) {
  uint32_t h[4];
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const int64_t i = s + e;
    h[e] = (i >= 0 && i < nsamp) ? (uint32_t)(uint16_t)__ldg(clip + i) : 0u;
  }
  return make_uint2(h[0] | (h[1] << 16), h[2] | (h[3] << 16));
}
__device__ __forceinline__ float4 tir_edge_unit(const float *clip, int64_t s, int64_t nsamp, float4) {
  float h[4];
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const int64_t i = s + e;
    h[e] = (i >= 0 && i < nsamp) ? __ldg(clip + i) : 0.f;
  }
  return make_float4(h[0], h[1], h[2], h[3]);
}

// P0: (T+1) hops of the clip -> buf, one unit (4 samples) per copy, asynchronously where the four
// samples lie inside the clip and the unit is aligned; samples outside the clip are zeros (the first
// hop of a clip sees the all-zero pvoc history, the last hop is zero padded: new_aubio_pvoc /
// aubio_source_do).  Hop chunk c of the tile lands at buf[c * PCH ...]: lane = frame reads of P1
// then have a stride of PCH units (odd), conflict free.
// Thread tid copies unit p = tid % UPC of chunks tid / UPC, + NT / UPC, ...
template <int WIN, class U, class S>
__device__ __forceinline__ void tir_issue_tile_load(U *buf, const S *__restrict__ pcm, const TirTile &td,
                                                    bool base_aligned, int tid) {
  using C = TirCfg<WIN>;
  constexpr int UPC = C::HOP / 4;       // units per hop chunk
  constexpr int CSTEP = C::NT / UPC;    // chunks advanced per iteration (4)
  static_assert(C::NT % UPC == 0, "thread <-> unit mapping");
  const S *clip = pcm + td.c0;
  const bool aligned = base_aligned && ((td.c0 & 3) == 0);
  const int64_t nsamp = td.nsamp;
  const int p = tid % UPC;
  int chunk = tid / UPC;
  int64_t s = ((int64_t)td.f0 - 1 + chunk) * C::HOP + 4 * p;
  U *dst = buf + chunk * C::PCH + p;
  const bool interior = aligned && td.f0 >= 1 && ((int64_t)td.f0 + C::T) * C::HOP <= nsamp; // CTA-uniform
  if (interior) {
#pragma unroll
    for (int i = 0; i < (C::T + 1 + CSTEP - 1) / CSTEP; i++) {
      if (chunk + i * CSTEP <= C::T) cp_async_unit(dst + i * CSTEP * C::PCH, clip + s + (int64_t)i * CSTEP * C::HOP);
    }
  } else {
#pragma unroll 1
    for (; chunk <= C::T; chunk += CSTEP, s += (int64_t)CSTEP * C::HOP, dst += CSTEP * C::PCH) {
      if (aligned && s >= 0 && s + 4 <= nsamp) cp_async_unit(dst, clip + s);
      else *dst = tir_edge_unit(clip, s, nsamp, U{});
    }
  }
  cp_async_commit();
}

template <int WIN, class SM, class U>
__device__ __forceinline__ void tir_pass1(SM &sm, const U *pcm, int role, int lane, TirP2 nz) {
  if constexpr (WIN == 512) tir_pass1_512(sm, pcm, role, lane, nz);
  else tir_pass1_1024(sm, pcm, role, lane, nz);
}

// coefficient warps: DCT row `warp` of the tile described by `td`, from its finished log-mel values
__device__ __forceinline__ void tir_emit_coefs(const float *lg, const TirMelParams &mp, const TirExtractArgs &a,
                                               const TirTile &td, int warp, int lane) {
  if (lane < td.nvalid) {
    float c;
    int32_t v;
    tir_dct_phase(lg, mp, warp, lane, c, v);
    const uint64_t o = (td.out0 + (uint64_t)lane) * (uint64_t)mp.n_coefs + (uint64_t)warp;
    if (a.coef) a.coef[o] = c;
    if (a.vq) a.vq[o] = v;
    if (a.cx.peer) { // NVLink peer stores: 4 bytes per coefficient and rank, issued as P4 produces them
      const unsigned long long g = (a.cx.frame_base + td.out0 + (uint64_t)lane) * (uint64_t)mp.n_coefs + (uint64_t)warp;
      for (int p = 0; p < a.cx.world; p++) reinterpret_cast<float *>(a.cx.peer[p] + a.cx.off)[g] = c;
    }
  }
}

// Per tile: P1 | sync | P2 load | sync | P2 compute | sync | P3a sweep (coefficient warps: P4 of the
// previous tile first) + wait for the next tile's PCM | sync | P3b logs -- no barrier before the next P1.
// F32IN: float samples in (the down-mixed multi-channel path, tir_extract_interleaved): twice the PCM tile, one CTA per SM
template <int WIN, bool F32IN>
__global__ void __launch_bounds__(TirCfg<WIN>::NT, F32IN ? 1 : TirCfg<WIN>::CTAS_PER_SM)
    tir_extract_kernel(const __grid_constant__ TirExtractArgs a, const __grid_constant__ TirMelParams mp) {
  using C = TirCfg<WIN>;
  using SM = TirSmem<WIN, F32IN>;
  using S = typename std::conditional<F32IN, float, int16_t>::type;
  const S *const a_pcm = static_cast<const S *>(a.pcm);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SM &sm = *reinterpret_cast<SM *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0); // warp-uniform for the compiler
  const bool base_aligned = a.pcm_aligned8 != 0;
  const TirP2 nz = tir_pmk(a.neg_zero.x, a.neg_zero.y);

  // Tiles are claimed dynamically (one atomic per tile, issued two tiles ahead of its use): the two
  // CTAs of an SM do not run at the same speed -- the warp scheduler favours one of them (measured:
  // 2.19 M vs 2.84 M cycles for the same 203 tiles) -- and with a static split the slow one finishes
  // the last quarter of its tiles alone on the SM.  Every CTA starts with tiles bid and bid + grid.
  __shared__ uint32_t s_claim;
  TirTile cur = tir_load_tile_desc(a.tiles + blockIdx.x); // grid <= n_tiles
  tir_issue_tile_load<WIN>(sm.pcm, a_pcm, cur, base_aligned, tid);
  bool have_nxt = blockIdx.x + gridDim.x < a.n_tiles;
  TirTile nxt = cur, prev = cur;
  if (have_nxt) nxt = tir_load_tile_desc(a.tiles + blockIdx.x + gridDim.x);

  for (int i = tid; i < 16 * C::NW; i += C::NT) sm.win4[i] = a.win4[i], sm.twp4[i] = a.twp4[i];
  for (int i = tid; i < C::NW * 8; i += C::NT) sm.twu4[i] = a.twu4[i];
  if (tid < 16) sm.logtab[tid] = k_logf_tab[tid];
  for (int i = tid; i < TIR_MAX_W2; i += C::NT) sm.w2[i] = mp.w2[i];
  for (int i = tid; i < TIR_MAX_RUNS; i += C::NT) sm.run_bins[i] = mp.run_bins[i], sm.run_emit[i] = mp.run_emit[i];
  // bins 0 and M of the magnitude buffer are never written; keep the whole buffer finite
  for (int i = tid; i < SM::XCH_WORDS; i += C::NT) sm.xch[i] = 0.f;
  // filters without weights: log10f(clamp) once, nothing overwrites it
  for (int i = tid; i < 2 * TIR_MAX_FILTERS * 32; i += C::NT) {
    const int fidx = (i / 32) % TIR_MAX_FILTERS;
    if (fidx < mp.n_filters && mp.dead[fidx]) (&sm.lg[0][0])[i] = mp.lg_dead;
  }
  cp_async_wait_all();
  __syncthreads();

  int b = 0;
  bool first = true;
  TIR_TRACE_DECL();
  for (;;) {
    uint32_t claim = 0;
    if (have_nxt && tid == 0) claim = atomicAdd(a.tile_counter, 1u) + 2u * gridDim.x; // the tile after nxt; lands under P1
    TIR_TRACE_MARK(0);
    tir_pass1<WIN>(sm, sm.pcm, warp, lane, nz);
    TIR_TRACE_MARK(1);
    if (tid == 0) s_claim = claim;
    __syncthreads(); // P1 is done with the PCM buffer
    TirTile nn = nxt;
    bool have_nn = false;
    if (have_nxt) {
      tir_issue_tile_load<WIN>(sm.pcm, a_pcm, nxt, base_aligned, tid); // streams in under P2..P3a
      const uint32_t t2 = s_claim;
      have_nn = t2 < a.n_tiles;
      if (have_nn) nn = tir_load_tile_desc(a.tiles + t2); // needed one tile from now
    }
    TirPass2Regs rg;
    tir_pass2_load<WIN>(sm, warp, lane, rg);
    TIR_TRACE_MARK(2);
    __syncthreads(); // the magnitudes overwrite the exchange buffer
    if (warp == 0) tir_pass2_compute<WIN, true>(sm, warp, lane, rg, nz);
    else tir_pass2_compute<WIN, false>(sm, warp, lane, rg, nz);
    __syncthreads();
    TIR_TRACE_MARK(3);
    if (warp < mp.n_coefs && !first) tir_emit_coefs(sm.lg[b ^ 1], mp, a, prev, warp, lane);
    tir_mel_sweep(sm.xch, sm.lg[b], mp, sm.w2, sm.run_bins, sm.run_emit, warp, lane, nz);
    TIR_TRACE_MARK(4);
    cp_async_wait_all();
    __syncthreads(); // raw mel sums complete; the next tile's PCM has landed
    if (mp.live_prefix) tir_log_phase_prefix<C::NW>(sm.lg[b], sm.logtab, mp.log_clamp, mp.n_live, warp, lane);
    else tir_log_phase<C::NW>(sm.lg[b], sm.logtab, mp, warp, lane);
    TIR_TRACE_MARK(5);
    TIR_TRACE_TILE();
    if (!have_nxt) break;
    prev = cur, cur = nxt, nxt = nn, have_nxt = have_nn, b ^= 1, first = false;
  }
  TIR_TRACE_END();
  __syncthreads();
  if (warp < mp.n_coefs) tir_emit_coefs(sm.lg[b], mp, a, cur, warp, lane);
  if (a.cx.peer) { // the last CTA to finish tells every rank that this rank's coefficients are in place
    __threadfence_system();
    __syncthreads();
    if (tid == 0 && atomicAdd(a.cx.done, 1u) == gridDim.x - 1) {
      __threadfence_system();
      for (int p = 0; p < a.cx.world; p++)
        tir_st_release_sys(reinterpret_cast<uint32_t *>(a.cx.peer[p]) + TIR_P2P_COEF_FLAG0 + a.cx.rank, a.cx.epoch);
      *a.cx.done = 0;
    }
  }
}

// G.711 mu-law -> PCM16 (what a channel's native ulaw frames decode to; same table as Asterisk's
// AST_MULAW).  16 samples per thread when the source is 16-byte aligned.
__device__ __forceinline__ int16_t tir_ulaw1(uint32_t b) {
  const uint32_t u = ~b & 0xffu;
  const int32_t mag = (int32_t)((((u & 0x0fu) << 3) + 0x84u) << ((u >> 4) & 7u));
  return (int16_t)((u & 0x80u) ? 0x84 - mag : mag - 0x84);
}

__global__ void tir_ulaw_decode_kernel(const uint8_t *__restrict__ in, int16_t *__restrict__ out, uint64_t n) {
  const uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  if (i0 >= n) return;
  if (i0 + 16 <= n && ((uintptr_t)(in + i0) & 15) == 0 && ((uintptr_t)(out + i0) & 15) == 0) {
    const uint4 v = *reinterpret_cast<const uint4 *>(in + i0);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[8];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      o[2 * k] = (uint32_t)(uint16_t)tir_ulaw1(w[k] & 0xff) | ((uint32_t)(uint16_t)tir_ulaw1((w[k] >> 8) & 0xff) << 16);
      o[2 * k + 1] = (uint32_t)(uint16_t)tir_ulaw1((w[k] >> 16) & 0xff) | ((uint32_t)(uint16_t)tir_ulaw1(w[k] >> 24) << 16);
    }
    reinterpret_cast<uint4 *>(out + i0)[0] = make_uint4(o[0], o[1], o[2], o[3]);
    reinterpret_cast<uint4 *>(out + i0)[1] = make_uint4(o[4], o[5], o[6], o[7]);
  } else {
    for (uint64_t i = i0; i < n && i < i0 + 16; i++) out[i] = tir_ulaw1(in[i]);
  }
}

int tir_ulaw_decode_launch(tir_ctx *ctx, const uint8_t *d_in, int16_t *d_out, uint64_t n) {
  if (n == 0) return TIR_OK;
  const uint64_t threads = (n + 15) / 16;
  tir_ulaw_decode_kernel<<<(uint32_t)((threads + 255) / 256), 256, 0, ctx->stream>>>(d_in, d_out, n);
  TIR_CUDA(ctx, cudaGetLastError());
  ctx->launches++;
  return TIR_OK;
}

// ---- self tests of the two float primitives the bit-exactness rests on ---------------------------
// (a) tir_sqrt_scaled64 against the IEEE square root, for EVERY float in [2^-149, 2^60) and 0;
// (b) the device build of tir_log10f_glibc over a strided range, returned for comparison with the
//     host build of the same source (which the CPU tests pin to libm's log10f).
__global__ void tir_selftest_sqrt_kernel(unsigned long long *__restrict__ bad) {
  const uint32_t last = 0x5d800000u; // 2^60
  unsigned long long n = 0;
  for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < last; u += (uint64_t)gridDim.x * blockDim.x) {
    const float x = __uint_as_float((uint32_t)u);
    const float a = tir_sqrt_scaled64(x), b = __fsqrt_rn(__fmul_rn(x, 18446744073709551616.0f));
    n += __float_as_uint(a) != __float_as_uint(b);
  }
  for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  if ((threadIdx.x & 31) == 0 && n) atomicAdd(bad, n);
}

__global__ void tir_selftest_log10f_kernel(uint32_t first, uint32_t step, uint32_t count, float *__restrict__ out) {
  __shared__ double2 tab[16];
  if (threadIdx.x < 16) tab[threadIdx.x] = k_logf_tab[threadIdx.x];
  __syncthreads();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = tir_log10f_glibc(__uint_as_float(first + i * step), tab);
}

int tir_selftest_launch(tir_ctx *ctx, uint64_t *sqrt_mismatches, uint32_t first, uint32_t step, uint32_t count, float *log10f_out) {
  if (sqrt_mismatches) {
    unsigned long long *d = nullptr;
    TIR_CUDA(ctx, cudaMalloc(&d, 8));
    TIR_CUDA(ctx, cudaMemsetAsync(d, 0, 8, ctx->stream));
    tir_selftest_sqrt_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(d);
    unsigned long long h = 0;
    TIR_CUDA(ctx, cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, ctx->stream));
    TIR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d);
    *sqrt_mismatches = h;
  }
  if (log10f_out && count) {
    float *d = nullptr;
    TIR_CUDA(ctx, cudaMalloc(&d, (size_t)count * 4));
    tir_selftest_log10f_kernel<<<(count + 255) / 256, 256, 0, ctx->stream>>>(first, step, count, d);
    TIR_CUDA(ctx, cudaMemcpyAsync(log10f_out, d, (size_t)count * 4, cudaMemcpyDeviceToHost, ctx->stream));
    TIR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d);
  }
  return TIR_OK;
}

size_t tir_extract_smem_bytes(int win) {
  return win == 512 ? sizeof(TirSmem<512>) : win == 1024 ? sizeof(TirSmem<1024>) : 0;
}

// aubio_source_*_do for a multi-channel file (source_wavread.c / source_sndfile.c): every channel's sample / 32768 in
// float, summed in channel order from 0.f, divided by the channel count.  The result is kept * 2^15 (exact), so that
// the kernel's window product rn(v * w * 2^-15) is aubio's rn(mono * w).
__global__ void tir_downmix_kernel(const int16_t *__restrict__ in, uint64_t n_frames, int channels, float *__restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_frames) return;
  const int16_t *p = in + i * (uint64_t)channels;
  float acc = 0.f;
  for (int c = 0; c < channels; c++) acc = __fadd_rn(acc, __fdiv_rn((float)p[c], 32768.f));
  out[i] = __fmul_rn(__fdiv_rn(acc, (float)channels), 32768.f);
}

int tir_downmix_launch(tir_ctx *ctx, const int16_t *d_in, uint64_t n_frames, int channels, float *d_out) {
  if (n_frames == 0) return TIR_OK;
  tir_downmix_kernel<<<(uint32_t)((n_frames + 255) / 256), 256, 0, ctx->stream>>>(d_in, n_frames, channels, d_out);
  TIR_CUDA(ctx, cudaGetLastError());
  ctx->launches++;
  return TIR_OK;
}

template <int WIN, bool F32IN>
static int tir_extract_launch_t(tir_ctx *ctx, const void *d_pcm, const uint64_t *clip_off, uint32_t n_clips,
                                float *d_coef, int32_t *d_vq, uint64_t *n_frames, const TirCoefX *cx) {
  using C = TirCfg<WIN>;
  // ---- host-side tile bookkeeping (metadata only) -> one pinned staging buffer -> device
  const size_t nc1 = (size_t)n_clips + 1;
  std::vector<uint64_t> frame_off(nc1);
  std::vector<uint32_t> tile_off(nc1);
  frame_off[0] = 0, tile_off[0] = 0;
  for (uint32_t c = 0; c < n_clips; c++) {
    if (clip_off[c + 1] < clip_off[c]) return tir_fail(ctx, TIR_ERR_ARG, "clip_off must be non-decreasing");
    const uint64_t nf = tir_n_frames(clip_off[c + 1] - clip_off[c], C::HOP);
    const uint64_t nt = (nf + C::T - 1) / C::T;
    frame_off[c + 1] = frame_off[c] + nf;
    if ((uint64_t)tile_off[c] + nt > 0xffffffffull) return tir_fail(ctx, TIR_ERR_ARG, "batch too large");
    tile_off[c + 1] = tile_off[c] + (uint32_t)nt;
  }
  const uint32_t n_tiles = tile_off[n_clips];
  if (n_frames) *n_frames = frame_off[n_clips];
  if (n_tiles == 0) return TIR_OK;

  const size_t off_clip = 0, off_frame = off_clip + nc1 * 8, off_tileoff = off_frame + nc1 * 8;
  const size_t meta_bytes = (off_tileoff + nc1 * 4 + 15) & ~(size_t)15;
  int rc;
  if ((rc = tir_reserve(ctx, ctx->d_clipmeta, meta_bytes))) return rc;
  if ((rc = tir_reserve(ctx, ctx->d_tilemeta, (size_t)n_tiles * sizeof(TirTile)))) return rc;
  void *hp;
  int slot;
  if ((rc = tir_stage_acquire(ctx, meta_bytes, &hp, &slot))) return rc;
  unsigned char *h = (unsigned char *)hp;
  std::memcpy(h + off_clip, clip_off, nc1 * 8);
  std::memcpy(h + off_frame, frame_off.data(), nc1 * 8);
  std::memcpy(h + off_tileoff, tile_off.data(), nc1 * 4);
  unsigned char *d = (unsigned char *)ctx->d_clipmeta.p;
  TIR_CUDA(ctx, cudaMemcpyAsync(d, h, meta_bytes, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = tir_stage_release(ctx, slot))) return rc;
  TirTile *d_tiles = (TirTile *)ctx->d_tilemeta.p;
  tir_build_tiles_kernel<C::T, C::HOP><<<(n_tiles + 255) / 256, 256, 0, ctx->stream>>>(
      (const uint64_t *)(d + off_clip), (const uint64_t *)(d + off_frame), (const uint32_t *)(d + off_tileoff), n_clips,
      n_tiles, d_tiles);
  TIR_CUDA(ctx, cudaGetLastError());
  ctx->launches++;

  TirExtractArgs a;
  a.pcm = d_pcm;
  a.tiles = d_tiles;
  a.win4 = ctx->d_win4, a.twp4 = ctx->d_twp4, a.twu4 = ctx->d_twu4;
  a.coef = d_coef, a.vq = d_vq;
  a.n_tiles = n_tiles;
  a.pcm_aligned8 = (((uintptr_t)d_pcm) & (F32IN ? 15 : 7)) == 0;
  a.neg_zero = make_float2(-0.0f, -0.0f);
  if ((rc = tir_reserve(ctx, ctx->d_counter, 256))) return rc;
  TIR_CUDA(ctx, cudaMemsetAsync(ctx->d_counter.p, 0, sizeof(uint32_t), ctx->stream));
  a.tile_counter = (uint32_t *)ctx->d_counter.p;
  a.cx = cx ? *cx : TirCoefX{nullptr, 0, 0, 1, 0, nullptr, 0};

  const size_t smem = sizeof(TirSmem<WIN, F32IN>);
  if (F32IN || !ctx->smem_attr_set) { // per context: the attribute belongs to the device the context is on
    TIR_CUDA(ctx, cudaFuncSetAttribute(tir_extract_kernel<WIN, F32IN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (!F32IN) ctx->smem_attr_set = true;
  }
  const uint32_t resident = (uint32_t)ctx->num_sms * (uint32_t)(F32IN ? 1 : C::CTAS_PER_SM);
  const uint32_t grid = n_tiles < resident ? n_tiles : resident;
  if (ctx->profiling) TIR_CUDA(ctx, cudaEventRecord(ctx->ev[0][0], ctx->stream));
  tir_extract_kernel<WIN, F32IN><<<grid, C::NT, smem, ctx->stream>>>(a, ctx->tab.mel);
  TIR_CUDA(ctx, cudaGetLastError());
  if (ctx->profiling) {
    TIR_CUDA(ctx, cudaEventRecord(ctx->ev[0][1], ctx->stream));
    ctx->ev_valid[0] = true;
  }
  ctx->launches++;
#ifdef TIR_TRACE
  {
    static long long h[TIR_MAX_WARPS][8], hc[2048][3];
    cudaStreamSynchronize(ctx->stream);
    cudaMemcpyFromSymbol(h, g_trace, sizeof h);
    cudaMemcpyFromSymbol(hc, g_trace_cta, sizeof hc);
    long long mn = 1ll << 60, mx = 0, sum = 0;
    for (uint32_t i = 0; i < grid && i < 2048; i++) mn = hc[i][0] < mn ? hc[i][0] : mn, mx = hc[i][0] > mx ? hc[i][0] : mx, sum += hc[i][0];
    fprintf(stderr, "trace: CTA loop cycles min %lld mean %lld max %lld (CTA 7: %lld tiles on SM %lld)\n", mn, sum / (long long)grid, mx, hc[7][1], hc[7][2]);
    for (int w = 0; w < C::NW; w++)
      fprintf(stderr, "trace warp %2d: P1 %6lld | bar+P2load %6lld | bar+P2compute %6lld | bar+P3a %6lld | bar+P3b %6lld\n", w,
              h[w][1] - h[w][0], h[w][2] - h[w][1], h[w][3] - h[w][2], h[w][4] - h[w][3], h[w][5] - h[w][4]);
  }
#endif
  return TIR_OK;
}

int tir_extract_launch(tir_ctx *ctx, const int16_t *d_pcm, uint64_t total_samples, const uint64_t *clip_off,
                       uint32_t n_clips, float *d_coef, int32_t *d_vq, uint64_t *n_frames, const TirCoefX *cx) {
  (void)total_samples;
  if (ctx->cfg.win == 512) return tir_extract_launch_t<512, false>(ctx, d_pcm, clip_off, n_clips, d_coef, d_vq, n_frames, cx);
  if (ctx->cfg.win == 1024) return tir_extract_launch_t<1024, false>(ctx, d_pcm, clip_off, n_clips, d_coef, d_vq, n_frames, cx);
  return tir_fail(ctx, TIR_ERR_ARG, "extraction kernels exist for win 512 / hop 256 and win 1024 / hop 512");
}

// the same from float samples (down-mixed multi-channel input * 2^15, tir_downmix_launch)
int tir_extract_launch_f32(tir_ctx *ctx, const float *d_pcm, const uint64_t *clip_off, uint32_t n_clips, float *d_coef,
                           int32_t *d_vq, uint64_t *n_frames) {
  if (ctx->cfg.win == 512) return tir_extract_launch_t<512, true>(ctx, d_pcm, clip_off, n_clips, d_coef, d_vq, n_frames, nullptr);
  if (ctx->cfg.win == 1024) return tir_extract_launch_t<1024, true>(ctx, d_pcm, clip_off, n_clips, d_coef, d_vq, n_frames, nullptr);
  return tir_fail(ctx, TIR_ERR_ARG, "extraction kernels exist for win 512 / hop 256 and win 1024 / hop 512");
}
