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

struct TirExtractArgs {
  const int16_t *pcm;
  const uint64_t *clip_off;  // [n_clips+1] samples
  const uint64_t *frame_off; // [n_clips+1] output frame index
  const uint32_t *tile_off;  // [n_clips+1] first tile of each clip
  const uint32_t *tile_clip; // [n_tiles]   owning clip of each tile
  const float2 *win2, *tw_pass, *tw_unt;
  float *coef;
  int32_t *vq;
  uint32_t n_tiles;
  uint32_t pcm_aligned16; // base pointer is 16-byte aligned
};

__device__ const double2 k_logf_tab[16] = TIR_LOGF_TAB_INIT;

__device__ __forceinline__ uint4 ldg_stream16(const void *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// P0: (T+1) hops of the clip -> sm.pcm ; samples outside the clip are zeros (the first hop of a
// clip sees the all-zero pvoc history, the last hop is zero padded: aubio_source_do / new_aubio_pvoc)
template <int WIN>
__device__ __forceinline__ void tir_load_tile(TirSmem<WIN> &sm, const int16_t *__restrict__ clip, int64_t nsamp,
                                              int64_t s_first, bool aligned, int tid) {
  using C = TirCfg<WIN>;
  constexpr int VPC = C::HOP / 8; // 16-byte vectors per hop chunk
  constexpr int TOTAL = (C::T + 1) * VPC;
  for (int v = tid; v < TOTAL; v += C::NT) {
    const int chunk = v / VPC, iv = v % VPC;
    const int64_t s = s_first + (int64_t)chunk * C::HOP + iv * 8;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (s >= 0 && s + 8 <= nsamp && aligned) {
      val = ldg_stream16(clip + s);
    } else if (s + 8 > 0 && s < nsamp) {
      uint32_t h[8];
#pragma unroll
      for (int e = 0; e < 8; e++) {
        const int64_t i = s + e;
        h[e] = (i >= 0 && i < nsamp) ? (uint32_t)(uint16_t)__ldg(clip + i) : 0u;
      }
      val = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
    }
    *reinterpret_cast<uint4 *>(&sm.pcm[chunk * C::PCM_STRIDE_W + iv * 4]) = val;
  }
}

template <int WIN>
__global__ void __launch_bounds__(TirCfg<WIN>::NT, 2)
    tir_extract_kernel(const __grid_constant__ TirExtractArgs a, const __grid_constant__ TirMelParams mp) {
  using C = TirCfg<WIN>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TirSmem<WIN> &sm = *reinterpret_cast<TirSmem<WIN> *>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < C::M; i += C::NT) sm.win2[i] = a.win2[i];
  for (int i = tid; i < C::N1 * 16; i += C::NT) sm.tw_pass[i] = a.tw_pass[i];
  for (int i = tid; i < 16 * C::TPF; i += C::NT) sm.tw_unt[i] = a.tw_unt[i];
  if (tid < 16) sm.logtab[tid] = k_logf_tab[tid];
  __syncthreads();

  for (uint32_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const uint32_t clip = __ldg(a.tile_clip + tile);
    const uint64_t c0 = __ldg(a.clip_off + clip), c1 = __ldg(a.clip_off + clip + 1);
    const int64_t nsamp = (int64_t)(c1 - c0);
    const int64_t f0 = (int64_t)(tile - __ldg(a.tile_off + clip)) * C::T;
    const int64_t nframes = (nsamp + C::HOP - 1) / C::HOP;
    const int nvalid = (int)min((int64_t)C::T, nframes - f0);
    const uint64_t out0 = __ldg(a.frame_off + clip) + (uint64_t)f0;
    const bool aligned = a.pcm_aligned16 && ((c0 & 7) == 0);

    tir_load_tile<WIN>(sm, a.pcm + c0, nsamp, (f0 - 1) * C::HOP, aligned, tid);
    __syncthreads();
    tir_pass1<WIN>(sm, tid);
    __syncthreads();
    TirPass2Regs rg;
    tir_pass2_load<WIN>(sm, tid, rg);
    __syncthreads(); // the magnitudes overwrite the exchange buffer
    tir_pass2_compute<WIN>(sm, tid, rg);
    __syncthreads();
    tir_mel_phase(sm.xch, sm.lg, sm.logtab, mp, warp, lane);
    __syncthreads();
    if (warp < mp.n_coefs && lane < nvalid) {
      float c;
      int32_t v;
      tir_dct_phase(sm.lg, mp, warp, lane, c, v);
      const uint64_t o = (out0 + (uint64_t)lane) * (uint64_t)mp.n_coefs + (uint64_t)warp;
      if (a.coef) a.coef[o] = c;
      if (a.vq) a.vq[o] = v;
    }
    // no barrier needed here: sm.lg is next written four barriers from now
  }
}

size_t tir_extract_smem_bytes(int win) { return win == 512 ? sizeof(TirSmem<512>) : 0; }

int tir_extract_launch(tir_ctx *ctx, const int16_t *d_pcm, uint64_t total_samples, const uint64_t *clip_off,
                       uint32_t n_clips, float *d_coef, int32_t *d_vq, uint64_t *n_frames) {
  (void)total_samples;
  if (ctx->cfg.win != 512) return tir_fail(ctx, TIR_ERR_ARG, "extraction kernel is built for win 512 / hop 256");
  using C = TirCfg<512>;
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
  const size_t off_tileclip = (off_tileoff + nc1 * 4 + 15) & ~(size_t)15;
  const size_t meta_bytes = off_tileclip + (size_t)n_tiles * 4;
  int rc;
  if ((rc = tir_reserve_host(ctx, ctx->h_meta, meta_bytes))) return rc;
  if ((rc = tir_reserve(ctx, ctx->d_tilemeta, meta_bytes))) return rc;
  unsigned char *h = (unsigned char *)ctx->h_meta.p;
  TIR_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // the staging buffer may still be in flight
  std::memcpy(h + off_clip, clip_off, nc1 * 8);
  std::memcpy(h + off_frame, frame_off.data(), nc1 * 8);
  std::memcpy(h + off_tileoff, tile_off.data(), nc1 * 4);
  uint32_t *tc = (uint32_t *)(h + off_tileclip);
  for (uint32_t c = 0; c < n_clips; c++)
    for (uint32_t t = tile_off[c]; t < tile_off[c + 1]; t++) tc[t] = c;
  unsigned char *d = (unsigned char *)ctx->d_tilemeta.p;
  TIR_CUDA(ctx, cudaMemcpyAsync(d, h, meta_bytes, cudaMemcpyHostToDevice, ctx->stream));

  TirExtractArgs a;
  a.pcm = d_pcm;
  a.clip_off = (const uint64_t *)(d + off_clip);
  a.frame_off = (const uint64_t *)(d + off_frame);
  a.tile_off = (const uint32_t *)(d + off_tileoff);
  a.tile_clip = (const uint32_t *)(d + off_tileclip);
  a.win2 = ctx->d_win2, a.tw_pass = ctx->d_tw_pass, a.tw_unt = ctx->d_tw_unt;
  a.coef = d_coef, a.vq = d_vq;
  a.n_tiles = n_tiles;
  a.pcm_aligned16 = (((uintptr_t)d_pcm) & 15) == 0;

  const size_t smem = sizeof(TirSmem<512>);
  static bool attr_set = false;
  if (!attr_set) {
    TIR_CUDA(ctx, cudaFuncSetAttribute(tir_extract_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const uint32_t resident = (uint32_t)ctx->num_sms * 2u;
  const uint32_t grid = n_tiles < resident ? n_tiles : resident;
  tir_extract_kernel<512><<<grid, C::NT, smem, ctx->stream>>>(a, ctx->tab.mel);
  TIR_CUDA(ctx, cudaGetLastError());
  ctx->launches++;
  return TIR_OK;
}
