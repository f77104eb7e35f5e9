// tc_mel_tail.cu -- EXPERIMENT (not part of the product): the tail of the extraction path
//   |X[k]| -> 40 Slaney mel bands -> log10 -> DCT rows 0,1 -> 10*log10|c| -> "%f" micro-units
// (aubio_mfcc_do + the quantiser, src/fp_handler.c:642-652) written the way BASELINE.json's north_star words it:
// the mel filterbank as a tensor-core GEMM [frames x 272 bins] . [272 x 48] on tcgen05 (fp16 hi/lo split operands,
// three products, fp32 accumulation in TMEM), the magnitude tiles staged by TMA (cp.async.bulk.tensor.2d), and
// log10 / DCT / quantiser fused into the TMEM epilogue.  It reads the magnitudes from HBM, which the fused product
// kernel never writes: the point is to MEASURE what the tensor-core formulation costs per frame and how close its
// results stay to a float64 evaluation, next to the SIMT phases P3a+P3b+P4 of tir_extract_kernel.
//
// Build (tools/ubench/build_tc_mel_tail.sh):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -I asterisk_tiresias_b200/csrc \
//        -o tools/ubench/tc_mel_tail.bin tools/ubench/tc_mel_tail.cu asterisk_tiresias_b200/csrc/tir_tables.cpp
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tir_tables.h"

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);      \
      return 1;                                                                            \
    }                                                                                      \
  } while (0)

constexpr int TM = 128;                 // frames per tile (= MMA M = TMEM lanes)
constexpr int KPAD = 272;               // bins per frame in HBM (257 padded to a multiple of 16; row pitch 1088 B)
constexpr int KC = 64;                  // bins per chunk (one TMA box: 128 rows x 64 floats = 32 KB)
constexpr int NCHUNK = 5;               // 5 x 64 = 320 >= 272: the last box is zero-filled beyond column 271 by the TMA unit
constexpr int KTOT = KC * NCHUNK;
constexpr int NF = 48;                  // filters padded (MMA N)
constexpr int NSTAGE = 3;
constexpr int NT = 256;
constexpr uint32_t STAGE_BYTES = TM * KC * 4;
constexpr uint32_t A_LBO = TM * 16 + 16; // +16: the 8-byte stores of one row land in 32 distinct banks
constexpr uint32_t A_SLICE = (KC / 8) * A_LBO;
constexpr uint32_t B_LBO = NF * 16;
constexpr uint32_t B_SLICE = (KTOT / 8) * B_LBO;
constexpr int MAG_SHIFT = 8;            // magnitudes are stored * 2^-8 (fp16 range), the weights * 2^+8

struct SmemLayout {
  static constexpr uint32_t stage = 0;
  static constexpr uint32_t a3 = stage + NSTAGE * STAGE_BYTES;   // [buf 2][slice 2]
  static constexpr uint32_t b3 = a3 + 4 * A_SLICE;               // [slice 2]
  static constexpr uint32_t part = b3 + 2 * B_SLICE;             // [128][2] float
  static constexpr uint32_t bars = part + TM * 2 * 4;            // full[3], a3_free[2], d3_full
  static constexpr uint32_t tmem = bars + 8 * 8;
  static constexpr uint32_t total = tmem + 16;
};

__constant__ float c_dct[2][40];

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(b),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int x, int y, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
      "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
      : "memory");
}
// K-major, no swizzle: element (r, k) at (k / 8) * lbo + (r / 8) * sbo + (r % 8) * 16 + (k % 8) * 2  (tc_probe.cu: verified)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); // D f32, A/B fp16, both K-major
}
#define LD8(taddr, v, o)                                                                                      \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                       \
               : "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), \
                 "=r"(v[o + 6]), "=r"(v[o + 7])                                                                \
               : "r"(taddr))

// 10 * log10(|c|) in double and the exact "%f" rounding: the product's own code (tir_fp.cuh)
__device__ __forceinline__ void emit(float c, float *coef, int32_t *vq) {
  *coef = c;
  *vq = tir_quantize_micro(tir_coef_to_y(c));
}

__global__ void __launch_bounds__(NT, 1)
    tc_mel_tail_kernel(const __grid_constant__ CUtensorMap map, const unsigned char *__restrict__ g_b3, float *__restrict__ coef,
                       int32_t *__restrict__ vq, uint32_t n_frames, uint32_t n_tiles, int n_live, float lg_dead, float clamp) {
  extern __shared__ __align__(1024) unsigned char sm[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t sbase = smem_u32(sm);
  const uint32_t bar_full = sbase + SmemLayout::bars, bar_afree = bar_full + 8 * NSTAGE, bar_d3 = bar_afree + 16;
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(sm + SmemLayout::tmem);
  float *s_part = reinterpret_cast<float *>(sm + SmemLayout::part);

  // ---- set-up: weights -> shared memory, barriers, TMEM
  for (uint32_t i = tid; i < 2 * B_SLICE / 16; i += NT) reinterpret_cast<uint4 *>(sm + SmemLayout::b3)[i] = reinterpret_cast<const uint4 *>(g_b3)[i];
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; s++) mbar_init(bar_full + 8 * s, 1);
    mbar_init(bar_afree, 1), mbar_init(bar_afree + 8, 1), mbar_init(bar_d3, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tm = *s_tmem;
  const uint32_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const uint32_t total_chunks = my_tiles * NCHUNK;
  constexpr uint32_t idesc = make_idesc(TM, NF);

  auto issue_tma = [&](uint32_t g) { // chunk g of this CTA -> stage g % NSTAGE
    const uint32_t tile = blockIdx.x + (g / NCHUNK) * gridDim.x, c = g % NCHUNK, s = g % NSTAGE;
    mbar_expect_tx(bar_full + 8 * s, STAGE_BYTES);
    tma_load_2d(sbase + SmemLayout::stage + s * STAGE_BYTES, &map, (int)(c * KC), (int)(tile * TM), bar_full + 8 * s);
  };
  if (tid == 0)
    for (uint32_t g = 0; g < NSTAGE && g < total_chunks; g++) issue_tma(g);

  const int u = tid & 15, row0 = tid >> 4; // 16-byte unit of a 64-float row; rows row0, row0 + 16, ...
  uint32_t g = 0;
  for (uint32_t it = 0; it < my_tiles; it++) {
    const uint32_t tile = blockIdx.x + it * gridDim.x;
    for (int c = 0; c < NCHUNK; c++, g++) {
      const uint32_t s = g % NSTAGE, buf = g & 1;
      mbar_wait(bar_full + 8 * s, (g / NSTAGE) & 1);
      if (g >= 2) mbar_wait(bar_afree + 8 * buf, ((g >> 1) - 1) & 1); // the MMAs of chunk g - 2 have read this A buffer
      // ---- fp32 -> fp16 hi + fp16 lo, K-major core-matrix layout
      const float4 *src = reinterpret_cast<const float4 *>(sm + SmemLayout::stage + s * STAGE_BYTES);
      unsigned char *a_hi = sm + SmemLayout::a3 + (buf * 2 + 0) * A_SLICE, *a_lo = sm + SmemLayout::a3 + (buf * 2 + 1) * A_SLICE;
      const uint32_t doff = (uint32_t)(u >> 1) * A_LBO + (uint32_t)(u & 1) * 8;
      const bool live_unit = c < NCHUNK - 1 || u < (KPAD - (NCHUNK - 1) * KC) / 4; // the last chunk holds 16 real columns
#pragma unroll
      for (int j = 0; j < TM / 16; j++) {
        if (!live_unit) break;
        const int row = row0 + 16 * j;
        const float4 v = src[row * (KC / 4) + u];
        const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
        uint2 oh, ol;
        oh.x = *reinterpret_cast<const uint32_t *>(&h01), oh.y = *reinterpret_cast<const uint32_t *>(&h23);
        ol.x = *reinterpret_cast<const uint32_t *>(&l01), ol.y = *reinterpret_cast<const uint32_t *>(&l23);
        *reinterpret_cast<uint2 *>(a_hi + doff + row * 16) = oh;
        *reinterpret_cast<uint2 *>(a_lo + doff + row * 16) = ol;
      }
      fence_async_smem(); // the MMA (async proxy) reads what these generic stores wrote; the stage is free for the next TMA
      fence_before();
      __syncthreads();
      if (tid == 0) {
        fence_after();
        const int nks = (c == NCHUNK - 1) ? (KPAD - (NCHUNK - 1) * KC) / 16 : KC / 16; // the last chunk holds 16 real columns
        const uint32_t ah = sbase + SmemLayout::a3 + (buf * 2 + 0) * A_SLICE, al = ah + A_SLICE;
        const uint32_t bh = sbase + SmemLayout::b3, bl = bh + B_SLICE;
        for (int ks = 0; ks < nks; ks++) {
          const uint32_t ao = ks * 2 * A_LBO, bo = (c * (KC / 16) + ks) * 2 * B_LBO;
          const uint64_t dah = make_desc(ah + ao, A_LBO, 128), dal = make_desc(al + ao, A_LBO, 128);
          const uint64_t dbh = make_desc(bh + bo, B_LBO, 128), dbl = make_desc(bl + bo, B_LBO, 128);
          mma_ss(tm, dah, dbh, idesc, (c | ks) != 0);
          mma_ss(tm, dah, dbl, idesc, 1);
          mma_ss(tm, dal, dbh, idesc, 1);
        }
        mma_commit(bar_afree + 8 * buf);
        if (c == NCHUNK - 1) mma_commit(bar_d3);
        if (g + NSTAGE < total_chunks) issue_tma(g + NSTAGE);
      }
    }
    // ---- epilogue out of TMEM: log10, DCT rows 0 and 1, 10*log10|c|, "%f" micro-units
    mbar_wait(bar_d3, it & 1);
    fence_after();
    const int q = warp & 3, half = warp >> 2; // TMEM lane quadrant of the warp; filters [24 half, 24 half + 24)
    const int row = q * 32 + (tid & 31);
    uint32_t v[24];
    const uint32_t taddr = tm + ((uint32_t)(q * 32) << 16) + half * 24;
    LD8(taddr, v, 0);
    LD8(taddr + 8, v, 8);
    LD8(taddr + 16, v, 16);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    float p0 = 0.f, p1 = 0.f;
#pragma unroll
    for (int j = 0; j < 24; j++) {
      const int f = half * 24 + j;
      if (f < 40) {
        const float m = __uint_as_float(v[j]);
        const float lg = (f < n_live && m > clamp) ? __log2f(m) * 0.30102999566398120f : lg_dead;
        p0 = fmaf(c_dct[0][f], lg, p0), p1 = fmaf(c_dct[1][f], lg, p1);
      }
    }
    // the half that does not finish coefficient j hands its partial sum over: half 0 finishes c0, half 1 finishes c1
    s_part[row * 2 + half] = half ? p0 : p1;
    fence_before();
    __syncthreads();
    {
      const uint64_t frame = (uint64_t)tile * TM + row;
      const float mine = half ? p1 : p0, other = s_part[row * 2 + (half ^ 1)];
      if (frame < n_frames) emit(half ? other + mine : mine + other, coef + frame * 2 + half, vq + frame * 2 + half);
    }
    __syncthreads(); // s_part is rewritten by the next tile
  }
  fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(64u) : "memory");
}

// synthetic magnitudes (already * 2^-8): a spectral tilt, a frame level, pseudo-random fine structure; columns >= 257 are 0
__global__ void gen_mags_kernel(float *mags, uint64_t n_frames) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_frames * KPAD) return;
  const uint64_t f = i / KPAD;
  const uint32_t k = (uint32_t)(i % KPAD);
  uint64_t h = (f * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)k * 0xC2B2AE3D27D4EB4Full);
  h ^= h >> 29, h *= 0xBF58476D1CE4E5B9ull, h ^= h >> 32;
  const float r = (float)(h & 0xffffff) * (1.f / 16777216.f);
  uint64_t hf = f * 0xD6E8FEB86659FD93ull;
  hf ^= hf >> 32;
  const float level = exp2f((float)(hf & 1023) * (12.f / 1024.f) - 4.f); // 2^-4 .. 2^8
  const float tilt = 1.f / (1.f + 0.02f * (float)k);
  mags[i] = k < 257 ? level * tilt * (0.05f + 2.f * r * r) : 0.f;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
  const uint64_t n_frames = argc > 1 ? strtoull(argv[1], nullptr, 10) : (1ull << 22);
  const int reps = argc > 2 ? atoi(argv[2]) : 10;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s sm_%d%d, %d SMs; %llu frames x %d bins (%.2f GB of magnitudes per pass)\n", prop.name, prop.major, prop.minor,
         prop.multiProcessorCount, (unsigned long long)n_frames, KPAD, n_frames * KPAD * 4e-9);
  TirHostTables T;
  if (!tir_build_tables(512, 256, 40, 2, 8000, T)) return printf("tables failed\n"), 1;
  const int L = T.L;
  for (int f = 0; f < 40; f++)
    if (T.filters[(size_t)f * L] != 0.f) return printf("bin 0 has weight\n"), 1;
  // weights * 2^8 as fp16 hi + lo, K-major core-matrix layout [n][k]
  std::vector<unsigned char> hb(2 * B_SLICE, 0);
  for (int n = 0; n < 40; n++)
    for (int k = 0; k < L && k < KTOT; k++) {
      const float w = ldexpf(T.filters[(size_t)n * L + k], MAG_SHIFT);
      const __half hi = __float2half_rn(w), lo = __float2half_rn(w - __half2float(hi));
      const size_t off = (size_t)(k / 8) * B_LBO + (size_t)n * 16 + (size_t)(k % 8) * 2;
      memcpy(&hb[off], &hi, 2), memcpy(&hb[B_SLICE + off], &lo, 2);
    }
  int n_live = 0;
  for (int f = 0; f < 40; f++) n_live += !T.mel.dead[f];
  for (int f = 0; f < n_live; f++)
    if (T.mel.dead[f]) return printf("live filters are not a prefix\n"), 1;
  CK(cudaMemcpyToSymbol(c_dct, T.mel.dct, sizeof(float) * 2 * 40));
  unsigned char *d_b3;
  float *d_mags, *d_coef;
  int32_t *d_vq;
  CK(cudaMalloc(&d_b3, hb.size()));
  CK(cudaMemcpy(d_b3, hb.data(), hb.size(), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_mags, n_frames * KPAD * 4));
  CK(cudaMalloc(&d_coef, n_frames * 2 * 4));
  CK(cudaMalloc(&d_vq, n_frames * 2 * 4));
  gen_mags_kernel<<<(unsigned)((n_frames * KPAD + 255) / 256), 256>>>(d_mags, n_frames);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());

  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) return printf("cuTensorMapEncodeTiled not found\n"), 1;
  CUtensorMap map;
  const cuuint64_t dims[2] = {(cuuint64_t)KPAD, (cuuint64_t)n_frames}, strides[1] = {(cuuint64_t)KPAD * 4};
  const cuuint32_t box[2] = {KC, TM}, estr[2] = {1, 1};
  const CUresult cr = ((EncodeTiledFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_mags, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return printf("cuTensorMapEncodeTiled failed: %d\n", (int)cr), 1;

  const uint32_t n_tiles = (uint32_t)((n_frames + TM - 1) / TM);
  const uint32_t grid = n_tiles < (uint32_t)prop.multiProcessorCount ? n_tiles : (uint32_t)prop.multiProcessorCount;
  CK(cudaFuncSetAttribute(tc_mel_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemLayout::total));
  printf("grid %u x %d threads, %u B shared memory, %u tiles of %d frames\n", grid, NT, SmemLayout::total, n_tiles, TM);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int w = 0; w < 3; w++) {
    tc_mel_tail_kernel<<<grid, NT, SmemLayout::total>>>(map, d_b3, d_coef, d_vq, (uint32_t)n_frames, n_tiles, n_live, T.mel.lg_dead, 1e-30f);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
  }
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; r++)
    tc_mel_tail_kernel<<<grid, NT, SmemLayout::total>>>(map, d_b3, d_coef, d_vq, (uint32_t)n_frames, n_tiles, n_live, T.mel.lg_dead, 1e-30f);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= reps;
  const double bytes = (double)n_frames * (KPAD * 4 + 16);
  printf("tensor-core mel tail: %.4f ms per pass, %.3f G frames/s, %.1f GB/s of HBM traffic (%.0f B/frame), %.3f ns/frame\n", ms,
         n_frames / (ms * 1e6), bytes / (ms * 1e6), bytes / n_frames, ms * 1e6 / n_frames);

  // ---- float64 evaluation of the same magnitudes for the first frames
  const size_t nchk = n_frames < 8192 ? n_frames : 8192;
  std::vector<float> hm(nchk * KPAD), hc(nchk * 2);
  std::vector<int32_t> hv(nchk * 2);
  CK(cudaMemcpy(hm.data(), d_mags, hm.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hc.data(), d_coef, hc.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hv.data(), d_vq, hv.size() * 4, cudaMemcpyDeviceToHost));
  double max_rel = 0, sum_rel = 0;
  size_t same_hash = 0, same_trunc = 0, n_val = 0, flt_same_hash = 0;
  for (size_t f = 0; f < nchk; f++) {
    double lg[40];
    float lgf[40];
    for (int n = 0; n < 40; n++) {
      double s = 0;
      float sf = 0.f; // the reference order in float32 (fmat_vecmul), for the hash rate a float32 SIMT evaluation would reach
      for (int k = 0; k < L; k++) {
        s += (double)T.filters[(size_t)n * L + k] * (double)ldexpf(hm[f * KPAD + k], MAG_SHIFT);
        sf += T.filters[(size_t)n * L + k] * ldexpf(hm[f * KPAD + k], MAG_SHIFT);
      }
      const double vsn = (double)2e-42f; // aubio's VERY_SMALL_NUMBER is a float (a denormal: 1.99965e-42)
      lg[n] = log10(s > vsn ? s : vsn);
      lgf[n] = log10f(sf > 2e-42f ? sf : 2e-42f);
    }
    for (int j = 0; j < 2; j++) {
      double c = 0;
      float cf = 0.f;
      for (int n = 0; n < 40; n++) c += (double)T.mel.dct[j][n] * lg[n], cf += lgf[n] * T.mel.dct[j][n];
      const double y = 10.0 * log10(fabs(c)), yg = 10.0 * log10(fabs((double)hc[f * 2 + j])), yf = 10.0 * log10(fabs((double)cf));
      const double rel = fabs((double)hc[f * 2 + j] - c) / fabs(c);
      if (rel > max_rel) max_rel = rel;
      sum_rel += rel, n_val++;
      same_hash += llrint(y * 1e6) == (long long)hv[f * 2 + j];
      flt_same_hash += llrint(y * 1e6) == llrint(yf * 1e6);
      if (j == 0) same_trunc += (long long)y == (long long)yg;
    }
  }
  printf("against float64 on %zu frames: MFCC relative error max %.3e mean %.3e; trunc(max1) identical %.5f; micro-hash identical %.4f"
         " (a float32 evaluation in the reference's order: %.4f)\n",
         nchk, max_rel, sum_rel / n_val, (double)same_trunc / nchk, (double)same_hash / n_val, (double)flt_same_hash / n_val);
  return 0;
}
