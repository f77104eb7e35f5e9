// tc_probe.cu -- feasibility probe for a tensor-core DFT stage on sm_100a (tcgen05 / TMEM), standalone.
//   1. correctness of tcgen05.mma kind::f16 (fp16 in, fp32 accumulate in TMEM) with both operands in
//      shared memory in the K-major no-swizzle canonical layout (8 x 16-byte core matrices), for both
//      readings of the descriptor's LBO / SBO fields, and with A taken from TMEM (written with tcgen05.st);
//   2. accumulation error of the fp32 accumulator against float64;
//   3. throughput: back-to-back MMAs of several N (SS and TS), tcgen05.ld and tcgen05.st rates.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe.bin tc_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);      \
      return 1;                                                                            \
    }                                                                                      \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
      "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
      "r"(a), "l"(bd), "r"(idesc), "r"(acc)
      : "memory");
}

// K-major, no swizzle: element (r, k) of an [R x K] fp16 operand lives at
//   (r / 8) * sbo + (k / 8) * lbo + (r % 8) * 16 + (k % 8) * 2   bytes
__host__ __device__ inline uint32_t op_off(int r, int k, uint32_t lbo, uint32_t sbo) {
  return (uint32_t)(r / 8) * sbo + (uint32_t)(k / 8) * lbo + (uint32_t)(r % 8) * 16u + (uint32_t)(k % 8) * 2u;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46; // descriptor version (Blackwell)
  return d;               // base offset 0, layout type 0 = no swizzle
}
__host__ __device__ inline uint32_t make_idesc(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                    // D format f32
  d |= 0u << 7;                    // A fp16
  d |= 0u << 10;                   // B fp16
  d |= 0u << 15;                   // A K-major
  d |= 0u << 16;                   // B K-major
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

#define LD32(taddr, v, o)                                                                                               \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"  \
               : "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]),        \
                 "=r"(v[o + 6]), "=r"(v[o + 7]), "=r"(v[o + 8]), "=r"(v[o + 9]), "=r"(v[o + 10]), "=r"(v[o + 11]),     \
                 "=r"(v[o + 12]), "=r"(v[o + 13]), "=r"(v[o + 14]), "=r"(v[o + 15])                                     \
               : "r"(taddr))
#define ST8(taddr, v, o)                                                                                     \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),        \
               "r"(v[o + 0]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]),    \
               "r"(v[o + 6]), "r"(v[o + 7])                                                                  \
               : "memory")
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- 1. correctness ---------------------------------------------------------------------------
// A [128 x K], B [N x K] (both K-major), D [128 x N]; mode 0: SS, LBO = K-chunk stride; mode 1: SS with the two
// fields swapped; mode 2: TS (A in TMEM columns N.., two fp16 per column).
template <int N, int K>
__global__ void __launch_bounds__(128) k_mma_check(const __half *gA, const __half *gB, float *gD, int mode) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  // layout: rows contiguous across core matrices (sbo = 128), K chunks far apart (lbo = R * 16)
  const uint32_t a_lbo = 128 * 16, a_sbo = 128, b_lbo = N * 16, b_sbo = 128;
  unsigned char *sA = sm, *sB = sm + 128 * K * 2;
  for (int i = tid; i < 128 * K; i += 128) {
    const int r = i / K, k = i % K;
    *reinterpret_cast<__half *>(sA + op_off(r, k, a_lbo, a_sbo)) = gA[i];
  }
  for (int i = tid; i < N * K; i += 128) {
    const int r = i / K, k = i % K;
    *reinterpret_cast<__half *>(sB + op_off(r, k, b_lbo, b_sbo)) = gB[i];
  }
  if (warp == 0) tmem_alloc(&s_tmem, 256);
  if (tid == 0) mbar_init(&s_bar, 1);
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tm = s_tmem;
  const uint32_t lane_base = tm + ((uint32_t)(warp * 32) << 16);
  if (mode == 2) { // A -> TMEM columns [128, 128 + K/2): thread = row, column c holds (A[r][2c], A[r][2c+1])
    uint32_t v[K / 2];
    for (int c = 0; c < K / 2; c++) {
      const __half2 h = __halves2half2(gA[tid * K + 2 * c], gA[tid * K + 2 * c + 1]);
      v[c] = *reinterpret_cast<const uint32_t *>(&h);
    }
#pragma unroll
    for (int c = 0; c < K / 2; c += 8) ST8(lane_base + 128 + c, v, c);
    wait_st();
    fence_before();
    __syncthreads();
    fence_after();
  }
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, N);
    for (int ks = 0; ks < K / 16; ks++) {
      const uint32_t aoff = ks * 2 * a_lbo, boff = ks * 2 * b_lbo;
      uint64_t ad, bd;
      if (mode == 1) ad = make_desc(smem_u32(sA) + aoff, a_sbo, a_lbo), bd = make_desc(smem_u32(sB) + boff, b_sbo, b_lbo);
      else ad = make_desc(smem_u32(sA) + aoff, a_lbo, a_sbo), bd = make_desc(smem_u32(sB) + boff, b_lbo, b_sbo);
      if (mode == 2) mma_ts(tm, tm + 128 + ks * 8, bd, idesc, ks > 0);
      else mma_ss(tm, ad, bd, idesc, ks > 0);
    }
    mma_commit(&s_bar);
  }
  mbar_wait(&s_bar, 0);
  fence_after();
  uint32_t v[N];
#pragma unroll
  for (int c = 0; c < N; c += 16) LD32(lane_base + c, v, c);
  wait_ld();
  for (int c = 0; c < N; c++) gD[tid * N + c] = __uint_as_float(v[c]);
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 256);
}

// ---- 3a. MMA issue throughput: R back-to-back MMAs into the same accumulator --------------------
template <int N>
__global__ void __launch_bounds__(128) k_mma_rate(long long *out, int reps, int ts, int ndst) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 * 16 * 2 + 256 * 16 * 2) / 4; i += 128) reinterpret_cast<uint32_t *>(sm)[i] = 0;
  if (warp == 0) tmem_alloc(&s_tmem, 512);
  if (tid == 0) mbar_init(&s_bar, 1);
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tm = s_tmem;
  if (warp == 1) {
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    if (leader) {
      const uint32_t idesc = make_idesc(128, N);
      const uint64_t ad = make_desc(smem_u32(sm), 128 * 16, 128), bd = make_desc(smem_u32(sm) + 4096, N * 16, 128);
      const long long t0 = clock64();
#pragma unroll 1
      for (int r = 0; r < reps; r += 8) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const uint32_t d = tm + (uint32_t)(((u & (ndst - 1)) * N) & 255); // rotate accumulators (ndst = 1: one chain)
          if (ts) mma_ts(d, tm + 384, bd, idesc, 1);
          else mma_ss(d, ad, bd, idesc, 1);
        }
      }
      mma_commit(&s_bar);
      const long long t1 = clock64();
      mbar_wait(&s_bar, 0);
      const long long t2 = clock64();
      out[0] = t1 - t0, out[1] = t2 - t0;
    }
    __syncwarp();
  }
  mbar_wait(&s_bar, 0);
  fence_after();
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// ---- 3b. tcgen05.ld / st throughput ---------------------------------------------------------------
__global__ void __launch_bounds__(256) k_tmem_rate(long long *out, int reps, int nwarps, int do_st) {
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&s_tmem, 512);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tm = s_tmem;
  const uint32_t lane_base = tm + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
  uint32_t v[64];
  for (int i = 0; i < 64; i++) v[i] = tid + i;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps) {
    if (do_st) {
      for (int r = 0; r < reps; r++) {
#pragma unroll
        for (int c = 0; c < 64; c += 8) ST8(lane_base + ((r & 3) * 64) + c, v, c);
      }
      wait_st();
    } else {
      for (int r = 0; r < reps; r++) {
#pragma unroll
        for (int c = 0; c < 64; c += 16) LD32(lane_base + ((r & 3) * 64) + c, v, c);
        wait_ld();
        acc += v[0] ^ v[17] ^ v[34] ^ v[63];
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) out[0] = t1 - t0;
  if (acc == 0x12345678u) out[1] = acc;
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <int N, int K>
static int check(int mode, const char *name) {
  std::vector<__half> hA(128 * K), hB(N * K);
  std::vector<double> dA(128 * K), dB(N * K);
  srand(1234 + N + K);
  for (size_t i = 0; i < hA.size(); i++) { float x = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2half(x); dA[i] = __half2float(hA[i]); }
  for (size_t i = 0; i < hB.size(); i++) { float x = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2half(x); dB[i] = __half2float(hB[i]); }
  __half *gA, *gB;
  float *gD;
  CK(cudaMalloc(&gA, hA.size() * 2)); CK(cudaMalloc(&gB, hB.size() * 2)); CK(cudaMalloc(&gD, 128 * N * 4));
  CK(cudaMemcpy(gA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(gB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(gD, 0xff, 128 * N * 4));
  const size_t smem = 128 * K * 2 + N * K * 2 + 1024;
  CK(cudaFuncSetAttribute(k_mma_check<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_mma_check<N, K><<<1, 128, smem>>>(gA, gB, gD, mode);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> hD(128 * N);
  CK(cudaMemcpy(hD.data(), gD, hD.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxabs = 0;
  int bad = 0;
  for (int m = 0; m < 128; m++)
    for (int n = 0; n < N; n++) {
      double s = 0;
      for (int k = 0; k < K; k++) s += dA[m * K + k] * dB[n * K + k];
      const double e = fabs(s - (double)hD[m * N + n]);
      if (!(e <= 1e-3)) bad++;
      if (e > maxerr || e != e) maxerr = e;
      if (fabs(s) > maxabs) maxabs = fabs(s);
    }
  printf("check %-28s N=%3d K=%3d : %s  bad=%d max_abs_err=%.3e (max |D| %.2f)  D[0][0..3]=%g %g %g %g\n", name, N, K,
         bad ? "MISMATCH" : "ok", bad, maxerr, maxabs, hD[0], hD[1], hD[2], hD[3]);
  cudaFree(gA); cudaFree(gB); cudaFree(gD);
  return 0;
}

template <int N>
static int rate(int ts, int ndst) {
  long long *d, h[2];
  CK(cudaMalloc(&d, 16));
  const size_t smem = 4096 + 256 * 32 + 1024;
  CK(cudaFuncSetAttribute(k_mma_rate<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int reps = 4000;
  for (int it = 0; it < 2; it++) {
    k_mma_rate<N><<<1, 128, smem>>>(d, reps, ts, ndst);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
  }
  CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
  printf("mma rate %s M=128 N=%3d K=16 ndst=%d : issue %.1f cyc/mma, complete %.1f cyc/mma  (%.0f MAC/cyc)\n", ts ? "TS" : "SS", N,
         ndst, (double)h[0] / reps, (double)h[1] / reps, 128.0 * N * 16 * reps / (double)h[1]);
  cudaFree(d);
  return 0;
}

static int tmem_rate(int nwarps, int do_st) {
  long long *d, h[2];
  CK(cudaMalloc(&d, 16));
  const int reps = 2000;
  for (int it = 0; it < 2; it++) {
    k_tmem_rate<<<1, 256>>>(d, reps, nwarps, do_st);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
  }
  CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
  const double bytes = (double)nwarps * 32 * 64 * 4 * reps;
  printf("tmem %s, %d warps, 64 columns per op : %.1f cycles per warp-op, %.1f B/cycle/SM\n", do_st ? "st" : "ld", nwarps,
         (double)h[0] / reps, bytes / (double)h[0]);
  cudaFree(d);
  return 0;
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s sm_%d%d, %d SMs\n", p.name, p.major, p.minor, p.multiProcessorCount);
  check<64, 64>(0, "SS lbo=Kchunk sbo=8rows");
  check<16, 16>(0, "SS lbo=Kchunk sbo=8rows");
  check<16, 64>(0, "SS lbo=Kchunk sbo=8rows");
  check<64, 64>(2, "TS (A in TMEM)");
  check<16, 16>(2, "TS (A in TMEM)");
  check<128, 128>(0, "SS lbo=Kchunk sbo=8rows");
  rate<16>(0, 1); rate<16>(0, 4); rate<32>(0, 4); rate<64>(0, 1); rate<64>(0, 4); rate<128>(0, 2); rate<256>(0, 1);
  rate<16>(1, 4); rate<64>(1, 4); rate<128>(1, 2);
  tmem_rate(1, 0); tmem_rate(4, 0); tmem_rate(8, 0);
  tmem_rate(1, 1); tmem_rate(4, 1); tmem_rate(8, 1);
  return 0;
}
