// microbenchmark: issue cost of FFMA / FFMA2 / FADD2 mixed with ALU ops on sm_100a.
// 148*2 CTAs x 512 threads, 100 KB dynamic smem each => exactly 2 CTAs (32 warps, 8 per SMSP) per SM.
// Reports cycles per loop iteration per SMSP-warp-slot: (ms * clk_hz) / (iters * 8 warps).
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float ffma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float fadd1(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned xor1(unsigned a, unsigned b) { unsigned r; asm volatile("prmt.b32 %0, %1, %2, 0x2541;" : "=r"(r) : "r"(a), "r"(b)); return r; }

// NF1 scalar FFMA, NF2 FFMA2, NA2 FADD2, NX xor per iteration, all independent chains
template <int NF1, int NF2, int NA2, int NX>
__global__ void __launch_bounds__(512) k(float* out, int iters, float s) {
  extern __shared__ float sm[];
  float a[NF1 + 1]; u64 p[NF2 + 1], q[NA2 + 1]; unsigned x[NX + 1];
  for (int i = 0; i <= NF1; i++) a[i] = threadIdx.x * 0.001f + i;
  for (int i = 0; i <= NF2; i++) p[i] = ((u64)__float_as_uint(1.f + i) << 32) | __float_as_uint(threadIdx.x * 1.f);
  for (int i = 0; i <= NA2; i++) q[i] = ((u64)__float_as_uint(2.f + i) << 32) | __float_as_uint(threadIdx.x * 1.f);
  for (int i = 0; i <= NX; i++) x[i] = threadIdx.x + i;
  u64 ss = ((u64)__float_as_uint(s) << 32) | __float_as_uint(s);
  unsigned xs = __float_as_uint(s);
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < NF1; i++) a[i] = ffma1(a[i], s, s);
#pragma unroll
      for (int i = 0; i < NF2; i++) p[i] = ffma2(p[i], ss, ss);
#pragma unroll
      for (int i = 0; i < NA2; i++) q[i] = fadd2(q[i], ss);
#pragma unroll
      for (int i = 0; i < NX; i++) x[i] = xor1(x[i], xs);
    }
  }
  float r = 0;
  for (int i = 0; i < NF1; i++) r += a[i];
  for (int i = 0; i < NF2; i++) r += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
  for (int i = 0; i < NA2; i++) r += __uint_as_float((unsigned)q[i]) + __uint_as_float((unsigned)(q[i] >> 32));
  for (int i = 0; i < NX; i++) r += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r + sm[threadIdx.x];
}
template <int NF1, int NF2, int NA2, int NX> void run(float* d, double clk_hz) {
  int iters = 5000;
  auto kern = k<NF1, NF2, NA2, NX>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<148 * 2, 512, 100 * 1024>>>(d, 100, 1.0001f);
  cudaEventRecord(e0);
  kern<<<148 * 2, 512, 100 * 1024>>>(d, iters, 1.0001f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double cyc = ms * 1e-3 * clk_hz / (iters * 4.0 * 8.0);  // per unrolled group per warp on one SMSP
  printf("FFMA=%2d FFMA2=%2d FADD2=%2d XOR=%2d : instr=%2d  cycles/group/warp=%.2f\n", NF1, NF2, NA2, NX, NF1 + NF2 + NA2 + NX, cyc);
}
int main() {
  float* d; cudaMalloc(&d, 148 * 2 * 512 * 4);
  int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double hz = khz * 1e3; printf("clock %.0f MHz (nominal max; actual may differ)\n", hz / 1e6);
  run<16, 0, 0, 0>(d, hz);
  run<0, 8, 0, 0>(d, hz);
  run<0, 16, 0, 0>(d, hz);
  run<0, 0, 8, 0>(d, hz);
  run<0, 0, 0, 16>(d, hz);
  run<8, 0, 0, 8>(d, hz);
  run<16, 0, 0, 8>(d, hz);
  run<16, 0, 0, 16>(d, hz);
  run<0, 8, 0, 4>(d, hz);
  run<0, 8, 0, 8>(d, hz);
  run<0, 8, 0, 12>(d, hz);
  run<0, 8, 0, 16>(d, hz);
  run<8, 4, 0, 0>(d, hz);
  run<8, 4, 0, 8>(d, hz);
  run<0, 4, 4, 8>(d, hz);
  return 0;
}
