// Dependent-chain latency of scalar and packed FP32 ops and MUFU on sm_100a (one warp, clock64 around a chain).
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
template <int MODE> __global__ void lat(long long* out, float* sink, float seed, u64 seed2) {
  float a = seed; u64 p = seed2; const u64 c2 = seed2 ^ 0x0000100000001000ull;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 256; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) a = __fmaf_rn(a, seed, seed);
      if (MODE == 1) p = fma2(p, c2, c2);
      if (MODE == 2) a = __fadd_rn(a, seed);
      if (MODE == 3) p = add2(p, c2);
      if (MODE == 4) asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(a) : "f"(a));
      if (MODE == 5) { asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(a) : "f"(a)); a = __fmaf_rn(a, seed, seed); }
      if (MODE == 6) { p = fma2(p, c2, c2); a = __fmaf_rn(__uint_as_float((unsigned)p), seed, seed); p = (p & 0xffffffff00000000ull) | __float_as_uint(a); }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[MODE] = t1 - t0;
  sink[threadIdx.x] = a + __uint_as_float((unsigned)p);
}
int main() {
  long long* d; float* s; cudaMalloc(&d, 64); cudaMalloc(&s, 1024);
  lat<0><<<1, 32>>>(d, s, 1.0001f, 0x3f8000013f800001ull); lat<1><<<1, 32>>>(d, s, 1.0001f, 0x3f8000013f800001ull);
  lat<2><<<1, 32>>>(d, s, 1.0001f, 0x3f8000013f800001ull); lat<3><<<1, 32>>>(d, s, 1.0001f, 0x3f8000013f800001ull);
  lat<4><<<1, 32>>>(d, s, 1.0001f, 0x3f8000013f800001ull); lat<5><<<1, 32>>>(d, s, 1.0001f, 0x3f8000013f800001ull);
  lat<6><<<1, 32>>>(d, s, 1.0001f, 0x3f8000013f800001ull);
  long long h[8]; cudaMemcpy(h, d, 56, cudaMemcpyDeviceToHost);
  const char* n[] = { "FFMA", "FFMA2", "FADD", "FADD2", "MUFU.RSQ", "MUFU.RSQ+FFMA", "FFMA2+FFMA(lo half)+pack" };
  for (int i = 0; i < 7; ++i) printf("%-28s %.2f cycles per link\n", n[i], h[i] / 4096.0);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
