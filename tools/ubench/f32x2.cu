// Microbenchmark: issue/throughput of packed FP32x2 (FADD2/FMUL2/FFMA2) against scalar FP32 on sm_100a,
// and a bitwise check that the packed ops round exactly like the scalar .rn ops (denormals included).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu ; run on one GPU.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <random>
typedef unsigned long long u64;
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int MODE>
__global__ void __launch_bounds__(128) bench(float* out, int iters, float seed, u64 nz) {
  // 16 scalar accumulators or 8 packed pairs: same flop count per iteration
  float a[16]; u64 p[8];
  for (int i = 0; i < 16; ++i) a[i] = seed + i + threadIdx.x;
  for (int i = 0; i < 8; ++i) p[i] = ((u64)__float_as_uint(a[2 * i + 1]) << 32) | __float_as_uint(a[2 * i]);
  const float m = 1.0000001f, c = 1e-7f;
  const u64 m2 = ((u64)__float_as_uint(m) << 32) | __float_as_uint(m), c2 = ((u64)__float_as_uint(c) << 32) | __float_as_uint(c);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = __fmaf_rn(a[i], m, c);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], m2, c2);
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = __fadd_rn(a[i], c);
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = add2(p[i], c2);
    } else if (MODE == 4) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = __fmul_rn(a[i], m);
    } else if (MODE == 5) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = mul2(p[i], m2);
    } else if (MODE == 6) {   // mixed: packed fma + scalar alu work to see co-issue
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], m2, nz);
    } else if (MODE == 7) {   // MUFU.RSQ throughput, 16 chains
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(a[i]) : "f"(a[i]));
    }
  }
  float s = 0.f;
  for (int i = 0; i < 16; ++i) s += a[i];
  for (int i = 0; i < 8; ++i) s += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void check(const uint32_t* x, const uint32_t* y, const uint32_t* z, int n, unsigned long long* bad, u64 nz) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n / 2) return;
  float a0 = __uint_as_float(x[2 * i]), a1 = __uint_as_float(x[2 * i + 1]);
  float b0 = __uint_as_float(y[2 * i]), b1 = __uint_as_float(y[2 * i + 1]);
  float c0 = __uint_as_float(z[2 * i]), c1 = __uint_as_float(z[2 * i + 1]);
  u64 A = ((u64)x[2 * i + 1] << 32) | x[2 * i], B = ((u64)y[2 * i + 1] << 32) | y[2 * i], C = ((u64)z[2 * i + 1] << 32) | z[2 * i];
  auto same = [](float s, uint32_t p) { return (__float_as_uint(s) == p) || (s != s && __uint_as_float(p) != __uint_as_float(p)); };
  u64 r; int nb = 0;
  r = add2(A, B); nb += !same(__fadd_rn(a0, b0), (uint32_t)r) + !same(__fadd_rn(a1, b1), (uint32_t)(r >> 32));
  r = mul2(A, B); nb += !same(__fmul_rn(a0, b0), (uint32_t)r) + !same(__fmul_rn(a1, b1), (uint32_t)(r >> 32));
  r = fma2(A, B, C); nb += !same(__fmaf_rn(a0, b0, c0), (uint32_t)r) + !same(__fmaf_rn(a1, b1, c1), (uint32_t)(r >> 32));
  r = fma2(A, B, nz); nb += !same(__fmul_rn(a0, b0), (uint32_t)r) + !same(__fmul_rn(a1, b1), (uint32_t)(r >> 32));   // mul as fma(a,b,-0)
  if (nb) atomicAdd(bad, (unsigned long long)nb);
}

template <int MODE> double run(const char* name, int blocks, int iters, float* out) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench<MODE><<<blocks, 128>>>(out, 16, 1.f, 0x8000000080000000ull);
  cudaEventRecord(e0);
  bench<MODE><<<blocks, 128>>>(out, iters, 1.f, 0x8000000080000000ull);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double lane_ops = (double)blocks * 128 * iters * 16;       // scalar-equivalent fp ops (fma counted once)
  printf("%-28s %8.3f ms  %7.2f Gop/s scalar-equivalent  (%.2f ops/clk/SM at 1.965 GHz, 148 SMs)\n", name, ms, lane_ops / ms * 1e-6,
         lane_ops / (ms * 1e-3) / 1.965e9 / 148);
  return ms;
}

int main() {
  float* out; cudaMalloc(&out, 148 * 16 * 128 * sizeof(float));
  const int blocks = 148 * 16, iters = 20000;   // 16 blocks x 4 warps = 64 warps / SM
  run<0>("FFMA  scalar x16", blocks, iters, out);
  run<1>("FFMA2 packed x8", blocks, iters, out);
  run<2>("FADD  scalar x16", blocks, iters, out);
  run<3>("FADD2 packed x8", blocks, iters, out);
  run<4>("FMUL  scalar x16", blocks, iters, out);
  run<5>("FMUL2 packed x8", blocks, iters, out);
  run<6>("FFMA2 (a*b + runtime -0)", blocks, iters, out);
  run<7>("MUFU.RSQ x16", blocks, iters / 4, out);
  // low occupancy: 3 warps per SMSP (12 warps / SM), like the step kernel
  printf("-- 12 warps / SM --\n");
  run<0>("FFMA  scalar x16", 148 * 3, iters, out);
  run<1>("FFMA2 packed x8", 148 * 3, iters, out);
  run<3>("FADD2 packed x8", 148 * 3, iters, out);
  // exactness
  const int n = 1 << 24;
  std::vector<uint32_t> hx(n), hy(n), hz(n);
  std::mt19937 rng(42);
  for (int i = 0; i < n; ++i) {
    hx[i] = rng(); hy[i] = rng(); hz[i] = rng();
    if (i % 7 == 0) { hx[i] &= 0x807fffffu; }                       // denormal operands
    if (i % 11 == 0) { hy[i] = (hy[i] & 0x807fffffu) | 0x00800000u; hx[i] = (hx[i] & 0x807fffffu) | 0x3f000000u; }   // products near the denormal range
    if (i % 13 == 0) { hz[i] = hx[i] ^ 0x80000000u; hy[i] = 0x3f800000u; }                                           // exact cancellation -> signed zero
    if (i % 17 == 0) { hx[i] &= 0x80000000u; }                      // signed zeros
  }
  uint32_t *dx, *dy, *dz; unsigned long long* bad;
  cudaMalloc(&dx, n * 4); cudaMalloc(&dy, n * 4); cudaMalloc(&dz, n * 4); cudaMalloc(&bad, 8); cudaMemset(bad, 0, 8);
  cudaMemcpy(dx, hx.data(), n * 4, cudaMemcpyHostToDevice); cudaMemcpy(dy, hy.data(), n * 4, cudaMemcpyHostToDevice); cudaMemcpy(dz, hz.data(), n * 4, cudaMemcpyHostToDevice);
  check<<<(n / 2 + 255) / 256, 256>>>(dx, dy, dz, n, bad, 0x8000000080000000ull);
  unsigned long long hb = 0; cudaMemcpy(&hb, bad, 8, cudaMemcpyDeviceToHost);
  printf("exactness: %llu mismatching results out of %d (add2, mul2, fma2, fma2(a,b,-0) vs scalar .rn)\n", hb, 4 * n);
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return hb != 0;
}
