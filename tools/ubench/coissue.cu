// Does a packed FFMA2 hold the warp scheduler's issue port for its second pipe cycle, or can another pipe's instruction
// issue there? Eight independent FFMA2 chains per thread, with M independent filler operations of another pipe per FFMA2,
// at 1 / 2 / 3 warps per scheduler on ONE SM; cycles per FFMA2 (per scheduler) from clock64.
//   filler 0: none   1: LOP3 (ALU)   2: FMNMX (ALU)   3: MUFU.RSQ (XU)   4: scalar FFMA (FMA pipe, for comparison)
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int FILL, int M> __global__ void k(long long* out, float* sink, u64 seed2, unsigned useed, float fseed) {
  u64 p[8]; unsigned u[8]; float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { p[i] = seed2 + i; u[i] = useed + i + threadIdx.x; f[i] = fseed + i; }
  const u64 c2 = seed2 ^ 0x0000100000001000ull;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 512; ++it) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        p[i] = fma2(p[i], c2, c2);
#pragma unroll
        for (int m = 0; m < M; ++m) {
          const int x = (i + 3 * m) & 7;
          if (FILL == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[x]) : "r"(useed), "r"(u[(x + 1) & 7]));
          if (FILL == 2) asm volatile("min.f32 %0, %0, %1;" : "+f"(f[x]) : "f"(f[(x + 1) & 7]));
          if (FILL == 3) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(f[x]));
          if (FILL == 4) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[x]) : "f"(fseed));
        }
      }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if ((threadIdx.x & 31) == 0) out[threadIdx.x >> 5] = t1 - t0;
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += __uint_as_float((unsigned)p[i]) + __uint_as_float(u[i]) + f[i];
  sink[threadIdx.x] = acc;
}

template <int FILL, int M> void run(const char* name, long long* d, float* s) {
  for (int wps = 1; wps <= 4; ++wps) {
    const int warps = 4 * wps;
    k<FILL, M><<<1, warps * 32>>>(d, s, 0x3f8000013f800001ull, 0x1234567u, 1.0001f);
    long long h[16]; cudaMemcpy(h, d, sizeof(long long) * warps, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < warps; ++i) mx = h[i] > mx ? h[i] : mx;
    // per scheduler: wps warps x 512 x 16 FFMA2 each
    printf("%-34s %d warps/scheduler: %.3f cycles per FFMA2 (+%d filler)\n", name, wps, (double)mx / (512.0 * 16.0 * wps), M);
  }
}
int main() {
  long long* d; float* s; cudaMalloc(&d, 256); cudaMalloc(&s, 4096);
  run<0, 0>("FFMA2 alone", d, s);
  run<1, 1>("FFMA2 + 1 LOP3", d, s);
  run<1, 2>("FFMA2 + 2 LOP3", d, s);
  run<1, 3>("FFMA2 + 3 LOP3", d, s);
  run<2, 1>("FFMA2 + 1 FMNMX", d, s);
  run<2, 2>("FFMA2 + 2 FMNMX", d, s);
  run<3, 1>("FFMA2 + 1 MUFU.RSQ", d, s);
  run<4, 1>("FFMA2 + 1 FFMA", d, s);
  run<4, 2>("FFMA2 + 2 FFMA", d, s);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
