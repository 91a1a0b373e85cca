// Probe (not product code): where does the branch-free exact 1/sqrt sequence hold, and does a MUFU-free reciprocal
// seeded with the rsqrt estimate still round correctly? Exhaustive over all positive finite binary32.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I barbu_b200/csrc -o isqrt_probe tools/ubench/isqrt_probe.cu
#include <cstdio>
#include <cstdint>
#include "hair_math.cuh"
using namespace bh;
__device__ __forceinline__ float variant_seeded(float x) {   // rcp seeded with y = rsqrt(x) instead of MUFU.RCP(s)
  float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  const float s0 = __fmul_rn(x, y), h = __fmul_rn(y, 0.5f);
  const float s = __fmaf_rn(__fmaf_rn(-s0, s0, x), h, s0);
  return __fmaf_rn(y, __fmaf_rn(-y, s, 1.0f), y);
}
__device__ __forceinline__ float variant_seeded2(float x) {  // two Newton steps on the seeded reciprocal
  float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  const float s0 = __fmul_rn(x, y), h = __fmul_rn(y, 0.5f);
  const float s = __fmaf_rn(__fmaf_rn(-s0, s0, x), h, s0);
  const float r1 = __fmaf_rn(y, __fmaf_rn(-y, s, 1.0f), y);
  return __fmaf_rn(r1, __fmaf_rn(-r1, s, 1.0f), r1);
}
__global__ void probe(unsigned long long* bad_by_exp, unsigned long long* bad_seeded, unsigned long long* bad_seeded2) {
  for (unsigned long long b = 0x00800000ull + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < 0x7f800000ull;
       b += (unsigned long long)gridDim.x * blockDim.x) {
    const float x = __uint_as_float((unsigned)b);
    const float want = __frcp_rn(__fsqrt_rn(x));
    const int e = (int)(b >> 23);
    if (__float_as_uint(MathExact::inversesqrt_in_range(x)) != __float_as_uint(want)) atomicAdd(&bad_by_exp[e], 1ull);
    if (__float_as_uint(variant_seeded(x)) != __float_as_uint(want)) atomicAdd(&bad_seeded[e], 1ull);
    if (__float_as_uint(variant_seeded2(x)) != __float_as_uint(want)) atomicAdd(&bad_seeded2[e], 1ull);
  }
}
int main() {
  unsigned long long *d, h[3 * 256];
  cudaMalloc(&d, sizeof h); cudaMemset(d, 0, sizeof h);
  probe<<<148 * 8, 256>>>(d, d + 256, d + 512);
  cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  printf("error: %s\n", cudaGetErrorString(cudaGetLastError()));
  for (int v = 0; v < 3; ++v) {
    unsigned long long tot = 0; int first = -1, last = -1, lo_ok = -1, hi_ok = -1;
    for (int e = 1; e < 255; ++e) { tot += h[v * 256 + e]; if (h[v * 256 + e]) { if (first < 0) first = e; last = e; } }
    // widest contiguous clean exponent range containing 127
    int a = 127, b = 127; while (a > 1 && !h[v * 256 + a - 1]) --a; while (b < 254 && !h[v * 256 + b + 1]) ++b;
    (void)lo_ok; (void)hi_ok;
    printf("variant %d: mismatches %llu; clean biased-exponent range around 1.0: [%d, %d] (2^%d .. 2^%d); mismatching exponents %d..%d\n",
           v, tot, a, b, a - 127, b - 127 + 1, first, last);
    if (v > 0) { printf("  per exponent 120..134:"); for (int e = 120; e < 135; ++e) printf(" %llu", h[v * 256 + e]); printf("\n"); }
  }
  return 0;
}
