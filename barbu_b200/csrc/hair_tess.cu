// hair_tess.cu — the tess-stream stage on the device: interpolated render strands written as
// the GL_LINES vertex stream (xyz, relPos) that glDrawTransformFeedback consumes in the reference
// (src/shaders/hair/02_tess_stream/*.glsl, draw call src/fx/hair.cc:141-173).
//
// One thread per tessellated POINT (instance, patch, line, k = 0..nsubsegments): it evaluates the three Hermite curves of
// the patch at x = k / nsubsegments, the barycentric sample and relPos once, and stores the vertex as the end of segment
// k-1 and the start of segment k — two adjacent float4, so a warp writes one contiguous run of 128-bit stores.
// Write-bound: 32 B per segment out, 6 control points x 2 planes x 16 B per patch in (amortised over
// ninstances * nlines * (nsubsegments + 1) points, mostly L2 hits). Arithmetic: explicit .rn operations in the order the
// oracle (oracle/barbu_hair_oracle.c: bho_tess_stream) defines; what the reference leaves to the GL implementation
// (tess coordinates, primitive order, the random table) is defined there as well.
#include "hair_gen.cuh"

#include <cstdint>

namespace bh {

namespace {

__device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ float u01(uint32_t h) { return __fmul_rn((float)(h >> 8), 1.0f / 16777216.0f); }
__device__ __forceinline__ float dot4(float a0, float a1, float a2, float a3, float b0, float b1, float b2, float b3) {
  return __fadd_rn(__fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1)), __fadd_rn(__fmul_rn(a2, b2), __fmul_rn(a3, b3)));
}

__global__ void __launch_bounds__(256) tess_stream_kernel(const float4* __restrict__ pos, const float4* __restrict__ tan,
                                                          const int* __restrict__ patch, long long npatches, int N, float scale,
                                                          int ninstances, int nlines, int nsub, uint32_t seed,
                                                          float4* __restrict__ out) {
  const int npts = nsub + 1;
  const long long total = (long long)ninstances * npatches * nlines * npts;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(q % npts);
    long long r = q / npts;
    const int line = (int)(r % nlines); r /= nlines;
    const long long pa = r % npatches;
    const int inst = (int)(r / npatches);
    const int* e = patch + 6 * pa;
    const float x = __fdiv_rn((float)k, (float)nsub);
    // hermite basis: vU * mHermit (inc_maths.glsl:210-228)
    const float u2 = __fmul_rn(x, x), u3 = __fmul_rn(u2, x);
    const float h0 = dot4(u3, u2, x, 1.0f, 2.0f, -3.0f, 0.0f, 1.0f);
    const float h1 = dot4(u3, u2, x, 1.0f, -2.0f, 3.0f, 0.0f, 0.0f);
    const float h2 = dot4(u3, u2, x, 1.0f, 1.0f, -2.0f, 1.0f, 0.0f);
    const float h3 = dot4(u3, u2, x, 1.0f, 1.0f, -1.0f, 0.0f, 0.0f);
    float qx[3], qy[3], qz[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float4 p0 = __ldg(pos + e[2 * c]), p1 = __ldg(pos + e[2 * c + 1]);
      float4 t0 = __ldg(tan + e[2 * c]), t1 = __ldg(tan + e[2 * c + 1]);
      t0.x = __fmul_rn(t0.x, scale); t0.y = __fmul_rn(t0.y, scale); t0.z = __fmul_rn(t0.z, scale);   // tcs_stream_hair.glsl:41
      t1.x = __fmul_rn(t1.x, scale); t1.y = __fmul_rn(t1.y, scale); t1.z = __fmul_rn(t1.z, scale);
      qx[c] = dot4(h0, h1, h2, h3, p0.x, p1.x, t0.x, t1.x);
      qy[c] = dot4(h0, h1, h2, h3, p0.y, p1.y, t0.y, t1.y);
      qz[c] = dot4(h0, h1, h2, h3, p0.z, p1.z, t0.z, t1.z);
    }
    // random barycentric sample of the (instance, line): index formula of tes_stream_hair.glsl:52-54
    const float y = __fdiv_rn((float)line, (float)nlines);
    const int idx = (int)__fadd_rn(__fmul_rn(y, 40.0f), (float)inst) % 4096;
    float sx = u01(lowbias32(seed + 0x9E3779B9u * (uint32_t)(2 * idx + 1)));
    float sy = u01(lowbias32(seed + 0x9E3779B9u * (uint32_t)(2 * idx + 2)));
    if (__fadd_rn(sx, sy) > 1.0f) {                       // sample_triangle2, inc_maths.glsl:107-115
      sx = fmaxf(sx, sy);
      sy = fminf(sx, sy);
      sx = __fsub_rn(1.0f, sx);
    }
    const float sz = __fsub_rn(1.0f, __fadd_rn(sx, sy));
    const float px = __fadd_rn(__fadd_rn(__fmul_rn(qx[0], sx), __fmul_rn(qx[1], sy)), __fmul_rn(qx[2], sz));
    const float py = __fadd_rn(__fadd_rn(__fmul_rn(qy[0], sx), __fmul_rn(qy[1], sy)), __fmul_rn(qy[2], sz));
    const float pz = __fadd_rn(__fadd_rn(__fmul_rn(qz[0], sx), __fmul_rn(qz[1], sy)), __fmul_rn(qz[2], sz));
    // relPos = smoothstep2(0, 1, relPos(first control point) + x / N)   (tes:63-64, vs:24)
    const float rel0 = __fdiv_rn((float)(e[0] % N), (float)N);
    float t = __fdiv_rn(__fsub_rn(__fadd_rn(rel0, __fdiv_rn(x, (float)N)), 0.0f), __fsub_rn(1.0f, 0.0f));
    t = fminf(fmaxf(t, 0.0f), 1.0f);
    const float rel = __fmul_rn(__fmul_rn(__fmul_rn(t, t), t),
                                __fadd_rn(10.0f, __fmul_rn(t, __fadd_rn(-15.0f, __fmul_rn(6.0f, t)))));
    const float4 v = make_float4(px, py, pz, rel);
    const long long base = (((long long)inst * npatches + pa) * nlines + line) * nsub * 2;
    if (k < nsub) __stcs(out + base + 2 * k, v);
    if (k > 0) __stcs(out + base + 2 * (k - 1) + 1, v);
  }
}

}  // namespace

cudaError_t launch_tess_stream(const float4* pos, const float4* tan, const int* patch, long long npatches, int nverts, float scale,
                               int ninstances, int nlines, int nsub, unsigned seed, float4* out, cudaStream_t stream) {
  const long long total = (long long)ninstances * npatches * nlines * (nsub + 1);
  if (total <= 0) return cudaSuccess;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 64) blocks = 148LL * 64;
  tess_stream_kernel<<<(unsigned)blocks, 256, 0, stream>>>(pos, tan, patch, npatches, nverts, scale, ninstances, nlines, nsub, seed, out);
  return cudaGetLastError();
}

}  // namespace bh
