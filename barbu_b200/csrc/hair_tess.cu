// hair_tess.cu — the tess-stream stage on the device: interpolated render strands written as
// the GL_LINES vertex stream (xyz, relPos) that glDrawTransformFeedback consumes in the reference
// (src/shaders/hair/02_tess_stream/*.glsl, draw call src/fx/hair.cc:141-173).
//
// One thread per curve POINT (patch, k = 0..nsubsegments): it evaluates the three Hermite curves of the patch there once, then
// walks the (instance, isoline) pairs — whose only own work is the barycentric mix — takes the far end of its segment from
// the next lane by shuffle and writes the segment as one aligned 32-byte store (two float4). Consecutive lanes own
// consecutive segments, so a warp writes contiguous 512-byte runs, every output sector exactly once. The (instance, isoline) samples are hashed once per block into
// shared memory. Write-bound: 32 B per segment out; 6 control points x 2 planes x 16 B per patch in, once per patch row.
// (The first version ran one thread per POINT and instance: 12 control-point loads, the Hermite basis and the sample hash
// per point — instruction-bound at 0.26 of the HBM write roofline, profiles/r01_stages.txt.)
// Arithmetic: explicit .rn operations in the order the
// oracle (oracle/barbu_hair_oracle.c: bho_tess_stream) defines; what the reference leaves to the GL implementation
// (tess coordinates, primitive order, the random table) is defined there as well.
#include "hair_gen.cuh"

#include <cstdint>
#include <cstdlib>

namespace bh {

namespace {

__device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ float u01(uint32_t h) { return __fmul_rn((float)(h >> 8), 1.0f / 16777216.0f); }
__device__ __forceinline__ float dot4(float a0, float a1, float a2, float a3, float b0, float b1, float b2, float b3) {
  // vec4 * mat4, one column: the four products summed left to right (third_party/glm/glm/detail/type_mat4x4.inl:584-595)
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1)), __fmul_rn(a2, b2)), __fmul_rn(a3, b3));
}

struct Sample { float x, y, z; };

// random barycentric sample of (instance, isoline): index formula of tes_stream_hair.glsl:52-54, sample_triangle2 of
// inc_maths.glsl:107-115
__device__ __forceinline__ Sample tess_sample(int inst, int line, int nlines, uint32_t seed) {
  const float y = __fdiv_rn((float)line, (float)nlines);
  const int idx = (int)__fadd_rn(__fmul_rn(y, 40.0f), (float)inst) % 4096;
  float sx = u01(lowbias32(seed + 0x9E3779B9u * (uint32_t)(2 * idx + 1)));
  float sy = u01(lowbias32(seed + 0x9E3779B9u * (uint32_t)(2 * idx + 2)));
  if (__fadd_rn(sx, sy) > 1.0f) {
    sx = fmaxf(sx, sy);
    sy = fminf(sx, sy);
    sx = __fsub_rn(1.0f, sx);
  }
  return { sx, sy, __fsub_rn(1.0f, __fadd_rn(sx, sy)) };
}

struct CurvePoint { float qx[3], qy[3], qz[3], rel; };

// The three Hermite curves of a patch at x = k / nsub, and relPos there. P0/P1/T0/T1: the control points (tangents scaled).
__device__ __forceinline__ CurvePoint tess_point(int k, int nsub, int N, float rel0, const float4 (&P0)[3], const float4 (&P1)[3],
                                                 const float4 (&T0)[3], const float4 (&T1)[3]) {
  CurvePoint r;
  const float x = __fdiv_rn((float)k, (float)nsub);
  // hermite basis: vU * mHermit (inc_maths.glsl:210-228)
  const float u2 = __fmul_rn(x, x), u3 = __fmul_rn(u2, x);
  const float h0 = dot4(u3, u2, x, 1.0f, 2.0f, -3.0f, 0.0f, 1.0f);
  const float h1 = dot4(u3, u2, x, 1.0f, -2.0f, 3.0f, 0.0f, 0.0f);
  const float h2 = dot4(u3, u2, x, 1.0f, 1.0f, -2.0f, 1.0f, 0.0f);
  const float h3 = dot4(u3, u2, x, 1.0f, 1.0f, -1.0f, 0.0f, 0.0f);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    r.qx[c] = dot4(h0, h1, h2, h3, P0[c].x, P1[c].x, T0[c].x, T1[c].x);
    r.qy[c] = dot4(h0, h1, h2, h3, P0[c].y, P1[c].y, T0[c].y, T1[c].y);
    r.qz[c] = dot4(h0, h1, h2, h3, P0[c].z, P1[c].z, T0[c].z, T1[c].z);
  }
  // relPos = smoothstep2(0, 1, relPos(first control point) + x / N)   (tes:63-64, vs:24)
  float t = __fdiv_rn(__fsub_rn(__fadd_rn(rel0, __fdiv_rn(x, (float)N)), 0.0f), __fsub_rn(1.0f, 0.0f));
  t = fminf(fmaxf(t, 0.0f), 1.0f);
  r.rel = __fmul_rn(__fmul_rn(__fmul_rn(t, t), t), __fadd_rn(10.0f, __fmul_rn(t, __fadd_rn(-15.0f, __fmul_rn(6.0f, t)))));
  return r;
}

__device__ __forceinline__ float4 tess_mix(const CurvePoint& q, Sample s) {
  return make_float4(__fadd_rn(__fadd_rn(__fmul_rn(q.qx[0], s.x), __fmul_rn(q.qx[1], s.y)), __fmul_rn(q.qx[2], s.z)),
                     __fadd_rn(__fadd_rn(__fmul_rn(q.qy[0], s.x), __fmul_rn(q.qy[1], s.y)), __fmul_rn(q.qy[2], s.z)),
                     __fadd_rn(__fadd_rn(__fmul_rn(q.qz[0], s.x), __fmul_rn(q.qz[1], s.y)), __fmul_rn(q.qz[2], s.z)), q.rel);
}

// one aligned 32-byte streaming store: both vertices of a segment
__device__ __forceinline__ void store_segment(float4* dst, float4 a, float4 b) {
  asm volatile("st.global.cs.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w),
               "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}

constexpr int kTessSmemSamples = 2048;           // (instance, isoline) pairs cached per block; beyond that they are hashed in place

__global__ void __launch_bounds__(256) tess_stream_points_kernel(const float4* __restrict__ pos, const float4* __restrict__ tan,
                                                          const int* __restrict__ patch, long long npatches, int N, float scale,
                                                          int ninstances, int nlines, int nsub, uint32_t seed,
                                                          float4* __restrict__ out) {
  __shared__ Sample smp[kTessSmemSamples];
  const int ncurves = ninstances * nlines;
  const bool cached = ncurves <= kTessSmemSamples;
  if (cached) {
    for (int i = threadIdx.x; i < ncurves; i += blockDim.x) smp[i] = tess_sample(i / nlines, i % nlines, nlines, seed);
    __syncthreads();
  }
  // One thread per curve POINT; a warp covers 32 consecutive points of the (patch, k = 0..nsub) sequence and owns the 31
  // segments between them (consecutive warps overlap by one point), so every point's Hermite evaluation happens once per
  // warp and the far end of a segment arrives by shuffle. Pairs that straddle two patches are not segments and write nothing.
  const int npts = nsub + 1;
  const long long total = npatches * npts;
  const long long nwarps_total = (total - 1 + 30) / 31;                       // warp w starts at point 31 * w
  const long long line_stride = 2LL * nsub;                                   // float4 per (instance, patch, isoline)
  const long long inst_stride = npatches * nlines * line_stride;
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, warp_step = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long w = warp0; w < nwarps_total; w += warp_step) {
    const long long g = w * 31 + lane;
    const bool live = g < total;
    const long long gg = live ? g : total - 1;                                // clamp: every lane takes part in the shuffles
    const int k = (int)(gg % npts);
    const long long pa = gg / npts;
    const int2* e2 = reinterpret_cast<const int2*>(patch + 6 * pa);           // 24-byte records: 8-byte aligned
    const int2 ea = __ldg(e2), eb = __ldg(e2 + 1), ec = __ldg(e2 + 2);
    const int e[6] = { ea.x, ea.y, eb.x, eb.y, ec.x, ec.y };
    float4 P0[3], P1[3], T0[3], T1[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      P0[c] = __ldg(pos + e[2 * c]); P1[c] = __ldg(pos + e[2 * c + 1]);
      T0[c] = __ldg(tan + e[2 * c]); T1[c] = __ldg(tan + e[2 * c + 1]);
      T0[c].x = __fmul_rn(T0[c].x, scale); T0[c].y = __fmul_rn(T0[c].y, scale); T0[c].z = __fmul_rn(T0[c].z, scale);   // tcs_stream_hair.glsl:41
      T1[c].x = __fmul_rn(T1[c].x, scale); T1[c].y = __fmul_rn(T1[c].y, scale); T1[c].z = __fmul_rn(T1[c].z, scale);
    }
    const float rel0 = __fdiv_rn((float)(e[0] % N), (float)N);
    const CurvePoint a = tess_point(k, nsub, N, rel0, P0, P1, T0, T1);
    const float rel_next = __shfl_down_sync(0xffffffffu, a.rel, 1);
    const bool writes = live && lane < 31 && k < nsub && g + 1 < total;       // this lane starts a segment whose end is in lane + 1
    float4* dst_inst = out + pa * nlines * line_stride + 2 * k;               // (instance 0, isoline 0) of this patch
    for (int inst = 0; inst < ninstances; ++inst, dst_inst += inst_stride) {
      float4* dst = dst_inst;
      for (int line = 0; line < nlines; ++line, dst += line_stride) {
        const Sample sm = cached ? smp[inst * nlines + line] : tess_sample(inst, line, nlines, seed);
        const float4 v = tess_mix(a, sm);
        const float4 vn = make_float4(__shfl_down_sync(0xffffffffu, v.x, 1), __shfl_down_sync(0xffffffffu, v.y, 1),
                                      __shfl_down_sync(0xffffffffu, v.z, 1), rel_next);
        if (writes) store_segment(dst, v, vn);
      }
    }
  }
}

// Variant for many (instance, isoline) pairs per patch: one thread per SEGMENT evaluates both of its ends itself (twice the
// Hermite work, no shuffles, no idle lanes at patch boundaries); ahead once the barycentric loop dominates.
__global__ void __launch_bounds__(256) tess_stream_segments_kernel(const float4* __restrict__ pos, const float4* __restrict__ tan,
                                                          const int* __restrict__ patch, long long npatches, int N, float scale,
                                                          int ninstances, int nlines, int nsub, uint32_t seed,
                                                          float4* __restrict__ out) {
  __shared__ Sample smp[kTessSmemSamples];
  const int ncurves = ninstances * nlines;
  const bool cached = ncurves <= kTessSmemSamples;
  if (cached) {
    for (int i = threadIdx.x; i < ncurves; i += blockDim.x) smp[i] = tess_sample(i / nlines, i % nlines, nlines, seed);
    __syncthreads();
  }
  const long long total = npatches * nsub;
  const long long line_stride = 2LL * nsub;                                   // float4 per (instance, patch, isoline)
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(q % nsub);
    const long long pa = q / nsub;
    const int2* e2 = reinterpret_cast<const int2*>(patch + 6 * pa);           // 24-byte records: 8-byte aligned
    const int2 ea = __ldg(e2), eb = __ldg(e2 + 1), ec = __ldg(e2 + 2);
    const int e[6] = { ea.x, ea.y, eb.x, eb.y, ec.x, ec.y };
    float4 P0[3], P1[3], T0[3], T1[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      P0[c] = __ldg(pos + e[2 * c]); P1[c] = __ldg(pos + e[2 * c + 1]);
      T0[c] = __ldg(tan + e[2 * c]); T1[c] = __ldg(tan + e[2 * c + 1]);
      T0[c].x = __fmul_rn(T0[c].x, scale); T0[c].y = __fmul_rn(T0[c].y, scale); T0[c].z = __fmul_rn(T0[c].z, scale);   // tcs_stream_hair.glsl:41
      T1[c].x = __fmul_rn(T1[c].x, scale); T1[c].y = __fmul_rn(T1[c].y, scale); T1[c].z = __fmul_rn(T1[c].z, scale);
    }
    const float rel0 = __fdiv_rn((float)(e[0] % N), (float)N);
    const CurvePoint a = tess_point(k, nsub, N, rel0, P0, P1, T0, T1), b = tess_point(k + 1, nsub, N, rel0, P0, P1, T0, T1);
    float4* seg = out + pa * nlines * line_stride + 2 * k;                    // (instance 0, isoline 0) of this patch
    const long long inst_stride = npatches * nlines * line_stride;
    for (int inst = 0; inst < ninstances; ++inst, seg += inst_stride) {
      float4* dst = seg;
      for (int line = 0; line < nlines; ++line, dst += line_stride) {
        const Sample s = cached ? smp[inst * nlines + line] : tess_sample(inst, line, nlines, seed);
        store_segment(dst, tess_mix(a, s), tess_mix(b, s));
      }
    }
  }
}

}  // namespace

cudaError_t launch_tess_stream(const float4* pos, const float4* tan, const int* patch, long long npatches, int nverts, float scale,
                               int ninstances, int nlines, int nsub, unsigned seed, float4* out, cudaStream_t stream) {
  if (npatches <= 0 || ninstances <= 0 || nlines <= 0) return cudaSuccess;
  // measured on B200 (profiles/r01_stages.txt): the point kernel wins while the per-point Hermite evaluation dominates
  // (0.91 vs 0.71 of the HBM write roofline at the reference's 3 x 2 pairs), the segment kernel once the pair loop does
  // (1.03 vs 0.90 at 6 x 4)
  static const int pairs_env = [] { const char* e = getenv("BH_TESS_POINT_KERNEL_MAX_PAIRS"); return e ? atoi(e) : -1; }();   // tuning knob
  const int max_pairs = pairs_env >= 0 ? pairs_env : 7;                      // crossover measured at 8 pairs (0.86 vs 0.87)
  if ((long long)ninstances * nlines <= max_pairs) {
    const long long total = npatches * (nsub + 1);                             // one thread per curve point, 31 new points per warp
    long long blocks = ((total + 30) / 31 * 32 + 255) / 256;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    tess_stream_points_kernel<<<(unsigned)blocks, 256, 0, stream>>>(pos, tan, patch, npatches, nverts, scale, ninstances, nlines, nsub, seed, out);
  } else {
    const long long total = npatches * nsub;                                   // one thread per (patch, segment)
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    tess_stream_segments_kernel<<<(unsigned)blocks, 256, 0, stream>>>(pos, tan, patch, npatches, nverts, scale, ninstances, nlines, nsub, seed, out);
  }
  return cudaGetLastError();
}

}  // namespace bh
