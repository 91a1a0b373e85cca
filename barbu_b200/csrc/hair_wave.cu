// hair_wave.cu — the LATENCY-oriented step kernel for small scalps (the reference's own asset is 448 roots x 4 control points,
// configs[0] is 4,096 x 16): same arithmetic as hair_step.cu / hair_stream.cu, organised for the shortest dependency chain
// instead of the fewest instructions.
//
// The streaming kernel (hair_stream.cu) gives a strand to one thread and runs the 8 constraint iterations as an 8-stage
// software pipeline inside it: minimal instruction count and HBM traffic, but a shard of a few hundred strands is then a
// handful of warps that each walk 16+ pipeline steps of ~1,000 cycles — 10-14 us per launch whatever the size. Here the
// iterations of ONE strand are spread over 8 LANES (candidate C): at step t lane k applies iteration k+1 to
// vertex 1 + t - k, takes its input C(i, k) from lane k-1 by a shuffle and keeps D(i-1, k+1) in a register. A strand of N
// vertices is done after N + 6 steps of one projection each, 4 strands share a warp, and a scalp of S strands spreads over
// S / 4 warps: about twice the instructions per vertex, a quarter of the latency, 8x the parallelism. Selected by
// bh_set_step_policy (BH_POLICY_LATENCY / BH_POLICY_AUTO); the throughput kernels stay the default.
//
// The substeps of a frame can run as passes of one launch here without any of the streaming kernel's group logic: a warp
// owns its strands entirely, so a pass re-reads what the same warp stored (ordered by __syncwarp).
//
// Per strand this is exactly one dispatch of cs_simulation.glsl:170-208 + PingPongBuffer::swap, in place (see hair_step.cu).
#include "hair_step.cuh"
#include "hair_math.cuh"
#include "hair_collide.cuh"

namespace bh {

namespace {

constexpr int kLanes = 8;                        // constraint iterations == lanes per strand
constexpr int kThreads = 128;                    // 16 strands per block

// vec3(mat4(1.0) * vec4(p, 1.0)) in GLM's operation order (see root_transform in hair_step.cu)
template <class M>
__device__ __forceinline__ V3 wave_root_transform(V3 p) {
  const float zx = M::mul(0.0f, p.x), zy = M::mul(0.0f, p.y), zz = M::mul(0.0f, p.z);
  const float z1 = M::mul(0.0f, 1.0f);
  return { M::add(M::add(M::mul(1.0f, p.x), zy), M::add(zz, z1)),
           M::add(M::add(zx, M::mul(1.0f, p.y)), M::add(zz, z1)),
           M::add(M::add(zx, zy), M::add(M::mul(1.0f, p.z), z1)) };
}

template <class M, bool CAPS>
__global__ void __launch_bounds__(kThreads) hair_step_wave_kernel(const __grid_constant__ StepArgs a) {
  const int k = threadIdx.x & (kLanes - 1);                                  // this lane's iteration is k + 1
  const long long strand = ((long long)blockIdx.x * kThreads + threadIdx.x) / kLanes;
  const bool active = strand < a.nstrands;                                   // an idle group still takes part in the shuffles
  const long long sidx = active ? strand : a.nstrands - 1;
  const int N = a.nverts;
  float4* P = a.pos + sidx * N;
  float4* V = a.vel + sidx * N;
  const int passes = a.passes > 1 ? a.passes : 1;

  for (int pass = 0; pass < passes; ++pass) {
    // every lane knows the root: D(0, k) = X[0] for all k (cs:190-192: the root is pinned)
    const float4 P0 = __ldcg(&P[0]);                                          // L2, not L1: a later pass reads what another lane stored
    const V3 x0 = wave_root_transform<M>(V3{ P0.x, P0.y, P0.z });
    V3 xp = x0;                                                              // D(i-1, k+1) of the vertex this lane handled last
    V3 outC = { 0.f, 0.f, 0.f };                                             // C(i, k+1): what lane k+1 takes next step
    float outRest = 0.f;
    V3 heldD = x0, heldd = { 0.f, 0.f, 0.f };                                // lane 7: the vertex waiting for d_{i+1}
    float heldRest = 0.f;
    const int steps = N - 1 + kLanes - 1;                                    // vertices 1 .. N-1 through 8 lanes
    for (int t = 0; t < steps; ++t) {
      // input of this step: C(i, k) from the lane below, as it left the previous step (lane 0: the integrated vertex)
      V3 prev = { __shfl_up_sync(0xffffffffu, outC.x, 1, kLanes), __shfl_up_sync(0xffffffffu, outC.y, 1, kLanes),
                  __shfl_up_sync(0xffffffffu, outC.z, 1, kLanes) };
      float rest = __shfl_up_sync(0xffffffffu, outRest, 1, kLanes);
      const int i = 1 + t - k;
      const bool valid = i >= 1 && i < N;
      if (k == 0 && valid) {
        const float4 p = __ldcg(&P[i]);
        float4 v = __ldcg(&V[i]);
        if (a.use_drag) { v.x = M::mul(v.x, a.keep); v.y = M::mul(v.y, a.keep); v.z = M::mul(v.z, a.keep); }
        prev = { __fmaf_rn(a.dt2, a.fx, __fmaf_rn(a.dt, v.x, p.x)), __fmaf_rn(a.dt2, a.fy, __fmaf_rn(a.dt, v.y, p.y)),
                 __fmaf_rn(a.dt2, a.fz, __fmaf_rn(a.dt, v.z, p.z)) };         // cs:181-182
        rest = p.w;
      }
      if (valid) {
        const V3 vd = vsub<M>(prev, xp);
        const float dp = M::dot(vd, vd);
        const float inv = (M::kRangeChecked && !M::in_fast_range(dp)) ? M::inversesqrt(dp) : M::inversesqrt_in_range(dp);
        const V3 D = M::project(xp, vd, inv, M::mul(a.sf, rest));             // cs:114
        xp = D;
        if (k < kLanes - 1) {
          outC = collide_all_pos<M, CAPS>(a, D);                              // cs:129-153, velocity dead before the last iteration
          outRest = rest;
        } else {
          const V3 d = vsub<M>(D, prev);                                      // cs:116
          if (i >= 2) {                                                       // vertex i-1 gets its velocity from d_i (cs:119-121)
            V3 w = M::scale(d, a.damp), q = heldD;
            collide_all_pos_vel<M, CAPS>(a, q, w);
            if (active) { P[i - 1] = make_float4(q.x, q.y, q.z, heldRest); V[i - 1] = make_float4(w.x, w.y, w.z, 0.f); }
          }
          heldD = D; heldd = d; heldRest = rest;
        }
      }
    }
    if (k == kLanes - 1 && active) {
      if (N >= 2) {                                                           // the tip keeps its own d
        V3 w = heldd, q = heldD;
        collide_all_pos_vel<M, CAPS>(a, q, w);
        P[N - 1] = make_float4(q.x, q.y, q.z, heldRest); V[N - 1] = make_float4(w.x, w.y, w.z, 0.f);
      }
      const V3 v0 = vsub<M>(x0, V3{ P0.x, P0.y, P0.z });                      // p.velocity = p.position - lastPosition (cs:192)
      P[0] = make_float4(x0.x, x0.y, x0.z, P0.w); V[0] = make_float4(v0.x, v0.y, v0.z, 0.f);
    }
    __syncwarp();                                                             // the next pass reads what lane 7 just stored
  }
}

template <class M, bool CAPS>
cudaError_t launch_wave_t(const StepArgs& a, cudaStream_t stream) {
  const long long strands_per_block = kThreads / kLanes;
  const long long blocks = (a.nstrands + strands_per_block - 1) / strands_per_block;
  if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
  hair_step_wave_kernel<M, CAPS><<<(unsigned)blocks, kThreads, 0, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace

bool wave_kernel_eligible(const StepArgs& a) {
  return a.iterations == kLanes && a.nverts >= 1 && a.nstrands >= 1 && a.ncaps >= 0 && a.ncaps <= kMaxCapsules;
}

cudaError_t launch_step_wave(const StepArgs& a, int math, cudaStream_t stream) {
  const bool caps = a.ncaps > 0;
  if (math == 0) return caps ? launch_wave_t<MathExact, true>(a, stream) : launch_wave_t<MathExact, false>(a, stream);
  return caps ? launch_wave_t<MathFast, true>(a, stream) : launch_wave_t<MathFast, false>(a, stream);
}

}  // namespace bh
