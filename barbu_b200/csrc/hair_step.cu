// hair_step.cu — the fused hair-strand step for sm_100a.
//
// What one launch computes per strand is exactly one dispatch of the reference shader
// (src/shaders/hair/01_simulation/cs_simulation.glsl:170-208) followed by PingPongBuffer::swap
// (src/memory/pingpong_buffer.cc:73-84), written back in place:
//
//   X[i] = fma(dt*dt, F, fma(dt, vel[i], pos[i]))           i >= 1          (cs:181-182)
//   X[0] = pos[0]  (root pinned, vel 0)                                      (cs:190-192)
//   K = 8 times:                                                             (cs:155-159)
//       serial FTL pass   D(i,k) = D(i-1,k) + sf*rest_i * normalize(C(i,k-1) - D(i-1,k))   (cs:110-117)
//       collision         C(i,k) = pushed out of the sphere (and capsules)                   (cs:129-153)
//   velocity fix: vel[i] = 0.8 * (D(i+1,K) - C(i+1,K-1)),  vel[N-1] = D(N-1,K) - C(N-1,K-1),
//       reflected where C(i,K) collided                                      (cs:116,119-121,137)
//
// Mapping to the machine. The only dependencies are D(i,k) <- D(i-1,k), C(i,k-1); so one THREAD owns
// one strand and runs a K-stage software pipeline along it: at step t stage k handles vertex t-k.
// The K stage bodies of a step are independent of each other (instruction-level parallelism K),
// the live state is K*(3+4) registers whatever the strand length, and a vertex is read once at
// step t and written once at step t+K. The HBM side is a pure stream: a warp owns 32 consecutive
// strands, moves them in 8-vertex chunks (one 128-byte line per strand per plane) with coalesced
// 128-bit loads/stores through a padded shared-memory transpose buffer, and because K == chunk
// length the output of vertex t-8 reuses the slot vertex t was just read from.
// Algorithmic traffic: 16 B pos + 16 B vel read, the same written = 64 B per vertex per launch.
#include "hair_step.cuh"
#include "hair_math.cuh"
#include "hair_collide.cuh"

#include <cstdlib>

namespace bh {

namespace {

constexpr int kChunk = 8;                 // vertices per staged chunk == pipeline depth
constexpr int kPitch = kChunk + 1;        // float4 per lane slot row: 144 B -> conflict-free LDS.128/STS.128
constexpr int kWarpsPerBlock = 4;
constexpr int kThreads = kWarpsPerBlock * 32;
// per warp: position slots, velocity slots (32 * kPitch float4 each) + rest-length ring (8 * 32 floats)
constexpr int kWarpSmemBytes = 2 * 32 * kPitch * 16 + kChunk * 32 * 4;

// One FTL projection: D = p0 + L * normalize(prev - p0).
template <class M>
__device__ __forceinline__ V3 ftl(V3 p0, V3 prev, float L) {
  const V3 vd = vsub<M>(prev, p0);
  return M::project(p0, vd, M::inversesqrt(M::dot(vd, vd)), L);
}

// Root vertex: vec3(mat4(1.0) * vec4(p, 1.0)) in GLM's mat4*vec4 order (type_mat4x4.inl:563-575):
// (m0*x + m1*y) + (m2*z + m3*1) with the identity's zeros and ones multiplied through, so that
// -0.0 and non-finite roots come out exactly as the reference host arithmetic gives them.
template <class M>
__device__ __forceinline__ V3 root_transform(V3 p) {
  const float zx = M::mul(0.0f, p.x), zy = M::mul(0.0f, p.y), zz = M::mul(0.0f, p.z);
  const float z1 = M::mul(0.0f, 1.0f);
  return { M::add(M::add(M::mul(1.0f, p.x), zy), M::add(zz, z1)),
           M::add(M::add(zx, M::mul(1.0f, p.y)), M::add(zz, z1)),
           M::add(M::add(zx, zy), M::add(M::mul(1.0f, p.z), z1)) };
}

template <class M>
__device__ __forceinline__ V3 integrate(const StepArgs& a, float4 P, float4 V) {
  if (a.use_drag) { V.x = M::mul(V.x, a.keep); V.y = M::mul(V.y, a.keep); V.z = M::mul(V.z, a.keep); }
  // fma(dt*dt, force, fma(dt, velocity, position)): fused in every profile (GLSL fma / std::fma)
  return { __fmaf_rn(a.dt2, a.fx, __fmaf_rn(a.dt, V.x, P.x)),
           __fmaf_rn(a.dt2, a.fy, __fmaf_rn(a.dt, V.y, P.y)),
           __fmaf_rn(a.dt2, a.fz, __fmaf_rn(a.dt, V.z, P.z)) };
}

// ------------------------------------------------------------------------------------------------
// Pipelined kernel: K == 8 constraint iterations (the reference's constant), any nverts >= 1.
//
// Two step bodies share the pipeline state:
//  * steady_step  — used for every chunk whose 8 steps have all 8 stages busy (8 <= t, t <= N-1). No stage
//    predicates; the work of a step is arranged in phases so that the 8 independent stage chains sit in ONE
//    basic block and ptxas interleaves them (ILP 8 hides FADD/FMUL/MUFU latency at 12-16 warps per SM):
//      A  8 x FTL projection; one warp-uniform test routes out-of-range normalisations to the IEEE builtins
//      T  8 x sphere test + the deferred test of the vertex being finalised
//      B  only if some lane of the warp touches the sphere: 9 x push-out, applied with selects
//      C  commit registers; the finished vertex t-8 goes into the slot vertex t was read from
//    The step loop is rolled (code of one step ~ 10 KB) so the hot loop stays inside the instruction cache.
//  * edge_step    — fill (t < 8) and drain (t >= N) steps, ragged N, capsules: per-stage warp-uniform
//    predicates, compact code, same arithmetic.
// ------------------------------------------------------------------------------------------------
struct Pipe {
  V3 Xp[kChunk];        // D(i-1, k+1): the already projected previous vertex of stage k
  float4 pass[kChunk];  // xyz = C(i, k): input of stage k ; w = sf * rest_i
  V3 heldD, heldd;      // D(i,K) and d_i of the vertex waiting for d_{i+1}
  V3 rootV;
};

template <class M> __device__ __noinline__ float inversesqrt_slow(float x) { return M::inversesqrt(x); }

// inv[k] = inversesqrt(x[k]). `ok`: every input of this lane lies inside the range of the branch-free
// sequence; the (rare) IEEE-builtin path is taken by the whole warp or not at all.
template <class M, int n>
__device__ __forceinline__ void inversesqrt_batch(const float (&x)[n], float (&inv)[n], bool ok) {
  if (!M::kRangeChecked || __all_sync(0xffffffffu, ok)) {
#pragma unroll
    for (int k = 0; k < n; ++k) inv[k] = M::inversesqrt_in_range(x[k]);
  } else {
#pragma unroll
    for (int k = 0; k < n; ++k) inv[k] = inversesqrt_slow<M>(x[k]);
  }
}

template <class M>
__device__ __forceinline__ void steady_step(const StepArgs& a, Pipe& s, float4* slotP, float4* slotV, float* slotR, bool first) {
  constexpr int K = kChunk;
  const V3 c = { a.cx, a.cy, a.cz };
  const float rest_out = *slotR;
  const V3 rootX = s.Xp[K - 1];                             // still the root when `first` (stage K-1 has not run yet)
  {
    const float4 P = *slotP, V = *slotV;
    *slotR = P.w;
    const V3 x = integrate<M>(a, P, V);
    s.pass[0] = make_float4(x.x, x.y, x.z, M::mul(a.sf, P.w));
  }
  // ---- phase A: projections --------------------------------------------------------------------
  V3 vd[K], D[K];
  float dp[K], inv[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    vd[k] = vsub<M>(V3{ s.pass[k].x, s.pass[k].y, s.pass[k].z }, s.Xp[k]);
    dp[k] = M::dot(vd[k], vd[k]);
  }
  bool ok = true;
  if (M::kRangeChecked) {                                   // NaN slips through min/max and stays NaN on both paths
    float mn = dp[0], mx = dp[0];
#pragma unroll
    for (int k = 1; k < K; ++k) { mn = fminf(mn, dp[k]); mx = fmaxf(mx, dp[k]); }
    ok = M::in_fast_range(mn) && M::in_fast_range(mx);
  }
  inversesqrt_batch<M>(dp, inv, ok);
#pragma unroll
  for (int k = 0; k < K; ++k) D[k] = M::project(s.Xp[k], vd[k], inv[k], s.pass[k].w);
  const V3 dF = vsub<M>(D[K - 1], V3{ s.pass[K - 1].x, s.pass[K - 1].y, s.pass[K - 1].z });   // cs:116
  V3 fp = s.heldD;                                          // the vertex being finalised: t-K
  V3 fw = M::scale(dF, a.damp);                             // cs:119-121

  // ---- phase T: collision tests (entry K-1 is the finalised vertex; stage K-1 itself is tested next step)
  V3 pt[K];
  float dpc[K], invc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    pt[k] = vsub<M>(k < K - 1 ? D[k] : fp, c);
    dpc[k] = M::dot(pt[k], pt[k]);
  }
  float mnc = dpc[0];
#pragma unroll
  for (int k = 1; k < K; ++k) mnc = fminf(mnc, dpc[k]);

  // ---- phase B: push-outs, only when some lane of the warp touches the sphere ------------------
  V3 C[K - 1];
  if (__any_sync(0xffffffffu, mnc < a.r2)) {
    // a hit has dpc < r2 < inf (launcher guarantees), so only the lower bound of the fast range can fail
    inversesqrt_batch<M>(dpc, invc, !(mnc < M::kFastLo));
#pragma unroll
    for (int k = 0; k < K - 1; ++k) {
      const bool hit = dpc[k] < a.r2;
      const V3 q = M::push_out(c, M::scale(pt[k], invc[k]), a.r);
      C[k].x = hit ? q.x : D[k].x; C[k].y = hit ? q.y : D[k].y; C[k].z = hit ? q.z : D[k].z;
    }
    const bool hit = dpc[K - 1] < a.r2;
    const V3 n = M::scale(pt[K - 1], invc[K - 1]);
    const V3 q = M::push_out(c, n, a.r), w = M::reflect(fw, n);
    fp.x = hit ? q.x : fp.x; fp.y = hit ? q.y : fp.y; fp.z = hit ? q.z : fp.z;
    fw.x = hit ? w.x : fw.x; fw.y = hit ? w.y : fw.y; fw.z = hit ? w.z : fw.z;
  } else {
#pragma unroll
    for (int k = 0; k < K - 1; ++k) C[k] = D[k];
  }

  // ---- phase C: commit ---------------------------------------------------------------------------
#pragma unroll
  for (int k = K - 2; k >= 0; --k) s.pass[k + 1] = make_float4(C[k].x, C[k].y, C[k].z, s.pass[k].w);
#pragma unroll
  for (int k = 0; k < K; ++k) s.Xp[k] = D[k];
  float4 oP = make_float4(fp.x, fp.y, fp.z, rest_out), oV = make_float4(fw.x, fw.y, fw.z, 0.f);
  if (first) {                                              // t == K: the slot is vertex 0
    oP = make_float4(rootX.x, rootX.y, rootX.z, rest_out);
    oV = make_float4(s.rootV.x, s.rootV.y, s.rootV.z, 0.f);
  }
  *slotP = oP;
  *slotV = oV;
  s.heldD = D[K - 1]; s.heldd = dF;
}

template <class M, bool CAPS>
__device__ __forceinline__ void edge_step(const StepArgs& a, Pipe& s, int t, int N, float4* slotP, float4* slotV, float* slotR) {
  constexpr int K = kChunk;
  const float rest_out = *slotR;                            // rest length of vertex t-K (valid once t >= K)
  V3 rootX = s.Xp[K - 1];                                   // still the root at t == K
  if (t < N) {
    const float4 P = *slotP, V = *slotV;
    *slotR = P.w;
    if (t == 0) {
      const V3 x0 = root_transform<M>(V3{ P.x, P.y, P.z });
      s.rootV = vsub<M>(x0, V3{ P.x, P.y, P.z });           // p.velocity = p.position - lastPosition (cs:192)
#pragma unroll
      for (int k = 0; k < K; ++k) s.Xp[k] = x0;
      rootX = x0;
      s.heldD = x0;
    } else {
      const V3 x = integrate<M>(a, P, V);
      s.pass[0] = make_float4(x.x, x.y, x.z, M::mul(a.sf, P.w));
    }
  }
  // last iteration (stage K-1) on vertex i = t-(K-1): its collision is deferred one step so that it can
  // also reflect the velocity, which needs d_{i+1}.
  const int iF = t - (K - 1);
  const bool validF = (iF >= 1) && (iF < N);
  V3 DF = s.heldD, dF = s.heldd;
  if (validF) {
    const V3 prev = { s.pass[K - 1].x, s.pass[K - 1].y, s.pass[K - 1].z };
    DF = ftl<M>(s.Xp[K - 1], prev, s.pass[K - 1].w);
    dF = vsub<M>(DF, prev);                                 // s_particles[i].velocity = p1_bis - p1 (cs:116)
    s.Xp[K - 1] = DF;
  }
  if (t >= K) {
    float4 oP, oV;
    if (t == K) {                                           // vertex 0
      oP = make_float4(rootX.x, rootX.y, rootX.z, rest_out);
      oV = make_float4(s.rootV.x, s.rootV.y, s.rootV.z, 0.f);
    } else {                                                // vertex t-K >= 1, held since the previous step
      V3 w = validF ? M::scale(dF, a.damp) : s.heldd;       // cs:119-121; the tip keeps its own d
      V3 p = s.heldD;
      collide_all_pos_vel<M, CAPS>(a, p, w);
      oP = make_float4(p.x, p.y, p.z, rest_out);
      oV = make_float4(w.x, w.y, w.z, 0.f);
    }
    *slotP = oP;
    *slotV = oV;
  }
  s.heldD = DF; s.heldd = dF;
  // stages K-2 .. 0 (descending: stage k consumes pass[k] before stage k-1 overwrites it)
#pragma unroll
  for (int k = K - 2; k >= 0; --k) {
    const int i = t - k;
    if (i >= 1 && i < N) {
      const V3 prev = { s.pass[k].x, s.pass[k].y, s.pass[k].z };
      const V3 D = ftl<M>(s.Xp[k], prev, s.pass[k].w);
      s.Xp[k] = D;
      const V3 C = collide_all_pos<M, CAPS>(a, D);
      s.pass[k + 1] = make_float4(C.x, C.y, C.z, s.pass[k].w);
    }
  }
}

template <class M, bool CAPS, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) hair_step_pipelined_kernel(const __grid_constant__ StepArgs a) {
  constexpr int K = kChunk;
  extern __shared__ float4 smem4[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long long strand0 = ((long long)blockIdx.x * kWarpsPerBlock + warp) * 32;
  if (strand0 >= a.nstrands) return;      // warp-uniform; no block-level barrier below
  const int N = a.nverts;
  const int nvalid = (int)min((long long)32, a.nstrands - strand0);

  float4* sP = smem4 + warp * (kWarpSmemBytes / 16);
  float4* sV = sP + 32 * kPitch;
  float* sR = reinterpret_cast<float*>(sV + 32 * kPitch);
  float4* gP = a.pos + strand0 * N;
  float4* gV = a.vel + strand0 * N;

  // cooperative chunk copy: float4 q = r*32 + lane of the warp's 32x8 tile -> strand q/8, vertex q%8
  const int cl_strand = lane >> 3, cl_vert = lane & 7;
  float4* myP = sP + lane * kPitch;
  float4* myV = sV + lane * kPitch;
  float* myR = sR + lane;

  Pipe s;
#pragma unroll
  for (int k = 0; k < K; ++k) { s.Xp[k] = { 0.f, 0.f, 0.f }; s.pass[k] = make_float4(0.f, 0.f, 0.f, 0.f); }
  s.heldD = s.heldd = s.rootV = { 0.f, 0.f, 0.f };

  const int nsteps = N + K;
  const int npass = (nsteps + kChunk - 1) / kChunk;
  // one 128-byte line per strand per plane per chunk: lane l asks L2 for a later chunk of strand l, so the
  // cooperative load two passes later finds it on chip
  auto prefetch_chunk = [&](int c) {
    if (c * kChunk < N && lane < nvalid) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(gP + (long long)lane * N + c * kChunk));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(gV + (long long)lane * N + c * kChunk));
    }
  };
  prefetch_chunk(1);
  for (int c = 0; c < npass; ++c) {
    prefetch_chunk(c + 2);
    if (c * kChunk < N) {
#pragma unroll
      for (int r = 0; r < kChunk; ++r) {
        const int sl = r * 4 + cl_strand, v = c * kChunk + cl_vert;
        if (sl < nvalid && v < N) {
          sP[sl * kPitch + cl_vert] = __ldcs(gP + (long long)sl * N + v);
          sV[sl * kPitch + cl_vert] = __ldcs(gV + (long long)sl * N + v);
        }
      }
    }
    __syncwarp();

    if (!CAPS && c >= 1 && c * kChunk + (kChunk - 1) <= N - 1) {
#pragma unroll 1
      for (int j = 0; j < kChunk; ++j) steady_step<M>(a, s, myP + j, myV + j, myR + j * 32, c == 1 && j == 0);
    } else {
      const int jend = min(kChunk, nsteps - c * kChunk);
#pragma unroll 1
      for (int j = 0; j < jend; ++j) edge_step<M, CAPS>(a, s, c * kChunk + j, N, myP + j, myV + j, myR + j * 32);
    }
    __syncwarp();

    if (c >= 1) {
#pragma unroll
      for (int r = 0; r < kChunk; ++r) {
        const int sl = r * 4 + cl_strand, v = (c - 1) * kChunk + cl_vert;
        if (sl < nvalid && v < N) {
          __stcs(gP + (long long)sl * N + v, sP[sl * kPitch + cl_vert]);
          __stcs(gV + (long long)sl * N + v, sV[sl * kPitch + cl_vert]);
        }
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// Generic kernel: any iteration count (including 0), thread per strand, iterates in place in
// global memory. Same operation sequence; used when iterations != 8 and as a GPU-side cross-check.
// ------------------------------------------------------------------------------------------------
template <class M, bool CAPS>
__global__ void __launch_bounds__(128) hair_step_generic_kernel(const __grid_constant__ StepArgs a) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= a.nstrands) return;
  const int N = a.nverts, K = a.iterations;
  float4* P = a.pos + s * N;
  float4* V = a.vel + s * N;

  const float4 P0 = P[0];
  const V3 x0 = root_transform<M>(V3{ P0.x, P0.y, P0.z });
  const V3 v0 = vsub<M>(x0, V3{ P0.x, P0.y, P0.z });
  P[0] = make_float4(x0.x, x0.y, x0.z, P0.w);
  V[0] = make_float4(v0.x, v0.y, v0.z, 0.f);
  for (int i = 1; i < N; ++i) {
    const float4 p = P[i], v = V[i];
    const V3 x = integrate<M>(a, p, v);
    P[i] = make_float4(x.x, x.y, x.z, p.w);
    if (K == 0) V[i] = make_float4(v.x, v.y, v.z, 0.f);
  }
  for (int k = 0; k < K; ++k) {
    const bool last = (k == K - 1);
    V3 xp = x0, heldD = x0, heldd = { 0.f, 0.f, 0.f };
    float heldRest = 0.f;
    for (int i = 1; i < N; ++i) {
      const float4 p = P[i];
      const V3 prev = { p.x, p.y, p.z };
      const V3 D = ftl<M>(xp, prev, M::mul(a.sf, p.w));
      xp = D;
      if (!last) {
        const V3 C = collide_all_pos<M, CAPS>(a, D);
        P[i] = make_float4(C.x, C.y, C.z, p.w);
      } else {
        const V3 d = vsub<M>(D, prev);
        if (i >= 2) {
          V3 w = M::scale(d, a.damp), q = heldD;
          collide_all_pos_vel<M, CAPS>(a, q, w);
          P[i - 1] = make_float4(q.x, q.y, q.z, heldRest);
          V[i - 1] = make_float4(w.x, w.y, w.z, 0.f);
        }
        heldD = D; heldd = d; heldRest = p.w;
      }
    }
    if (last && N >= 2) {
      V3 w = heldd, q = heldD;
      collide_all_pos_vel<M, CAPS>(a, q, w);
      P[N - 1] = make_float4(q.x, q.y, q.z, heldRest);
      V[N - 1] = make_float4(w.x, w.y, w.z, 0.f);
    }
  }
}

template <class M, bool CAPS>
cudaError_t launch_t(const StepArgs& a, cudaStream_t stream) {
  if (a.nstrands <= 0 || a.nverts <= 0) return cudaSuccess;
  // steady_step assumes a colliding vertex has |p - c|^2 < r^2 <= 2^64; absurd radii take the generic kernel
  const bool pipelined = a.iterations == kChunk && a.r2 <= 1.8446744073709551616e19f;
  if (pipelined) {
    constexpr int smem = kWarpsPerBlock * kWarpSmemBytes;
    static_assert(smem <= 48 * 1024, "stays under the default dynamic shared-memory limit: no per-device opt-in needed");
    const long long strands_per_block = kWarpsPerBlock * 32;
    const long long blocks = (a.nstrands + strands_per_block - 1) / strands_per_block;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    static const int occ = [] { const char* e = getenv("BH_SCHED_OCC"); return e ? atoi(e) : 3; }();   // tuning knob
    if (occ == 4) hair_step_pipelined_kernel<M, CAPS, 4><<<(unsigned)blocks, kThreads, smem, stream>>>(a);
    else hair_step_pipelined_kernel<M, CAPS, 3><<<(unsigned)blocks, kThreads, smem, stream>>>(a);
  } else {
    const long long blocks = (a.nstrands + 127) / 128;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    hair_step_generic_kernel<M, CAPS><<<(unsigned)blocks, 128, 0, stream>>>(a);
  }
  return cudaGetLastError();
}

}  // namespace

int step_kernel_kind(const StepArgs& a) {
  if (a.prefer_latency && wave_kernel_eligible(a)) return 3;
  if (stream_kernel_eligible(a)) return 0;
  return a.iterations == kChunk && a.r2 <= 1.8446744073709551616e19f ? 1 : 2;
}

cudaError_t launch_step(const StepArgs& a, int math, cudaStream_t stream, unsigned int* tile_counter) {
  if (a.nstrands <= 0 || a.nverts <= 0) return cudaSuccess;
  if (a.prefer_latency && wave_kernel_eligible(a)) return launch_step_wave(a, math, stream);
  if (tile_counter && stream_kernel_eligible(a)) {
    if (a.nverts != 4 || a.nstrands % 2 == 0) return launch_step_stream(a, math, stream, tile_counter);
    // nverts == 4 pairs strands into 128-byte tensor rows: the last strand of an odd count takes the per-strand kernel
    StepArgs head = a, tail = a;
    head.nstrands = a.nstrands - 1;
    tail.nstrands = 1; tail.pos = a.pos + head.nstrands * 4; tail.vel = a.vel + head.nstrands * 4;
    const cudaError_t e = launch_step_stream(head, math, stream, tile_counter);
    return e != cudaSuccess ? e : launch_step(tail, math, stream, nullptr);
  }
  const bool caps = a.ncaps > 0;
  if (math == 0) return caps ? launch_t<MathExact, true>(a, stream) : launch_t<MathExact, false>(a, stream);
  return caps ? launch_t<MathFast, true>(a, stream) : launch_t<MathFast, false>(a, stream);
}

}  // namespace bh
