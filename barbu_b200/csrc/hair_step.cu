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

namespace bh {

namespace {

constexpr int kChunk = 8;                 // vertices per staged chunk == pipeline depth
constexpr int kPitch = kChunk + 1;        // float4 per lane slot row: 144 B -> conflict-free LDS.128/STS.128
constexpr int kWarpsPerBlock = 4;
constexpr int kThreads = kWarpsPerBlock * 32;
// per warp: position slots, velocity slots (32 * kPitch float4 each) + rest-length ring (8 * 32 floats)
constexpr int kWarpSmemBytes = 2 * 32 * kPitch * 16 + kChunk * 32 * 4;

// Closest point on the capsule axis, then the same push-out as the sphere. Extension (no reference).
template <class M>
__device__ __forceinline__ V3 capsule_center(const Capsule& c, V3 p) {
  V3 a = { c.ax, c.ay, c.az };
  const V3 ab = { M::sub(c.bx, c.ax), M::sub(c.by, c.ay), M::sub(c.bz, c.az) };
  const float l2 = M::dot(ab, ab);
  if (l2 > 0.0f) {
    const V3 ap = vsub<M>(p, a);
    float t = __fdiv_rn(M::dot(ap, ab), l2);
    t = fminf(fmaxf(t, 0.0f), 1.0f);
    a = { M::add(c.ax, M::mul(t, ab.x)), M::add(c.ay, M::mul(t, ab.y)), M::add(c.az, M::mul(t, ab.z)) };
  }
  return a;
}

// CollideSphere(+1, c, r, p) on the position only (iterations whose velocity is dead).
template <class M>
__device__ __forceinline__ V3 collide_pos(V3 p, V3 c, float r, float r2) {
  const V3 pt = vsub<M>(p, c);
  const float dp = M::dot(pt, pt);
  if (dp < r2) {
    const V3 n = M::scale(pt, M::inversesqrt(dp));
    p = M::push_out(c, n, r);
  }
  return p;
}
// ... and on position + velocity (last iteration): vel = reflect(vel, n).
template <class M>
__device__ __forceinline__ void collide_pos_vel(V3& p, V3& w, V3 c, float r, float r2) {
  const V3 pt = vsub<M>(p, c);
  const float dp = M::dot(pt, pt);
  if (dp < r2) {
    const V3 n = M::scale(pt, M::inversesqrt(dp));
    p = M::push_out(c, n, r);
    w = M::reflect(w, n);
  }
}

template <class M, bool CAPS>
__device__ __forceinline__ V3 collide_all_pos(const StepArgs& a, V3 p) {
  p = collide_pos<M>(p, V3{ a.cx, a.cy, a.cz }, a.r, a.r2);
  if (CAPS) {
    for (int q = 0; q < a.ncaps; ++q) {
      const float r = a.caps[q].r;
      p = collide_pos<M>(p, capsule_center<M>(a.caps[q], p), r, M::mul(r, r));
    }
  }
  return p;
}
template <class M, bool CAPS>
__device__ __forceinline__ void collide_all_pos_vel(const StepArgs& a, V3& p, V3& w) {
  collide_pos_vel<M>(p, w, V3{ a.cx, a.cy, a.cz }, a.r, a.r2);
  if (CAPS) {
    for (int q = 0; q < a.ncaps; ++q) {
      const float r = a.caps[q].r;
      collide_pos_vel<M>(p, w, capsule_center<M>(a.caps[q], p), r, M::mul(r, r));
    }
  }
}

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
// ------------------------------------------------------------------------------------------------
template <class M, bool CAPS>
__global__ void __launch_bounds__(kThreads, 4) hair_step_pipelined_kernel(const __grid_constant__ StepArgs a) {
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

  V3 Xp[K];        // D(i-1, k+1): the already projected previous vertex of stage k
  float4 pass[K];  // xyz = C(i, k): input of stage k ; w = sf * rest_i
  V3 heldD = { 0.f, 0.f, 0.f }, heldd = { 0.f, 0.f, 0.f };   // D(i,K) and d_i of the vertex awaiting d_{i+1}
  V3 rootV = { 0.f, 0.f, 0.f };
#pragma unroll
  for (int k = 0; k < K; ++k) { Xp[k] = { 0.f, 0.f, 0.f }; pass[k] = make_float4(0.f, 0.f, 0.f, 0.f); }

  const int nsteps = N + K;
  const int npass = (nsteps + kChunk - 1) / kChunk;
  for (int c = 0; c < npass; ++c) {
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

    const int jend = min(kChunk, nsteps - c * kChunk);
#pragma unroll 1
    for (int j = 0; j < jend; ++j) {
      const int t = c * kChunk + j;
      const float rest_out = myR[j * 32];             // rest length of vertex t-K (valid once t >= K)
      V3 rootX = Xp[K - 1];                           // still the root at t == K (stage K-1 first runs there)

      if (t < N) {
        const float4 P = myP[j], V = myV[j];
        myR[j * 32] = P.w;
        if (t == 0) {
          const V3 x0 = root_transform<M>(V3{ P.x, P.y, P.z });
          rootV = vsub<M>(x0, V3{ P.x, P.y, P.z });   // p.velocity = p.position - lastPosition (cs:192)
#pragma unroll
          for (int k = 0; k < K; ++k) Xp[k] = x0;
          rootX = x0;
        } else {
          const V3 x = integrate<M>(a, P, V);
          pass[0] = make_float4(x.x, x.y, x.z, M::mul(a.sf, P.w));
        }
      }

      // last iteration (stage K-1) on vertex i = t-(K-1): no collision yet, it is deferred one step
      // so that it can also reflect the velocity, which needs d_{i+1}.
      const int iF = t - (K - 1);
      const bool validF = (iF >= 1) && (iF < N);
      V3 DF = heldD, dF = heldd;
      if (validF) {
        const V3 prev = { pass[K - 1].x, pass[K - 1].y, pass[K - 1].z };
        DF = ftl<M>(Xp[K - 1], prev, pass[K - 1].w);
        dF = vsub<M>(DF, prev);                       // s_particles[i].velocity = p1_bis - p1 (cs:116)
        Xp[K - 1] = DF;
      }
      if (t >= K) {
        float4 oP, oV;
        if (t == K) {                                 // vertex 0
          oP = make_float4(rootX.x, rootX.y, rootX.z, rest_out);
          oV = make_float4(rootV.x, rootV.y, rootV.z, 0.f);
        } else {                                      // vertex t-K >= 1, held since the previous step
          V3 w = validF ? M::scale(dF, a.damp) : heldd;   // cs:119-121; the tip keeps its own d
          V3 p = heldD;
          collide_all_pos_vel<M, CAPS>(a, p, w);
          oP = make_float4(p.x, p.y, p.z, rest_out);
          oV = make_float4(w.x, w.y, w.z, 0.f);
        }
        myP[j] = oP;
        myV[j] = oV;
      }
      heldD = DF; heldd = dF;

      // stages K-2 .. 0 (descending: stage k consumes pass[k] before stage k-1 overwrites it)
#pragma unroll
      for (int k = K - 2; k >= 0; --k) {
        const int i = t - k;
        if (i >= 1 && i < N) {
          const V3 prev = { pass[k].x, pass[k].y, pass[k].z };
          const V3 D = ftl<M>(Xp[k], prev, pass[k].w);
          Xp[k] = D;
          const V3 C = collide_all_pos<M, CAPS>(a, D);
          pass[k + 1] = make_float4(C.x, C.y, C.z, pass[k].w);
        }
      }
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
  if (a.iterations == kChunk) {
    constexpr int smem = kWarpsPerBlock * kWarpSmemBytes;
    static_assert(smem <= 48 * 1024, "stays under the default dynamic shared-memory limit: no per-device opt-in needed");
    const long long strands_per_block = kWarpsPerBlock * 32;
    const long long blocks = (a.nstrands + strands_per_block - 1) / strands_per_block;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    hair_step_pipelined_kernel<M, CAPS><<<(unsigned)blocks, kThreads, smem, stream>>>(a);
  } else {
    const long long blocks = (a.nstrands + 127) / 128;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    hair_step_generic_kernel<M, CAPS><<<(unsigned)blocks, 128, 0, stream>>>(a);
  }
  return cudaGetLastError();
}

}  // namespace

int step_kernel_kind(int /*nverts*/, int iterations, int /*ncaps*/) { return iterations == kChunk ? 0 : 1; }

cudaError_t launch_step(const StepArgs& a, int math, cudaStream_t stream) {
  const bool caps = a.ncaps > 0;
  if (math == 0) return caps ? launch_t<MathExact, true>(a, stream) : launch_t<MathExact, false>(a, stream);
  return caps ? launch_t<MathFast, true>(a, stream) : launch_t<MathFast, false>(a, stream);
}

}  // namespace bh
