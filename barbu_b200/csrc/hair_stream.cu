// hair_stream.cu — the streaming step kernel for sm_100a (the hot path: nverts % 8 == 0 or the reference's nverts = 4,
// 8 constraint iterations, sphere collider and up to 8 capsules).
//
// Same arithmetic as hair_step.cu (one dispatch of cs_simulation.glsl:170-208 + PingPongBuffer::swap,
// written back in place), organised for Blackwell:
//
//  * TMA tiles. A warp owns TILES of 32 consecutive strands. A tile moves between HBM and shared memory
//    in chunks of 8 vertices: one cp.async.bulk.tensor.2d box of 32 rows x 128 B per plane, written
//    with the 128-byte swizzle, so that lane l finds vertex j of ITS strand at 16-byte column
//    j ^ (l & 7) of row l — conflict-free LDS.128/STS.128 with no transposing copy loop. Two chunk
//    buffers per warp form a ring on two mbarriers; a finished chunk leaves through a TMA tile store
//    from the same buffer (the result of vertex t-8 is written into the slot vertex t was read from).
//  * Persistent warps. The grid is (#SMs x resident blocks); every warp pulls tiles from an atomic
//    counter and runs ONE software pipeline across all of them: at step t stage k applies constraint
//    iteration k+1 to stream vertex t-k. A strand root entering the stream passes through the stages
//    unchanged and resets their "previous vertex", so the 8-step fill/drain is paid once per warp
//    instead of once per strand, and every step of the launch is the same steady-state body.
//  * Packed FP32x2. The eight stage chains of a thread are mutually independent, so stages (a, a+4)
//    share one 64-bit register pair per coordinate and every FADD/FMUL/FFMA of the projection and of
//    the collision test issues as FADD2/FMUL2/FFMA2 — half the issue slots at the same FP32 lane
//    throughput (measured: tools/ubench/f32x2.cu). Packed add/mul/fma round exactly like the scalar
//    .rn forms; in the exact profile a product is written fma(a, b, -0.0) with the -0.0 taken from a
//    kernel parameter, because ptxas contracts `mul.rn.f32x2` + `add.rn.f32x2` into FFMA2 even with
//    explicit rounding modifiers and --fmad=false.
//  * The instruction cache decides what may be inlined: the hot loop (two chunk bodies x two input variants + their
//    push-out blocks) is ~2,000 instructions against an L1.5 I-cache of 32 KB, so every rare path — the IEEE fallback of
//    the inverse square root, the exact capsule chain — is ONE out-of-line call, and the capsule variant runs a single
//    step variant (every step hands its positions on through P) instead of a free and a contact one.
//  * Tile order alternates between consecutive launches (StepArgs::reverse, set by bh_step): a launch starts where the
//    previous one ended, in the part of the state that is still in L2.
//
// Algorithmic traffic: 16 B pos + 16 B vel read and the same written = 64 B per vertex per launch.
#include "hair_step.cuh"
#include "hair_math.cuh"
#include "hair_collide.cuh"

#include <cuda.h>

#include <cstdint>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <mutex>

namespace bh {

namespace {

typedef unsigned long long u64;

#ifndef BH_CAPS_ONE_BODY
#define BH_CAPS_ONE_BODY 0
#endif
#ifndef BH_CAPS_SINGLE
#define BH_CAPS_SINGLE 2
#endif
#ifndef BH_CAPS_PAIRS
#define BH_CAPS_PAIRS 1
#endif
constexpr int kK = 8;                            // constraint iterations == pipeline depth == vertices per chunk
// Warps per block is a LAUNCH parameter (the kernel reads blockDim): small shards get blocks of kWarps warps, three to an SM, so
// that their tiles spread over the SMs; a shard that fills the GPU anyway gets ONE block of kWarpsBig warps per SM — the same
// twelve resident warps, 1 % faster in the exact profile (0.4903 -> 0.4860 ms per launch at configs[1], fused frame 1.922 ->
// 1.900 ms; A/B on one box). The HBM-bound fast launch keeps the small blocks (0.3398 vs 0.3409 ms).
#ifndef BH_STREAM_WARPS
#define BH_STREAM_WARPS 4
#endif
#ifndef BH_STREAM_WARPS_BIG
#define BH_STREAM_WARPS_BIG 12
#endif
constexpr int kWarps = BH_STREAM_WARPS;
constexpr int kWarpsBig = BH_STREAM_WARPS_BIG;
constexpr int kMaxThreads = (kWarpsBig > kWarps ? kWarpsBig : kWarps) * 32;
constexpr int kPlaneTile = 32 * 128;             // bytes of one plane of one chunk: 32 strands x 8 vertices x 16 B
constexpr int kStageBytes = 2 * kPlaneTile;      // position + velocity
constexpr int kWarpTileBytes = 2 * kStageBytes;  // two stages
constexpr int kRingBytes = kK * 32 * 4;          // rest lengths of the 8 vertices in flight, [slot][lane]
constexpr int smem_bytes(int warps) { return 1024 /* alignment slack */ + warps * (kWarpTileBytes + kRingBytes) + warps * 2 * 8; }

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "BH_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra BH_DONE;\n"
      "bra BH_WAIT;\n"
      "BH_DONE:\n"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_tile(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_tile(const CUtensorMap* map, int c0, int c1, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(map), "r"(c0), "r"(c1), "r"(src) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// all bulk groups but the most recent one are COMPLETE (their global writes performed), not merely read out of shared memory
__device__ __forceinline__ void tma_wait_complete_but1() { asm volatile("cp.async.bulk.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- packed pairs ------------------------------------------------------------------------------
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ float lo(u64 v) { float l, h; asm("mov.b64 {%0, %1}, %2;" : "=f"(l), "=f"(h) : "l"(v)); (void)h; return l; }
__device__ __forceinline__ float hi(u64 v) { float l, h; asm("mov.b64 {%0, %1}, %2;" : "=f"(l), "=f"(h) : "l"(v)); (void)l; return h; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 mul2_contractable(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
// written as two float negations so that ptxas folds it into the consumer's operand modifier (FFMA2 -R); an integer xor
// of the sign bits costs two LOP3 per use
__device__ __forceinline__ u64 neg2(u64 a) { return pk(-lo(a), -hi(a)); }
__device__ __forceinline__ float rsq_approx(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

constexpr u64 kOne2 = 0x3f8000003f800000ull;
constexpr u64 kHalf2 = 0x3f0000003f000000ull;

#ifdef BH_STATS   // debug build only (tools/stats_build.sh): contact statistics of the step loop, printed by the last warp of a launch
__device__ unsigned long long g_stats[16];
__device__ unsigned long long g_span[4];       // first start / last end of a warp (globaltimer ns), sum of the warps' busy ns, warps
__device__ __forceinline__ unsigned long long* bh_warp_stats() {           // per-warp counters in shared memory, flushed once per launch
  __shared__ unsigned long long w[16][16];
  return w[threadIdx.x >> 5];
}
#define BH_STAT(i, v) (bh_warp_stats()[i] += (unsigned long long)(v))
#define BH_T0(name) const long long name = clock64()
#define BH_T1(i, name) do { if ((threadIdx.x & 31) == 0) BH_STAT(i, clock64() - name); } while (0)
#else
#define BH_T0(name)
#define BH_T1(i, name)
#endif

struct V3p { u64 x, y, z; };                     // lo half: stage a, hi half: stage a + 4

// Packed counterparts of MathExact / MathFast (hair_math.cuh). `nz` is (-0.0f, -0.0f) from a kernel parameter.
struct PackedExact {
  typedef MathExact S;
  static constexpr bool kRangeChecked = true;
  static constexpr bool kTwoStageTest = true;      // stream_step phase T: a contracted sum of squares decides "nobody is near the sphere"
  static __device__ __forceinline__ u64 mul(u64 a, u64 b, u64 nz) { return fma2(a, b, nz); }   // RN(a*b + -0) == RN(a*b), sign of zero included
  static __device__ __forceinline__ u64 dot(V3p a, V3p b, u64 nz) { return add2(add2(mul(a.x, b.x, nz), mul(a.y, b.y, nz)), mul(a.z, b.z, nz)); }
  // -(1 / sqrt(x)) for both halves, finite x >= 2^-102: the sequence of MathExact::inversesqrt_in_range, with the
  // reciprocal seeded by rcp(-s) so that no packed negation is needed afterwards (RN is sign-symmetric).
  static __device__ __forceinline__ u64 neg_inversesqrt_in_range(u64 x, u64 nz) {
    const u64 y = pk(rsq_approx(lo(x)), rsq_approx(hi(x)));
    const u64 s0 = mul(x, y, nz);
    const u64 h = mul(y, kHalf2, nz);
    const u64 s = fma2(fma2(neg2(s0), s0, x), h, s0);                       // sqrt_rn(x)
    const u64 rn = pk(rcp_approx(-lo(s)), rcp_approx(-hi(s)));              // -r
    return fma2(rn, fma2(rn, s, kOne2), rn);                                // -(r + r*(1 - r*s)) = -rcp_rn(s)
  }
  static __device__ __forceinline__ u64 neg_inversesqrt_ieee(u64 x) {
    return pk(-__frcp_rn(__fsqrt_rn(lo(x))), -__frcp_rn(__fsqrt_rn(hi(x))));
  }
  static __device__ __forceinline__ float inv_of(float ninv) { return -ninv; }
  // p0 + L * (vd * inv) given ninv = -inv:  p0 - L * (vd * ninv), bit-identical
  static __device__ __forceinline__ V3p project(V3p p0, V3p vd, u64 ninv, u64 L, u64 nz) {
    return { sub2(p0.x, mul(L, mul(vd.x, ninv, nz), nz)), sub2(p0.y, mul(L, mul(vd.y, ninv, nz), nz)), sub2(p0.z, mul(L, mul(vd.z, ninv, nz), nz)) };
  }
  // c + r * (pt * inv) given ninv
  static __device__ __forceinline__ V3p push_out(V3p c, V3p pt, u64 ninv, u64 r, u64 nz) {
    return { sub2(c.x, mul(r, mul(pt.x, ninv, nz), nz)), sub2(c.y, mul(r, mul(pt.y, ninv, nz), nz)), sub2(c.z, mul(r, mul(pt.z, ninv, nz), nz)) };
  }
  // The same for a sphere centred on (+0, +0, +0): 0 + RN(r * n) as ONE fused operation, RN(r * n + (+0)). For a product
  // that does not round to zero the two agree trivially (adding zero does not move a rounding), and an exactly zero product
  // gives +0 both ways. They would differ — by the sign of a zero — only for a non-zero product that UNDERFLOWS to zero,
  // and a vertex that is pushed out cannot produce one: it lies inside the sphere, so r / |pt| > 1 - 2^-21 and
  // |r * n.x| >= |pt.x| (1 - 2^-21) up to one rounding of n.x, which keeps every non-zero coordinate (the smallest
  // subnormal included: RN(k d / |pt|) >= 1 d whenever r * that can matter) away from the underflow threshold; the lanes that
  // are NOT pushed out discard this value. Saves the three packed additions per stage pair.
  static __device__ __forceinline__ V3p push_out_origin(V3p pt, u64 ninv, u64 r, u64 nz) {
    const u64 pz = 0ull;                                                      // (+0.0f, +0.0f)
    return { fma2(r, neg2(mul(pt.x, ninv, nz)), pz), fma2(r, neg2(mul(pt.y, ninv, nz)), pz), fma2(r, neg2(mul(pt.z, ninv, nz)), pz) };
  }
};

struct PackedFast {
  typedef MathFast S;
  static constexpr bool kRangeChecked = false;
  static constexpr bool kTwoStageTest = false;     // its dot IS the contracted sum
  static __device__ __forceinline__ u64 mul(u64 a, u64 b, u64) { return mul2_contractable(a, b); }
  static __device__ __forceinline__ u64 dot(V3p a, V3p b, u64) { return fma2(a.z, b.z, fma2(a.y, b.y, mul2_contractable(a.x, b.x))); }
  // this profile keeps the inverse square root with its own sign (one MUFU.RSQ per half)
  static __device__ __forceinline__ u64 neg_inversesqrt_in_range(u64 x, u64) { return pk(rsq_approx(lo(x)), rsq_approx(hi(x))); }
  static __device__ __forceinline__ u64 neg_inversesqrt_ieee(u64 x) { return neg_inversesqrt_in_range(x, 0); }
  static __device__ __forceinline__ float inv_of(float isq) { return isq; }
  static __device__ __forceinline__ V3p project(V3p p0, V3p vd, u64 isq, u64 L, u64) {
    const u64 s = mul2_contractable(L, isq);                                // L * inv
    return { fma2(vd.x, s, p0.x), fma2(vd.y, s, p0.y), fma2(vd.z, s, p0.z) };
  }
  static __device__ __forceinline__ V3p push_out(V3p c, V3p pt, u64 isq, u64 r, u64) {
    const u64 s = mul2_contractable(r, isq);
    return { fma2(pt.x, s, c.x), fma2(pt.y, s, c.y), fma2(pt.z, s, c.z) };
  }
  static __device__ __forceinline__ V3p push_out_origin(V3p pt, u64 isq, u64 r, u64 nz) { return push_out(V3p{ 0ull, 0ull, 0ull }, pt, isq, r, nz); }
};

__device__ __forceinline__ V3p sub3(V3p a, V3p b) { return { sub2(a.x, b.x), sub2(a.y, b.y), sub2(a.z, b.z) }; }
__device__ __forceinline__ V3 lo3(V3p v) { return { lo(v.x), lo(v.y), lo(v.z) }; }
__device__ __forceinline__ V3 hi3(V3p v) { return { hi(v.x), hi(v.y), hi(v.z) }; }
__device__ __forceinline__ V3p pk3(V3 l, V3 h) { return { pk(l.x, h.x), pk(l.y, h.y), pk(l.z, h.z) }; }

// (d < r) ? t : f per component, as one FSETP and three FSEL: left to itself ptxas writes such a select, when its result goes
// to a register of its own, as two predicated moves per component
__device__ __forceinline__ V3 sel3_lt(float d, float r, V3 t, V3 f) {
  V3 o;
  asm("{\n.reg .pred p;\nsetp.lt.f32 p, %9, %10;\nselp.f32 %0, %3, %6, p;\nselp.f32 %1, %4, %7, p;\nselp.f32 %2, %5, %8, p;\n}"
      : "=f"(o.x), "=f"(o.y), "=f"(o.z) : "f"(t.x), "f"(t.y), "f"(t.z), "f"(f.x), "f"(f.y), "f"(f.z), "f"(d), "f"(r));
  return o;
}

// vec3(mat4(1.0) * vec4(p, 1.0)) in GLM's operation order (see root_transform in hair_step.cu).
template <class M>
__device__ __forceinline__ V3 root_transform(V3 p) {
  const float zx = M::mul(0.0f, p.x), zy = M::mul(0.0f, p.y), zz = M::mul(0.0f, p.z);
  const float z1 = M::mul(0.0f, 1.0f);
  return { M::add(M::add(M::mul(1.0f, p.x), zy), M::add(zz, z1)),
           M::add(M::add(zx, M::mul(1.0f, p.y)), M::add(zz, z1)),
           M::add(M::add(zx, zy), M::add(M::mul(1.0f, p.z), z1)) };
}

// Capsule variant: ONE step variant (every step hands its positions on through P, as a step after a push-out does) instead of
// a free and a contact variant. 1: exact profile only, 2: both profiles. See stream_chunk.
template <class PM> __device__ __forceinline__ constexpr bool caps_single() { return BH_CAPS_SINGLE == 2 || (BH_CAPS_SINGLE == 1 && PM::kRangeChecked); }

// Pipeline registers. Pair a holds stages a (lo) and a + 4 (hi).
//
// X[k] is D(., k+1) as stage k left it in the previous step. The vertex now in stage k was in stage k-1 then, so its
// input C(i, k) equals X[k-1] of the previous step unless the collider moved it. Collisions are rare per warp-step, so
// the common step (SEP = false) reads its inputs straight from X shifted by one stage — nothing is copied from step to
// step — and only a step that follows a push-out (SEP = true) reads the separately kept C values in P.
struct Pipe {
  V3p X[4];        // D of each stage, previous step; hi half of X[3] is D(i, 8) of the vertex waiting for d_{i+1}
  V3p P[4];        // C(i, k): input of each stage; valid only after a step with a push-out
  u64 L[4];        // sf * rest_i of the vertex in each stage
  V3 heldd;        // d_i of the vertex waiting for d_{i+1}
  V3 rootV[2];     // velocity to write for the root(s) in flight: one per 8 slots, two when a chunk holds two strands (N = 4)
  // collision of the vertex waiting for d_{i+1} (its position is final, its velocity is reflected one step later);
  // written by a push-out step, read by the SEP step that follows it
  V3 heldC, heldN;
  bool heldHit;
  bool heldCap;    // capsule variant, warp-uniform: that vertex went through the capsule chain -> recompute it with its velocity
  // capsule variant, temporal bound (per lane): every position this lane has in flight is at least `slack` away from every
  // capsule's widened surface; <= 0 (or NaN): nothing is known. `lmax`: largest |sf * rest| this lane has seen.
  float slack, lmax;
};

struct Quad { u64 v[4]; };
template <class PM> __device__ __noinline__ Quad neg_inversesqrt_slow(Quad x) {
#pragma unroll 1
  for (int a = 0; a < 4; ++a) x.v[a] = PM::neg_inversesqrt_ieee(x.v[a]);
  return x;
}

// ninv[a] = -inversesqrt(x[a]). The branch-free sequence is issued unconditionally; when some lane of the warp has an
// operand outside its range (`ok` false: rare), the whole warp recomputes with the IEEE builtins.
template <class PM>
__device__ __forceinline__ bool neg_inversesqrt_batch(const u64 (&x)[4], u64 (&ninv)[4], bool ok, u64 nz) {
#pragma unroll
  for (int a = 0; a < 4; ++a) ninv[a] = PM::neg_inversesqrt_in_range(x[a], nz);
  if (PM::kRangeChecked && !__all_sync(0xffffffffu, ok)) {
    const Quad r = neg_inversesqrt_slow<PM>(Quad{ { x[0], x[1], x[2], x[3] } });
#pragma unroll
    for (int a = 0; a < 4; ++a) ninv[a] = r.v[a];
    return true;
  }
  return false;                                                               // warp-uniform: whether the IEEE path ran
}

__device__ __forceinline__ float min8(const u64 (&v)[4]) {
  return fminf(fminf(fminf(lo(v[0]), hi(v[0])), fminf(lo(v[1]), hi(v[1]))), fminf(fminf(lo(v[2]), hi(v[2])), fminf(lo(v[3]), hi(v[3]))));
}
__device__ __forceinline__ float max8(const u64 (&v)[4]) {
  return fmaxf(fmaxf(fmaxf(lo(v[0]), hi(v[0])), fmaxf(lo(v[1]), hi(v[1]))), fmaxf(fmaxf(lo(v[2]), hi(v[2])), fmaxf(lo(v[3]), hi(v[3]))));
}

// One step of the stream pipeline: slot j of the current chunk. Returns whether any lane was pushed out of the
// collider (warp-uniform), i.e. whether the next step must be the SEP variant.
//   RS   : root stride of this chunk. 0: no root in it (an inner chunk of a strand); 8: vertex 0 of a strand is in slot 0,
//          so at step j stage j holds a root, and step a.tip_step finalises the tip of the previous strand (step 7 when
//          nverts % 8 == 0; earlier when the previous strand's last chunk was ragged, its empty slots being NaN);
//          4: the chunk holds two whole strands of the reference's own nverts = 4 (interop.h:8), roots in slots 0 and 4.
//          Stage k holds at step j the vertex of slot j - k (mod 8; negative: of the previous chunk).
//   SEP  : the previous step had a push-out; inputs come from s.P.
//   fin_root (runtime, warp-uniform): the vertex finalised by this step is a root.
template <int RS>
__device__ __forceinline__ bool root_in_stage(const int j, const int stage) {
  return RS == 8 ? j == stage : (RS != 0 && ((j - stage) & (RS - 1)) == 0);  // j, stage in 0..7; j >= 8 (RS == 8): no root in this chunk
}

//   CAPS : capsule colliders are present (extension, a.ncaps > 0). Every step tests the eight new positions against the
//          capsules' bounding spheres (packed, conservative); only a warp that may touch one runs the exact capsule
//          arithmetic (scalar, hair_collide.cuh). The vertex leaving the pipeline then recomputes its whole collision
//          chain with its final velocity one step later, so no collision normals are carried.
// The exact capsule arithmetic runs for few warps and must not bloat the step body (the instruction cache holds the hot
// loop only if the rare paths stay out of line): one out-of-line copy each, one call per step.
// The capsule chain of collide_all_pos (hair_collide.cuh) for the eight positions of a step, branch-free per lane: pair by
// pair, for every capsule in turn the closest point on its axis (IEEE division, scalar), then the sphere push-out in the
// packed form the step uses for the sphere itself. Operation for operation the scalar chain — a lane's arithmetic does not
// depend on what other lanes do; the votes only skip work whose result every lane would discard. The positions travel
// through local memory and one pair is in registers at a time: the function must fit into the registers the step leaves
// free, or the step itself starts spilling (all eight in registers: 280 bytes of spills in the hot loop, capsules nothing
// can reach 2.36 -> 3.2 ms). `skip`: positions that do not collide (roots; the vertex in stage 7, which the next step
// recomputes). `cmask`: capsules whose bound some lane passed on the positions as they came in — a capsule outside it is
// visited only once an earlier one has moved something in that pair.
template <class PM> __device__ __noinline__ void caps_packed(const StepArgs& a, V3p* x, const unsigned skip, const unsigned cmask, const u64 nz) {
  typedef typename PM::S M;
  constexpr int W = BH_CAPS_PAIRS;                                            // pairs in registers at a time (independent chains)
  const float inf = __int_as_float(0x7f800000);
#pragma unroll 1
  for (int q0 = 0; q0 < 4; q0 += W) {
    V3p p[W];
#pragma unroll
    for (int i = 0; i < W; ++i) p[i] = x[q0 + i];
    bool moved = false;
#pragma unroll 1
    for (int k = 0; k < a.ncaps; ++k) {
      if (!moved && !((cmask >> k) & 1u)) continue;
      const Capsule& c = a.caps[k];
      const V3 ab = { a.capx[k][0], a.capx[k][1], a.capx[k][2] };               // b - a, |b - a|^2 and r * r as the scalar chain forms them
      const float l2 = a.capx[k][3];
      const float r = c.r, r2 = a.capx[k][4];
      const V3p A2 = { pk(c.ax, c.ax), pk(c.ay, c.ay), pk(c.az, c.az) };
      V3p ctr[W], pt[W];
      u64 dp[W];
      float mn = inf;
#pragma unroll
      for (int i = 0; i < W; ++i) {
        ctr[i] = A2;
        if (l2 > 0.0f) {                                                      // capsule_center(); warp-uniform
          const V3p AB2 = { pk(ab.x, ab.x), pk(ab.y, ab.y), pk(ab.z, ab.z) };
          const V3p ap = sub3(p[i], A2);
          const u64 d = PM::dot(ap, AB2, nz);
          const float tl = fminf(fmaxf(__fdiv_rn(lo(d), l2), 0.0f), 1.0f), th = fminf(fmaxf(__fdiv_rn(hi(d), l2), 0.0f), 1.0f);
          const u64 t2 = pk(tl, th);
          ctr[i] = { add2(A2.x, PM::mul(t2, AB2.x, nz)), add2(A2.y, PM::mul(t2, AB2.y, nz)), add2(A2.z, PM::mul(t2, AB2.z, nz)) };
        }
        pt[i] = sub3(p[i], ctr[i]);                                           // collide_pos()
        dp[i] = PM::dot(pt[i], pt[i], nz);
        if ((skip >> (q0 + i)) & 1u) dp[i] = pk(inf, hi(dp[i]));              // +inf: below no radius
        if ((skip >> (q0 + i + 4)) & 1u) dp[i] = pk(lo(dp[i]), inf);
        mn = fminf(mn, fminf(lo(dp[i]), hi(dp[i])));                          // a NaN takes no part: it is not below r2 either
      }
      if (!__any_sync(0xffffffffu, mn < r2)) continue;
      moved = true;
      const bool ieee = PM::kRangeChecked && __any_sync(0xffffffffu, mn < M::kFastLo);   // a hit has dp < r2: only the lower end of the range can fail
#pragma unroll
      for (int i = 0; i < W; ++i) {
        u64 ninv = PM::neg_inversesqrt_in_range(dp[i], nz);
        if (ieee) ninv = PM::neg_inversesqrt_ieee(dp[i]);
        const V3p Q = PM::push_out(ctr[i], pt[i], ninv, pk(r, r), nz);
        p[i] = pk3(sel3_lt(lo(dp[i]), r2, lo3(Q), lo3(p[i])), sel3_lt(hi(dp[i]), r2, hi3(Q), hi3(p[i])));
      }
    }
    if (moved) {
#pragma unroll
      for (int i = 0; i < W; ++i) x[q0 + i] = p[i];
    }
  }
}
struct PosVel { V3 p, w; };
template <class M> __device__ __noinline__ PosVel all_pos_vel_slow(const StepArgs& a, V3 p, V3 w) {
  collide_all_pos_vel_bounded<M>(a, p, w);
  return { p, w };
}

__device__ __forceinline__ float sqrt_approx(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// Capsule-shaped conservative test of eight positions: squared distance to each capsule's axis in fast packed arithmetic
// against the radius with a margin (fill_capsule_bounds). It only decides whether the exact chain is entered. Returns whether
// some lane may touch a capsule, and in `slack` a lower bound on the distance of its eight positions to the widened surfaces
// (before the error term the caller subtracts): the smallest sqrt(d^2) (1 - 2^-10) - r_tight. A NaN position takes no part in
// either (fminf skips it: it cannot collide, and stays NaN until the next root resets the bound).
// Returns (warp-uniform) the set of capsules whose bound some lane passed.
__device__ __forceinline__ unsigned caps_tight_touch(const StepArgs& a, const V3p (&X)[4], float& slack) {
  unsigned touch = 0u;
  float sl = __int_as_float(0x7f800000);
#pragma unroll 1
  for (int k = 0; k < a.ncaps; ++k) {
    const Capsule& c = a.caps[k];
    const V3p A = { pk(c.ax, c.ax), pk(c.ay, c.ay), pk(c.az, c.az) };
    const V3p AB = { pk(a.capt[k][0], a.capt[k][0]), pk(a.capt[k][1], a.capt[k][1]), pk(a.capt[k][2], a.capt[k][2]) };
    const float inv_l2 = a.capt[k][3];
    u64 dd[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const V3p ap = sub3(X[q], A);
      const u64 d = fma2(ap.z, AB.z, fma2(ap.y, AB.y, mul2_contractable(ap.x, AB.x)));
      const u64 nt = pk(-__saturatef(lo(d) * inv_l2), -__saturatef(hi(d) * inv_l2));   // -clamp(t, 0, 1); NaN -> 0
      const V3p e = { fma2(nt, AB.x, ap.x), fma2(nt, AB.y, ap.y), fma2(nt, AB.z, ap.z) };
      dd[q] = fma2(e.z, e.z, fma2(e.y, e.y, mul2_contractable(e.x, e.x)));
    }
    const float m = min8(dd);
    if (__any_sync(0xffffffffu, m < a.capt[k][4])) touch |= 1u << k;
    sl = fminf(sl, __fmaf_rn(sqrt_approx(m), 0.9990234375f, -a.capt[k][5]));
  }
  slack = sl;
  return touch;
}

template <class PM, bool ORIGIN, int RS, bool SEP, bool CAPS>
__device__ __forceinline__ bool stream_step(const StepArgs& a, const u64 nz, Pipe& s, const int j, const bool fin_root,
                                            float4* slotP, float4* slotV, float* slotR) {
  typedef typename PM::S M;
  const float rest_out = *slotR;
  const bool set1 = RS == 4 && (j & 4);                                     // which root-velocity set slot j belongs to
  const V3 rootV_out = set1 ? s.rootV[1] : s.rootV[0];
  V3 x;
  float lx;
  {
    const float4 Pin = *slotP, Vin = *slotV;
    *slotR = Pin.w;
    if (root_in_stage<RS>(j, 0)) {
      x = root_transform<M>(V3{ Pin.x, Pin.y, Pin.z });
      const V3 rv = vsub<M>(x, V3{ Pin.x, Pin.y, Pin.z });                  // p.velocity = p.position - lastPosition (cs:192)
      if (set1) s.rootV[1] = rv; else s.rootV[0] = rv;
    } else {
      float4 V = Vin;
      if (a.use_drag) { V.x = M::mul(V.x, a.keep); V.y = M::mul(V.y, a.keep); V.z = M::mul(V.z, a.keep); }
      x = { __fmaf_rn(a.dt2, a.fx, __fmaf_rn(a.dt, V.x, Pin.x)), __fmaf_rn(a.dt2, a.fy, __fmaf_rn(a.dt, V.y, Pin.y)),
            __fmaf_rn(a.dt2, a.fz, __fmaf_rn(a.dt, V.z, Pin.z)) };         // cs:181-182
    }
    lx = M::mul(a.sf, Pin.w);
    if (CAPS) {
      s.lmax = fmaxf(s.lmax, fabsf(lx));                                      // a NaN (out-of-bounds slot) is skipped
      // a root starts a new chain of positions: nothing is known about it until the next test
      if (root_in_stage<RS>(j, 0)) s.slack = __int_as_float(0xff800000);
    }
  }
  // ---- the vertex finalised by this step (t-8): position known since the previous step ------------
  const V3 fin = hi3(s.X[3]);                                               // D(i, 8)
  V3 fp = fin;
  const bool recompute = SEP && CAPS && s.heldCap;
  if (SEP && !recompute) { fp.x = s.heldHit ? s.heldC.x : fin.x; fp.y = s.heldHit ? s.heldC.y : fin.y; fp.z = s.heldHit ? s.heldC.z : fin.z; }

  // inputs of the eight stages: stage 0 takes the new vertex, stage k > 0 what stage k-1 produced in the previous step
  V3p in[4];
  in[0] = pk3(x, SEP ? hi3(s.P[0]) : lo3(s.X[3]));
#pragma unroll
  for (int q = 1; q < 4; ++q) in[q] = SEP ? s.P[q] : s.X[q - 1];
  // sf * rest moves along with its vertex
  const float l3 = lo(s.L[3]);
#pragma unroll
  for (int q = 3; q >= 1; --q) s.L[q] = s.L[q - 1];
  s.L[0] = pk(lx, l3);

  // ---- phase A: the eight projections, as four packed chains ---------------------------------------
  V3p vd[4], D[4];
  u64 dp[4], ninv[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    vd[q] = sub3(in[q], s.X[q]);
    dp[q] = PM::dot(vd[q], vd[q], nz);
  }
  bool ok = true;
  if (PM::kRangeChecked) {
    // All eight operands inside the branch-free range? fmin/fmax skip NaN operands, and a NaN is no reason to leave the
    // branch-free sequence (it returns NaN like the builtins): a strand that went NaN — the reference's arithmetic makes
    // some, normalize(0) — must not drag its warp through the IEEE path at every step for the rest of the run.
    ok = !(min8(dp) < M::kFastLo) && !(max8(dp) >= __int_as_float(0x7f800000));
  }
  const bool ieeeA = neg_inversesqrt_batch<PM>(dp, ninv, ok, nz);
  // outside the branch-free range (a segment shorter than 2^-51 or overflowing) the computed direction need not have unit length
  if (CAPS && ieeeA) s.slack = __int_as_float(0xff800000);
#pragma unroll
  for (int q = 3; q >= 0; --q) {
    D[q] = PM::project(s.X[q], vd[q], ninv[q], s.L[q], nz);
    if (RS == 8) {                                                          // the root passes through: D(0, k) = X[0]
      if (j == q) D[q] = pk3(lo3(in[q]), hi3(D[q]));
      if (j == q + 4) D[q] = pk3(lo3(D[q]), hi3(in[q]));
    } else if (RS == 4) {                                                   // stages q and q + 4 hold roots at the same steps
      if (root_in_stage<RS>(j, q)) D[q] = in[q];
    }
    if (q == 3) {
      const V3 dF = vsub<M>(hi3(D[3]), hi3(in[3]));                         // s_particles[i].velocity = p1_bis - p1 (cs:116)
      V3 fw = M::scale(dF, a.damp);                                         // cs:119-121
      // the tip keeps its own d. nverts % 8 == 0: a root follows it (step 7 of the root chunk); otherwise the slots after it
      // in its last chunk are out of bounds for the tensor map — TMA fills them with NaN on the way in and drops them on the
      // way out — and the tip leaves at step (nverts % 8) - 1 of the next root chunk (tip_step, set by the launcher)
      if (RS == 8 ? j == a.tip_step : root_in_stage<RS>(j, 7)) fw = s.heldd;
      s.heldd = dF;
      if (SEP && !recompute && s.heldHit) fw = M::reflect(fw, s.heldN);     // cs:137, with the normal found one step ago
      if (recompute) {                                                      // sphere + capsules again from D(i, 8), now with the velocity
        BH_T0(t4);
        const PosVel pv = all_pos_vel_slow<M>(a, fp, fw); fp = pv.p; fw = pv.w;
        BH_T1(12, t4);
      }
      const float4 oP = make_float4(fp.x, fp.y, fp.z, rest_out);
      float4 oV = make_float4(fw.x, fw.y, fw.z, 0.f);
      if (fin_root) oV = make_float4(rootV_out.x, rootV_out.y, rootV_out.z, 0.f);   // a root is neither moved nor reflected
      *slotP = oP;
      *slotV = oV;
    }
    s.X[q] = D[q];
  }

  // ---- phase T: collision tests of the eight new positions -----------------------------------------
  const V3p c2 = { pk(a.cx, a.cx), pk(a.cy, a.cy), pk(a.cz, a.cz) };
  V3p pt[4];
  u64 dpc[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) pt[q] = ORIGIN ? D[q] : sub3(D[q], c2);
  // roots do not collide (cs:149-151: index > 0): their distance becomes +inf, which no radius exceeds; in the capsule variant
  // NaN, which the minimum AND the maximum below skip (fminf / fmaxf return the other operand) and which compares false too
  const float no_hit = __int_as_float(CAPS ? 0x7fffffff : 0x7f800000);
  auto mark_roots = [&](u64 (&v)[4]) {
    if (RS == 8) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (j == q) v[q] = pk(no_hit, hi(v[q]));
        if (j == q + 4) v[q] = pk(lo(v[q]), no_hit);
      }
    } else if (RS == 4) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (root_in_stage<RS>(j, q)) v[q] = pk(no_hit, no_hit);
    }
  };
  bool maybe_hit = true;                                                      // warp-uniform
  float mnc = 0.f, mxc = 0.f;                                                 // smallest / (capsule variant) largest |p - c|^2 of this lane
  if (PM::kTwoStageTest && (!SEP || (CAPS && caps_single<PM>()))) {
    // The exact squared distance (five packed operations per pair, the reference's rounding sequence) is only NEEDED by a
    // vertex that is pushed out; whether any vertex is, a contracted sum of squares (three operations) decides for all
    // but the warps within 2^-20 relative of the surface: both sums carry at most three roundings of at most the true value,
    // so fma-chain >= r^2 (1 + 2^-20) implies exact >= r^2 (StepArgs::r2_maybe, +inf where that argument does not hold:
    // subnormal radii). A step that follows a push-out (SEP) nearly always pushes out again and goes straight to the
    // exact sum. The capsule variant's shell test and error term take the contracted sums just as well (their margins are
    // a thousand times the difference).
    u64 df[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) df[q] = fma2(pt[q].z, pt[q].z, fma2(pt[q].y, pt[q].y, mul2_contractable(pt[q].x, pt[q].x)));
    mark_roots(df);
    mnc = min8(df);
    maybe_hit = __any_sync(0xffffffffu, mnc < a.r2_maybe);
    if (!maybe_hit) {
      if (!CAPS) return false;
      mxc = max8(df);
    }
  }
  if (maybe_hit) {
#pragma unroll
    for (int q = 0; q < 4; ++q) dpc[q] = PM::dot(pt[q], pt[q], nz);
    mark_roots(dpc);
    mnc = min8(dpc);
    if (CAPS) mxc = max8(dpc);
  }

  // ---- phase B: push-outs, only when some lane of the warp touches the sphere --------------------
  const bool any_hit = maybe_hit && __any_sync(0xffffffffu, mnc < a.r2);
#ifdef BH_STATS
  {
    int pairs = 0, stages = 0, lanes = 0;
#if BH_STATS > 1   // -DBH_STATS=2: per-pair / per-stage / per-lane hit counts as well (eight more ballots per step)
    for (int q = 0; q < 4; ++q) {
      const unsigned ml = __ballot_sync(0xffffffffu, lo(dpc[q]) < a.r2), mh = __ballot_sync(0xffffffffu, hi(dpc[q]) < a.r2);
      pairs += (ml | mh) != 0; stages += (ml != 0) + (mh != 0); lanes += __popc(ml) + __popc(mh);
    }
#endif
    if ((threadIdx.x & 31) == 0) { BH_STAT(0, 1); BH_STAT(1, any_hit); BH_STAT(2, pairs); BH_STAT(3, stages); BH_STAT(4, lanes); BH_STAT(5, SEP); }
  }
#endif
  V3p C[4];
  auto sphere_push_out = [&]() {
    // a hit has dpc < r2 < inf (launcher guarantees), so only the lower bound of the branch-free range can fail
    u64 ninvc[4];
    neg_inversesqrt_batch<PM>(dpc, ninvc, !(mnc < M::kFastLo), nz);
    const u64 r2p = pk(a.r, a.r);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const V3p Q = ORIGIN ? PM::push_out_origin(pt[q], ninvc[q], r2p, nz) : PM::push_out(c2, pt[q], ninvc[q], r2p, nz);
      C[q] = pk3(sel3_lt(lo(dpc[q]), a.r2, lo3(Q), lo3(D[q])), sel3_lt(hi(dpc[q]), a.r2, hi3(Q), hi3(D[q])));
    }
    // stage 7 is done: keep its final position and the collision normal for the step that writes it out
    s.heldHit = hi(dpc[3]) < a.r2;
    s.heldC = hi3(C[3]);
    s.heldN = M::scale(hi3(pt[3]), PM::inv_of(hi(ninvc[3])));
  };
  auto enter_through_P = [&]() {                                            // stage k+1 takes C of stage k
#pragma unroll
    for (int q = 3; q >= 1; --q) s.P[q] = C[q - 1];
    s.P[0] = pk3(V3{ 0.f, 0.f, 0.f }, lo3(C[3]));
  };
  if (!CAPS) {
    if (any_hit) { sphere_push_out(); enter_through_P(); }
    return any_hit;
  }
  // ---- capsule extension: bound tests on the positions as the sphere left them, exact chain out of line ------------
  // Level 0, nearly free: the squared distances to the sphere centre are there already. All capsules lie inside the shell
  // cap_lo2 < |p - c|^2 < cap_hi2 (fill_capsule_bounds, with margins). The push-out above only moves a vertex outwards, onto
  // the sphere, and cap_lo2 is either beyond the sphere's surface or zero: a lane whose eight positions all stayed below
  // cap_lo2 or above cap_hi2 BEFORE the push-out has none inside the shell after it.
  // Level 1, temporal: the position a stage tests at this step lies within one rest length of the one it tested a step ago —
  // D(i, k) = D(i-1, k) + L_i n with |n| = 1, the same iteration of the previous vertex (a root passes through the stages
  // unchanged, so it is "within" the position it had one stage earlier) — plus, around a sphere push-out, the depth pushed
  // out (tested positions are the pushed-out ones, the chain runs on the projected ones). A lane therefore keeps a lower
  // bound on its distance to the capsules from its last test and lowers it every step by its largest rest length, twice the
  // largest push-out depth of the step and the rounding terms cap_e0 + cap_e1 max |p - c|^2; while the bound stays positive
  // no test is needed. Nothing here looks at other lanes, and a lane without knowledge (a new root, a direction from the
  // IEEE path, NaN anywhere) has a bound that is not positive.
  // Outside the shell nothing is tracked (the common case for capsules the hair does not reach: one vote, like a step without
  // the temporal bound); inside it, lanes that are not in the shell themselves keep counting down.
  const bool in_shell = mnc < a.cap_hi2 && mxc > a.cap_lo2;
  bool any_cap = __any_sync(0xffffffffu, in_shell);
  float eb = 0.f;
  if (!any_cap) s.slack = __int_as_float(0xff800000);
  else {
    eb = __fmaf_rn(mxc, a.cap_e1, a.cap_e0);                                  // NaN when all eight positions are roots / NaN: forces a test
    if (any_hit) eb = __fmaf_rn(fmaxf(0.f, a.r - sqrt_approx(mnc) * 0.9990234375f), 2.0f, eb);   // + 2 x (r - smallest |p - c|)
    s.slack -= __fmaf_rn(s.lmax, 1.0009765625f, eb);
    any_cap = __any_sync(0xffffffffu, in_shell && !(s.slack > 0.f));
  }
#ifdef BH_STATS
  if ((threadIdx.x & 31) == 0) BH_STAT(6, any_cap);
#endif
  // written out for both kinds of step so that the common one — the sphere touched nobody and no lane asks for a test — reads
  // the positions where they are (D) and leaves without having copied them
  if (!any_hit) {
    s.heldHit = false;
#ifdef BH_STATS   // a skipped test that would have let a lane through (must stay zero)
    if (!any_cap) { float d_; if (caps_tight_touch(a, D, d_) && (threadIdx.x & 31) == 0) BH_STAT(14, 1); }
#endif
    if (!any_cap) {
      s.heldCap = false;
      if (caps_single<PM>()) {                              // one step variant: every step hands its positions on through P
#pragma unroll
        for (int q = 0; q < 4; ++q) C[q] = D[q];
        enter_through_P();
        return true;
      }
      return false;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) C[q] = D[q];
  } else {
    sphere_push_out();                                                        // sets every C[q]
#ifdef BH_STATS
    if (!any_cap) { float d_; if (caps_tight_touch(a, C, d_) && (threadIdx.x & 31) == 0) BH_STAT(14, 1); }
#endif
  }
  unsigned cmask = 0u;
  if (any_cap) {                                                              // level 2: distance to the axes, fast arithmetic; renews the bound
    BH_T0(t2);
    float sl;
    cmask = caps_tight_touch(a, C, sl);
    s.slack = sl - eb;
    any_cap = cmask != 0u;
    BH_T1(10, t2);
  }
#ifdef BH_STATS
  if ((threadIdx.x & 31) == 0) BH_STAT(8, any_cap);
#endif
  if (any_cap) {
    // every vertex in flight that is not a root (cs:149-151: index > 0); the vertex in stage 7 is skipped, the next step
    // recomputes its whole chain (sphere included) together with its velocity
    unsigned skip = 0x80u;
#pragma unroll
    for (int k = 0; k < 7; ++k) skip |= root_in_stage<RS>(j, k) ? (1u << k) : 0u;
    BH_T0(t3);
    V3p buf[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) buf[q] = C[q];
    caps_packed<PM>(a, buf, skip, cmask, nz);
#pragma unroll
    for (int q = 0; q < 4; ++q) C[q] = buf[q];
    BH_T1(11, t3);
  }
  s.heldCap = any_cap && !root_in_stage<RS>(j, 7);                          // conservative: recomputing without a hit is a no-op
  if (any_hit || any_cap || caps_single<PM>()) enter_through_P();
  return any_hit || any_cap;
}

// The eight steps of one chunk; `sep` carries the push-out state from step to step.
// `fin_mask`: bit j set when the vertex finalised at step j (slot j of the previous chunk) is a root.
template <class PM, bool ORIGIN, int RS, bool CAPS>
__device__ __forceinline__ void stream_chunk(const StepArgs& a, const u64 nz, Pipe& s, bool& sep, const unsigned fin_mask,
                                             float4* bP, float4* bV, float* myR, const int sw, const int joff = 0) {
  // joff = 8 (RS == 8 only): the chunk holds no root; the step sees slot numbers 8..15, which match no stage
  // Two innermost loops instead of one loop with two bodies: a run of free steps and a run of contact steps each turn on a
  // back-edge of their own (contact comes in runs: 98 % of the steps that follow a push-out push out again), which ptxas
  // allocates and lays out better than one loop that picks its body every step — free step 228 -> 221 instructions, contact
  // step 362 -> 359, configs[1] exact 0.4977 -> 0.4913 ms per launch (A/B on one box).
  int j = 0;
  // Capsule variant: ONE step variant. Its loop (tests, push-out, call sites) in two variants and two chunk kinds does not
  // fit the instruction cache; one variant does, at the price of 24 moves in a step that pushed nothing out (measured at
  // configs[2], 4M x 32: "arms" exact 0.427 -> 0.478, fast 0.56 -> 0.67 of the HBM peak; the plain kernel, whose loop
  // fits, loses with it: 0.4966 -> 0.5094 ms per launch for its root chunks alone).
  if (CAPS && caps_single<PM>()) {
#pragma unroll 1
    for (; j < kK; ++j) stream_step<PM, ORIGIN, RS, true, CAPS>(a, nz, s, j + joff, (fin_mask >> j) & 1u, bP + (j ^ sw), bV + (j ^ sw), myR + j * 32);
    sep = true;
    return;
  }
#pragma unroll 1
  for (;;) {
    if (!sep) {
#pragma unroll 1
      do {
        sep = stream_step<PM, ORIGIN, RS, false, CAPS>(a, nz, s, j + joff, (fin_mask >> j) & 1u, bP + (j ^ sw), bV + (j ^ sw), myR + j * 32);
        ++j;
      } while (!sep && j < kK);
    } else {
#pragma unroll 1
      do {
        sep = stream_step<PM, ORIGIN, RS, true, CAPS>(a, nz, s, j + joff, (fin_mask >> j) & 1u, bP + (j ^ sw), bV + (j ^ sw), myR + j * 32);
        ++j;
      } while (sep && j < kK);
    }
    if (j >= kK) break;
  }
}

// NS = 8: a tensor row is one strand of nverts % 8 == 0 vertices. NS = 4: nverts == 4 and a row is TWO consecutive strands
// (the same 128 bytes), so every chunk is a whole row with roots in slots 0 and 4; a.nstrands is even (launcher).
// FUSED: the launch runs StepArgs::passes steps over groups of tiles (frame-level substep fusion); its own instantiation, so
// that the plain one-step launch keeps its code (and registers) exactly.
template <class PM, bool ORIGIN, int NS, bool CAPS, bool FUSED>
__global__ void __launch_bounds__(kMaxThreads, 1)                           // 384 threads: 168 registers, as with three blocks of 128
hair_step_stream_kernel(const __grid_constant__ StepArgs a, const __grid_constant__ CUtensorMap mapP,
                        const __grid_constant__ CUtensorMap mapV, unsigned int* tile_counter) {
  // (-0.0f, -0.0f), read from device memory: a value ptxas can neither fold into the packed products (it would then
  // contract them with the following add) nor re-load from the constant bank in the middle of the step
  const u64 nz = *reinterpret_cast<const u64*>(tile_counter + 2);
  // Programmatic dependent launch: the next launch of the stream may move onto an SM as soon as this launch's blocks leave it
  // and run its prologue (shared-memory carve-up, barrier init) there; it waits below, before its first tile, for this grid to
  // have completed and flushed. Hides the launch gap behind the tail of the previous launch; changes nothing else.
  asm volatile("griddepcontrol.launch_dependents;");
  extern __shared__ unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;              // 128B-swizzled tiles repeat every 1024 B
  unsigned char* gen = smem_raw + (base - smem_u32(smem_raw));
  unsigned char* tiles = gen + warp * kWarpTileBytes;                       // [stage][plane][32 x 128 B]
  float* ring = reinterpret_cast<float*>(gen + nwarps * kWarpTileBytes + warp * kRingBytes);
  const uint32_t tiles_s = base + warp * kWarpTileBytes;
  const uint32_t bar_s = base + nwarps * (kWarpTileBytes + kRingBytes) + warp * 16;

  const int chunks = NS == 8 ? (a.nverts + kK - 1) / kK : 1;                // per row; the last one may be ragged (see tip_step)
  const long long nrows = NS == 8 ? a.nstrands : a.nstrands / 2;
  const unsigned int ntiles = (unsigned int)((nrows + 31) / 32);
  // Work is handed out in GROUPS of tiles; a warp streams a group `passes` times (pass after pass, tile after tile, chunk
  // after chunk) before it takes the next. passes == 1: a group is one tile and this is the plain one-step launch. The last
  // group absorbs the remainder, so every group has at least group_tiles tiles.
  const int passes = FUSED ? a.passes : 1, gtiles = FUSED ? a.group_tiles : 1;
  const unsigned int ngroups = FUSED ? ntiles / (unsigned int)gtiles : ntiles;

  if (lane == 0) {
    mbar_init(bar_s, 1);
    mbar_init(bar_s + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  auto grab = [&]() -> int {                                                // next group of this warp, -1 when none is left
    unsigned int t = 0;
    if (lane == 0) t = atomicAdd(tile_counter, 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= ngroups) return -1;
    return (int)(a.reverse ? ngroups - 1u - t : t);                           // strands are independent: any order is the same result
  };
  // A stream position is (absolute tile, meta) with meta = chunk | tile-in-group << 16 | pass << 24: advanced without divisions.
  const int lastbase = (int)((ngroups - 1u) * (unsigned int)gtiles);         // first tile of the last group (which absorbs the remainder)
  auto chunk_of = [&](int meta) -> int { return FUSED ? meta & 0xffff : meta; };
  auto issue_load = [&](int tile, int meta, int b) {                        // lane 0 only
    const int c = chunk_of(meta);
    const uint32_t bar = bar_s + 8 * b, dst = tiles_s + b * kStageBytes;
    // a chunk of a later pass is what this warp stored one pass ago: that store is at least one bulk group older than the most
    // recent one (StepArgs::group_tiles), and it must have been performed, not just read out of shared memory
    if (FUSED && (meta >> 24) != 0) tma_wait_complete_but1();
    mbar_expect_tx(bar, kStageBytes);
    tma_load_tile(dst, &mapP, c * 32, tile * 32, bar);
    tma_load_tile(dst + kPlaneTile, &mapV, c * 32, tile * 32, bar);
  };
  auto issue_store = [&](int tile, int meta, int b) {                       // lane 0 only
    const int c = chunk_of(meta);
    const uint32_t src = tiles_s + b * kStageBytes;
    tma_store_tile(&mapP, c * 32, tile * 32, src);
    tma_store_tile(&mapV, c * 32, tile * 32, src + kPlaneTile);
    tma_commit();
  };

  // chunk positions: S = being stored (q-1), C = being computed (q), N = loaded (q+1), L = next to load (q+2)
  // The last warp of the grid to leave puts both words back to zero: no memset between launches.
  auto leave = [&]() {
    if (lane == 0) {
      __threadfence();
      if (atomicAdd(tile_counter + 1, 1u) == gridDim.x * (unsigned)nwarps - 1u) {
        tile_counter[0] = 0u; tile_counter[1] = 0u; __threadfence();
#ifdef BH_STATS
        printf("BH_STATS steps %llu hit %llu pairs %llu stages %llu lanes %llu sep %llu | caps: tested %llu touched %llu | cycles: test %llu chain %llu recompute %llu warp-total %llu | skipped tests that would have passed %llu\n",
               g_stats[0], g_stats[1], g_stats[2], g_stats[3], g_stats[4], g_stats[5], g_stats[6], g_stats[8], g_stats[10], g_stats[11], g_stats[12], g_stats[13], g_stats[14]);
        printf("BH_SPAN launch %llu ns, warps %llu, mean busy %llu ns (%.2f %% of the span idle)\n", g_span[1] - g_span[0], g_span[3], g_span[2] / g_span[3],
               100.0 - 100.0 * (double)g_span[2] / ((double)(g_span[1] - g_span[0]) * (double)g_span[3]));
        for (int i = 0; i < 16; ++i) g_stats[i] = 0;
        for (int i = 0; i < 4; ++i) g_span[i] = 0;
#endif
      }
    }
  };
  asm volatile("griddepcontrol.wait;" ::: "memory");                        // the state (and the scheduler words) as the previous launch left them
  int tC = grab(), cC = 0;
  if (tC < 0) { leave(); return; }
  if (FUSED) tC *= gtiles;
  auto next_pos = [&](int& t, int& c) {                                     // t: tile, c: chunk (FUSED: meta)
    if (t < 0) return;
    if (!FUSED) { if (++c == chunks) { c = 0; t = grab(); } return; }
    if ((c & 0xffff) + 1 < chunks) { c += 1; return; }                      // next chunk of the tile
    c &= ~0xffff;
    const int ti = (c >> 16) & 0xff, gbase = t - ti;
    const int gcount = gbase == lastbase ? (int)ntiles - gbase : gtiles;
    if (ti + 1 < gcount) { c += 1 << 16; t += 1; return; }                  // next tile of the group
    const int p = c >> 24;
    if (p + 1 < passes) { c = (p + 1) << 24; t = gbase; return; }           // next pass over the group
    const int g = grab();                                                   // next group
    t = g < 0 ? -1 : g * gtiles; c = 0;
  };
  int tN = tC, cN = cC; next_pos(tN, cN);
  int tL = tN, cL = cN; next_pos(tL, cL);
  if (lane == 0) {
    issue_load(tC, cC, 0);
    if (tN >= 0) issue_load(tN, cN, 1);
  }
  int tS = -1, cS = 0;

  Pipe s;
  {
    // benign fill values: far from any collider, non-degenerate segments, zero length
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const u64 x = pk(1.0e3f + 10.f * q, 1.0e3f + 10.f * (q + 4)), p = pk(1.0e3f + 10.f * q + 10.f, 1.0e3f + 10.f * (q + 4) + 10.f);
      s.X[q] = { x, x, x }; s.P[q] = { p, p, p }; s.L[q] = 0ull;
    }
    s.heldd = s.rootV[0] = s.rootV[1] = s.heldC = s.heldN = { 0.f, 0.f, 0.f };
    s.heldHit = s.heldCap = false;
    s.slack = __int_as_float(0xff800000); s.lmax = 0.f;
  }
  for (int j = 0; j < kK; ++j) ring[j * 32 + lane] = 0.f;

  const int sw = lane & 7;
  float* myR = ring + lane;
  bool prev_root_chunk = false, sep = false;
  unsigned int q = 0;
#ifdef BH_STATS
  if (lane == 0) for (int i = 0; i < 16; ++i) bh_warp_stats()[i] = 0;
#endif
  BH_T0(t_warp);
#ifdef BH_STATS
  unsigned long long ns0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));
#endif
  for (;; ++q) {
    const int b = q & 1;
    const bool live = tC >= 0;                                              // false: the drain chunk after the last tile
    if (live) mbar_wait(bar_s + 8 * b, (q >> 1) & 1);
    float4* bP = reinterpret_cast<float4*>(tiles + b * kStageBytes) + lane * 8;
    float4* bV = bP + kPlaneTile / 16;
    const bool root_chunk = !live || chunk_of(cC) == 0;
    if (NS == 4) stream_chunk<PM, ORIGIN, 4, CAPS>(a, nz, s, sep, 0x11u, bP, bV, myR, sw);
    // capsule variant: with BH_CAPS_ONE_BODY one step body serves both kinds of chunk (root slots compared at run time). It was
    // the better choice while the variant had a free and a contact step variant; with the single variant of stream_chunk two
    // bodies fit the instruction cache and win ("arms" exact 0.435 -> 0.478 of the HBM peak), so it is off.
    else if (BH_CAPS_ONE_BODY && CAPS && PM::kRangeChecked) stream_chunk<PM, ORIGIN, 8, CAPS>(a, nz, s, sep, prev_root_chunk ? 1u : 0u, bP, bV, myR, sw, root_chunk ? 0 : 8);
    else if (root_chunk) stream_chunk<PM, ORIGIN, 8, CAPS>(a, nz, s, sep, prev_root_chunk ? 1u : 0u, bP, bV, myR, sw);
    else stream_chunk<PM, ORIGIN, 0, CAPS>(a, nz, s, sep, prev_root_chunk ? 1u : 0u, bP, bV, myR, sw);
    prev_root_chunk = root_chunk;
    fence_async_smem();                                                     // generic-proxy writes -> visible to the TMA store
    __syncwarp();
    if (lane == 0) {
      if (tS >= 0) issue_store(tS, cS, b);
      if (tL >= 0 || !live) tma_wait_read0();                               // buffer b is free again
      if (tL >= 0) issue_load(tL, cL, b);
    }
    if (!live) break;
    tS = tC; cS = cC;
    tC = tN; cC = cN;
    tN = tL; cN = cL;
    next_pos(tL, cL);
    if (tC < 0 && lane == 0) tma_wait_read0();                              // the drain chunk reuses buffer b^1
    __syncwarp();
  }
  BH_T1(13, t_warp);
#ifdef BH_STATS
  if (lane == 0) for (int i = 0; i < 16; ++i) atomicAdd(&g_stats[i], bh_warp_stats()[i]);
  if (lane == 0) {                                                          // how much of the launch the warps spend without a tile (tail)
    unsigned long long ns1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
    if (g_span[0] == 0ull) atomicCAS(&g_span[0], 0ull, ns0);
    atomicMin(&g_span[0], ns0); atomicMax(&g_span[1], ns1); atomicAdd(&g_span[2], ns1 - ns0); atomicAdd(&g_span[3], 1ull);
  }
#endif
  leave();
}

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// A plane as a 2-D tensor of fp32: row = strand (stride nverts * 16 B), 4 * nverts floats per row; box = 32 floats x 32 rows.
// nverts == 4: a row is two consecutive strands (8 vertices, 128 B); nstrands is even.
bool make_plane_map(CUtensorMap* map, float4* plane, long long nstrands, int nverts) {
  EncodeTiledFn fn = encode_tiled();
  if (!fn) return false;
  if (nverts == 4) { nstrands /= 2; nverts = 8; }
  const cuuint64_t dims[2] = { (cuuint64_t)nverts * 4, (cuuint64_t)nstrands };
  const cuuint64_t strides[1] = { (cuuint64_t)nverts * 16 };
  const cuuint32_t box[2] = { 32, 32 };
  const cuuint32_t estr[2] = { 1, 1 };
  // L2 promotion: a tile row of >= 1 KB (64+ vertices per strand) gives every 128-byte piece of a box its own DRAM page, and
  // fetching 256 B per request halves the page openings (+5 % measured at 64 and 128 vertices per strand); with short rows
  // the box is (nearly) contiguous and 256 B costs bandwidth (-16 % at the reference's 4 vertices per strand).
  static const int promo_env = [] { const char* e = getenv("BH_TMA_L2_PROMOTION"); return e ? atoi(e) : -1; }();   // tuning knob: 0, 64, 128, 256
  const int promo = promo_env >= 0 ? promo_env : (nverts >= 64 ? 256 : 128);
  const CUtensorMapL2promotion l2 = promo == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : promo == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B :
                                    promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  // Out-of-bounds elements (the slots after the tip when nverts % 8 != 0, the rows after the last strand of a ragged tile)
  // arrive as NaN: NaN vertices run through the branch-free arithmetic at full speed, touch no collider and are never
  // stored, whereas zeros would be degenerate segments (length 0) and send the warp through the IEEE fallback.
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, plane, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA) == CUDA_SUCCESS;
}

struct DeviceInfo { int sms = 0; bool ready[24] = {}; int blocks_per_sm[24] = {}; int big_blocks_per_sm[24] = {}; };

// Bounds in front of the exact capsule arithmetic. Each is widened by 1e-3 relative plus an absolute term — orders of
// magnitude above the fp32 rounding of either the bound or the exact test it guards, so "outside the bound" implies "the
// exact test cannot fire".
void fill_capsule_bounds(StepArgs& b) {
  // Level 0: the shell around the sphere's centre that contains every capsule, lo <= |p - c| <= hi for each of their points:
  // lo = min_k (distance from c to axis k - r_k), hi = max_k (farthest axis end + r_k); margins as below. lo only counts when
  // it lies beyond the sphere's surface (a vertex pushed out of the sphere lands ON it), otherwise it is zero.
  {
    double lo = INFINITY, hi = 0.0;
    bool bad = false;
    for (int k = 0; k < b.ncaps && k < kMaxCapsules; ++k) {
      const Capsule& c = b.caps[k];
      const double ax = (double)c.ax - b.cx, ay = (double)c.ay - b.cy, az = (double)c.az - b.cz;
      const double ux = (double)c.bx - c.ax, uy = (double)c.by - c.ay, uz = (double)c.bz - c.az;
      const double l2 = ux * ux + uy * uy + uz * uz;
      double t = l2 > 0.0 ? -(ax * ux + ay * uy + az * uz) / l2 : 0.0;
      t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
      const double nx = ax + t * ux, ny = ay + t * uy, nz = az + t * uz;
      const double dnear = std::sqrt(nx * nx + ny * ny + nz * nz), r = std::fabs((double)c.r);
      const double da = std::sqrt(ax * ax + ay * ay + az * az);
      const double bx = ax + ux, by = ay + uy, bz = az + uz, db = std::sqrt(bx * bx + by * by + bz * bz);
      if (!std::isfinite(dnear) || !std::isfinite(da) || !std::isfinite(db) || !std::isfinite(r)) bad = true;
      lo = std::fmin(lo, dnear - r);
      hi = std::fmax(hi, std::fmax(da, db) + r);
    }
    lo = lo * (1.0 - 1e-3) - 1e-6;
    hi = hi * (1.0 + 1e-3) + 1e-6;
    const double rs = std::fabs((double)b.r) * (1.0 + 2e-3) + 1e-6;          // where a pushed-out vertex can end up, generously
    float lo2 = (!bad && lo > rs) ? std::nextafter((float)(lo * lo), 0.0f) : 0.0f;
    float hi2 = !bad ? std::nextafter((float)(hi * hi), INFINITY) : INFINITY;
    if (!(lo2 == lo2) || !(hi2 == hi2) || !std::isfinite((double)b.cx + b.cy + b.cz) || !std::isfinite(b.r)) { lo2 = 0.0f; hi2 = INFINITY; }
    b.cap_lo2 = lo2; b.cap_hi2 = hi2;
  }
  double emag = 0.0;
  for (int k = 0; k < b.ncaps && k < kMaxCapsules; ++k) {
    const Capsule& c = b.caps[k];
    // capsule-shaped bound: radius widened by 1e-3 relative plus 1e-5 of the coordinate magnitudes involved
    const float abx = c.bx - c.ax, aby = c.by - c.ay, abz = c.bz - c.az;
    const float l2xx = abx * abx, l2yy = aby * aby, l2zz = abz * abz, l2xy = l2xx + l2yy;
    const float l2 = l2xy + l2zz;                                             // MathExact::dot(ab, ab): one rounding per operation, in its order
    const double mag = 1.0 + std::fmax(std::fmax(std::fabs((double)c.ax), std::fabs((double)c.ay)), std::fabs((double)c.az)) +
                       std::fmax(std::fmax(std::fabs((double)c.bx), std::fabs((double)c.by)), std::fabs((double)c.bz));
    const double Rt = std::fabs((double)c.r) * (1.0 + 1e-3) + 1e-5 * mag;
    b.capt[k][0] = abx; b.capt[k][1] = aby; b.capt[k][2] = abz;
    b.capx[k][0] = abx; b.capx[k][1] = aby; b.capx[k][2] = abz; b.capx[k][3] = l2; b.capx[k][4] = c.r * c.r;
    b.capx[k][5] = b.capx[k][6] = b.capx[k][7] = 0.0f;
    b.capt[k][3] = l2 > 0.0f ? 1.0f / l2 : 0.0f;
    b.capt[k][4] = std::nextafter((float)(Rt * Rt), INFINITY);
    b.capt[k][5] = std::nextafter(std::sqrt(b.capt[k][4]), INFINITY);         // >= sqrt(r2_tight): the surface the temporal bound measures from
    if (!std::isfinite(b.capt[k][4]) || !std::isfinite(b.capt[k][3]) || !std::isfinite(l2)) {   // NaN / overflowing capsule: no second bound
      b.capt[k][0] = b.capt[k][1] = b.capt[k][2] = b.capt[k][3] = 0.0f; b.capt[k][4] = b.capt[k][5] = INFINITY;
    }
    emag += std::fabs((double)c.ax) + std::fabs((double)c.ay) + std::fabs((double)c.az) + std::fabs((double)c.bx) + std::fabs((double)c.by) + std::fabs((double)c.bz);
  }
  // Temporal bound: |p| <= |p - c| + |c| <= (1 + |p - c|^2) / 2 + |c|, and every rounding involved (a position coordinate per
  // step; p - a, the closest point on the axis and the distance to it at a test) is below 2^-20 of |p| + |a| + |b|.
  // 2^-17 of (1 + |p - c|^2 + |c| + |a| + |b|) covers them eight times over.
  emag += 1.0 + std::fabs((double)b.cx) + std::fabs((double)b.cy) + std::fabs((double)b.cz);
  b.cap_e1 = 7.62939453125e-06f;                                               // 2^-17
  b.cap_e0 = std::nextafter((float)(emag * 7.62939453125e-06), INFINITY);
  if (!std::isfinite(b.cap_e0)) b.cap_e0 = INFINITY;                            // never skip a test
}

// Tiles per group of a fused launch: the smallest count with group_tiles * chunks >= 4 (see StepArgs::group_tiles).
int fusion_group_tiles(int nverts) {
  const int chunks = nverts == 4 ? 1 : (nverts + kK - 1) / kK;
  return chunks >= 4 ? 1 : (4 + chunks - 1) / chunks;
}

template <class PM, bool ORIGIN, int NS, bool CAPS, bool FUSED>
cudaError_t launch_stream_tf(const StepArgs& a_in, cudaStream_t stream, unsigned int* tile_counter, int variant) {
  StepArgs a = a_in;
  a.tip_step = a.nverts % kK == 0 ? kK - 1 : a.nverts % kK - 1;
  {
    // threshold of the two-stage sphere test: r^2 (1 + 2^-20), rounded up; the error argument needs normal numbers
    const double t = (double)a.r2 * (1.0 + 9.5367431640625e-07);
    a.r2_maybe = (a.r2 >= 7.888609052210118e-31f && t < 3.0e38) ? std::nextafter((float)t, INFINITY) : INFINITY;   // r^2 >= 2^-100
    if (!(a.r2_maybe >= a.r2)) a.r2_maybe = INFINITY;                          // NaN radius: never skip the exact test
  }
  if (CAPS) fill_capsule_bounds(a);
  static DeviceInfo info[64];
  static std::mutex mu;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  auto kernel = hair_step_stream_kernel<PM, ORIGIN, NS, CAPS, FUSED>;
  int sms = 0, blocks_per_sm = 0, big_blocks_per_sm = 0;                    // copies taken under the lock
  {
    std::lock_guard<std::mutex> g(mu);
    DeviceInfo& di = info[dev];
    if (!di.ready[variant]) {
      if ((e = cudaDeviceGetAttribute(&di.sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
      if ((e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(kWarpsBig > kWarps ? kWarpsBig : kWarps))) != cudaSuccess) return e;
      if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&di.blocks_per_sm[variant], kernel, kWarps * 32, smem_bytes(kWarps))) != cudaSuccess) return e;
      if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&di.big_blocks_per_sm[variant], kernel, kWarpsBig * 32, smem_bytes(kWarpsBig))) != cudaSuccess) return e;
      if (di.blocks_per_sm[variant] < 1) return cudaErrorLaunchOutOfResources;
      di.ready[variant] = true;
    }
    sms = di.sms; blocks_per_sm = di.blocks_per_sm[variant]; big_blocks_per_sm = di.big_blocks_per_sm[variant];
  }
  // Tensor maps are pure functions of (plane address, shape): keep the last few so a frame of substeps on the same
  // shard (or the slices of bh_step_host) does not re-encode them at every launch. Process-wide and never invalidated on
  // purpose: an entry whose buffer was freed can only be hit again by a new buffer at the same address with the same shape,
  // for which the encoded map is the same 128 bytes (cuTensorMapEncodeTiled reads its arguments, not the memory).
  struct MapEntry { const void* pos; const void* vel; long long s; int n; CUtensorMap mapP, mapV; };
  static MapEntry cache[64];
  static int cache_next = 0;
  CUtensorMap mapP, mapV;
  {
    std::lock_guard<std::mutex> g(mu);
    const MapEntry* hit = nullptr;
    for (const MapEntry& m : cache)
      if (m.pos == a.pos && m.vel == a.vel && m.s == a.nstrands && m.n == a.nverts) { hit = &m; break; }
    if (!hit) {
      MapEntry& m = cache[cache_next];
      cache_next = (cache_next + 1) % 64;
      m.pos = nullptr;
      if (!make_plane_map(&m.mapP, a.pos, a.nstrands, a.nverts) || !make_plane_map(&m.mapV, a.vel, a.nstrands, a.nverts))
        return cudaErrorInvalidValue;
      m.pos = a.pos; m.vel = a.vel; m.s = a.nstrands; m.n = a.nverts;
      hit = &m;
    }
    mapP = hit->mapP; mapV = hit->mapV;
  }
  const long long ntiles = ((NS == 8 ? a.nstrands : a.nstrands / 2) + 31) / 32;
  if (a.passes < 1) a.passes = 1;
  a.group_tiles = a.passes > 1 ? fusion_group_tiles(a.nverts) : 1;
  if (a.passes > 1 && ntiles < a.group_tiles) return cudaErrorInvalidValue;   // callers ask stream_fusion_eligible() first
  const long long ngroups = ntiles / a.group_tiles;
  static const int occ_cap = [] { const char* e = getenv("BH_STREAM_BLOCKS_PER_SM"); return e ? atoi(e) : 0; }();   // tuning knob
  static const int big_ok = [] { const char* e = getenv("BH_STREAM_BIG_BLOCKS"); return e ? atoi(e) : 1; }();        // tuning knob: 0 = small blocks always, 2 = big blocks for every shape and profile (tests)
  // one big block per SM when the shard gives every warp of every SM at least two groups and the big block holds no fewer
  // warps than the small ones together; otherwise small blocks, which spread a small shard over the SMs
  // (exact profile and fused frames: the HBM-bound fast launch measures 0.3 % slower with it)
  const bool big = big_ok && occ_cap == 0 && big_blocks_per_sm >= 1 && big_blocks_per_sm * kWarpsBig >= blocks_per_sm * kWarps &&
                   (big_ok == 2 || ((PM::kRangeChecked || FUSED) && ngroups >= 2ll * sms * big_blocks_per_sm * kWarpsBig));
  const int warps = big ? kWarpsBig : kWarps;
  long long blocks = (ngroups + warps - 1) / warps;
  int per_sm = big ? big_blocks_per_sm : blocks_per_sm;
  if (occ_cap > 0 && occ_cap < per_sm) per_sm = occ_cap;
  const long long resident = (long long)sms * per_sm;
  if (blocks > resident) blocks = resident;
  static const bool pdl = [] { const char* e = getenv("BH_STREAM_PDL"); return !e || atoi(e) != 0; }();   // tuning knob: 0 = plain stream order
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)blocks); cfg.blockDim = dim3(warps * 32); cfg.dynamicSmemBytes = smem_bytes(warps); cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  e = cudaLaunchKernelEx(&cfg, kernel, a, mapP, mapV, tile_counter);
  return e != cudaSuccess ? e : cudaGetLastError();
}

// Exhaustive check of the branch-free inverse square roots against the IEEE builtins: every finite binary32
// >= 2^-102, scalar (MathExact) and packed (PackedExact, both halves).
__global__ void selftest_inversesqrt_kernel(unsigned long long* bad, const u64 nz) {
  const unsigned int lo_bits = 0x0C800000u, hi_bits = 0x7F800000u;         // 2^-102, +inf
  unsigned long long nb = 0;
  for (unsigned long long b = lo_bits + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < hi_bits;
       b += (unsigned long long)gridDim.x * blockDim.x) {
    const float x = __uint_as_float((unsigned int)b);
    const float want = __frcp_rn(__fsqrt_rn(x));
    const float got_s = MathExact::inversesqrt_in_range(x);
    const float x2 = __uint_as_float((unsigned int)(hi_bits - 1 - (b - lo_bits)));   // a different value in the other half
    const u64 got_p = PackedExact::neg_inversesqrt_in_range(pk(x, x2), nz);
    const float want2 = __frcp_rn(__fsqrt_rn(x2));
    nb += (__float_as_uint(got_s) != __float_as_uint(want)) + (__float_as_uint(-lo(got_p)) != __float_as_uint(want)) +
          (__float_as_uint(-hi(got_p)) != __float_as_uint(want2));
  }
  if (nb) atomicAdd(bad, nb);
}

}  // namespace

cudaError_t selftest_inversesqrt(unsigned long long* mismatches) {
  unsigned long long* d = nullptr;
  cudaError_t e = cudaMalloc(&d, sizeof *d);
  if (e != cudaSuccess) return e;
  e = cudaMemset(d, 0, sizeof *d);
  if (e == cudaSuccess) { selftest_inversesqrt_kernel<<<148 * 8, 256>>>(d, 0x8000000080000000ull); e = cudaGetLastError(); }
  if (e == cudaSuccess) e = cudaMemcpy(mismatches, d, sizeof *d, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return e;
}

bool stream_kernel_eligible(const StepArgs& a) {
  static const bool disabled = [] { const char* e = getenv("BH_NO_STREAM_KERNEL"); return e && e[0] == '1'; }();
  // any nverts >= 2 (a ragged last chunk is handled by the tensor map's bounds, see make_plane_map); the reference's own
  // nverts = 4 packs two strands per tensor row (an even number of strands: the launcher gives the last strand of an odd
  // count to the per-strand kernel)
  static const bool ragged_ok = [] { const char* e = getenv("BH_NO_RAGGED_STREAM"); return !(e && e[0] == '1'); }();
  const bool shape_ok = a.nverts == 4 ? a.nstrands >= 2 : (a.nverts >= 2 && (ragged_ok || a.nverts % kK == 0));
  return !disabled && a.iterations == kK && a.ncaps >= 0 && a.ncaps <= kMaxCapsules && shape_ok &&
         a.nstrands <= 0x7fffffffLL && a.r2 <= 1.8446744073709551616e19f && encode_tiled() != nullptr;
}

template <class PM, bool ORIGIN, int NS, bool CAPS = false>
cudaError_t launch_stream_t(const StepArgs& a, cudaStream_t stream, unsigned int* tile_counter, int variant) {
  return a.passes > 1 ? launch_stream_tf<PM, ORIGIN, NS, CAPS, true>(a, stream, tile_counter, variant + 12)
                      : launch_stream_tf<PM, ORIGIN, NS, CAPS, false>(a, stream, tile_counter, variant);
}

bool stream_fusion_eligible(const StepArgs& a, int passes, bool always) {
  static const bool disabled = [] { const char* e = getenv("BH_NO_SUBSTEP_FUSION"); return e && e[0] == '1'; }();   // tuning knob
  if (disabled || passes < 2 || passes > 64 || !stream_kernel_eligible(a)) return false;   // the pass counter shares a word with the chunk index
  if (a.nverts == 4 && a.nstrands % 2) return false;                        // the odd last strand runs on the per-strand kernel: step by step
  const long long ntiles = ((a.nverts == 4 ? a.nstrands / 2 : a.nstrands) + 31) / 32;
  const int gtiles = fusion_group_tiles(a.nverts);
  if (ntiles < gtiles) return false;
  return always || gtiles == 1 || ntiles / gtiles >= 1024;
}

cudaError_t launch_step_stream(const StepArgs& a, int math, cudaStream_t stream, unsigned int* tile_counter) {
#ifdef BH_ONLY_MAIN   // compile-time experiments (tools/quick_sass.sh -DBH_ONLY_MAIN): the configs[1] exact kernel alone
  return launch_stream_tf<PackedExact, true, 8, false, false>(a, stream, tile_counter, 0);
#endif
  // x - (+0.0f) == x bit for bit, for every x (a -0.0f centre component would turn a -0.0f coordinate into +0.0f)
  const bool origin = __builtin_bit_cast(uint32_t, a.cx) == 0u && __builtin_bit_cast(uint32_t, a.cy) == 0u && __builtin_bit_cast(uint32_t, a.cz) == 0u;
  if (a.ncaps > 0) {                                                        // capsule extension: one variant per profile and row shape
    if (a.nverts == 4) return math == 0 ? launch_stream_t<PackedExact, false, 4, true>(a, stream, tile_counter, 8) : launch_stream_t<PackedFast, false, 4, true>(a, stream, tile_counter, 9);
    return math == 0 ? launch_stream_t<PackedExact, false, 8, true>(a, stream, tile_counter, 10) : launch_stream_t<PackedFast, false, 8, true>(a, stream, tile_counter, 11);
  }
  if (a.nverts == 4) {                                                      // a.nstrands is even here (launch_step)
    if (math == 0) return origin ? launch_stream_t<PackedExact, true, 4>(a, stream, tile_counter, 4) : launch_stream_t<PackedExact, false, 4>(a, stream, tile_counter, 5);
    return origin ? launch_stream_t<PackedFast, true, 4>(a, stream, tile_counter, 6) : launch_stream_t<PackedFast, false, 4>(a, stream, tile_counter, 7);
  }
  if (math == 0) return origin ? launch_stream_t<PackedExact, true, 8>(a, stream, tile_counter, 0) : launch_stream_t<PackedExact, false, 8>(a, stream, tile_counter, 1);
  return origin ? launch_stream_t<PackedFast, true, 8>(a, stream, tile_counter, 2) : launch_stream_t<PackedFast, false, 8>(a, stream, tile_counter, 3);
}

}  // namespace bh
