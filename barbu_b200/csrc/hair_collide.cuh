// hair_collide.cuh — scalar collision helpers shared by the step kernels (hair_step.cu, hair_stream.cu):
// CollideSphere (cs_simulation.glsl:129-139) and its capsule extension, in either arithmetic profile.
#pragma once
#include "hair_math.cuh"
#include "hair_step.cuh"

namespace bh {

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

// ---- variants for the streaming kernel, which fills StepArgs::capt ----------------------------------------------
// The capsule-shaped bound of StepArgs::capt for one position: squared distance to the axis in fast arithmetic against the
// widened radius — the scalar form of caps_tight_touch() in hair_stream.cu, operation for operation, so that a lane passes
// here exactly when it made its warp pass there. Out of the warps that reach the exact chain only few LANES (and few of the
// eight vertices in flight) are near a capsule: gating the IEEE divisions and square roots with this instead of the
// bounding sphere keeps the chain to the vertices that are about to be pushed out.
__device__ __forceinline__ bool capsule_tight_hit(const Capsule& c, const float (&t)[8], V3 p) {
  const float apx = p.x - c.ax, apy = p.y - c.ay, apz = p.z - c.az;
  const float d = fmaf(apz, t[2], fmaf(apy, t[1], apx * t[0]));
  const float nt = -__saturatef(d * t[3]);                                   // -clamp(t, 0, 1); NaN -> 0
  const float ex = fmaf(nt, t[0], apx), ey = fmaf(nt, t[1], apy), ez = fmaf(nt, t[2], apz);
  return fmaf(ez, ez, fmaf(ey, ey, ex * ex)) < t[4];
}
// Sphere, then capsules, position and velocity (the last iteration of a vertex).
template <class M>
__device__ __forceinline__ void collide_all_pos_vel_bounded(const StepArgs& a, V3& p, V3& w) {
  collide_pos_vel<M>(p, w, V3{ a.cx, a.cy, a.cz }, a.r, a.r2);
#pragma unroll 1
  for (int q = 0; q < a.ncaps; ++q) {
    if (!capsule_tight_hit(a.caps[q], a.capt[q], p)) continue;
    const float r = a.caps[q].r;
    collide_pos_vel<M>(p, w, capsule_center<M>(a.caps[q], p), r, M::mul(r, r));
  }
}

}  // namespace bh
