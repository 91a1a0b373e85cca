// ref_glsl_harness.cpp — runs the reference's cs_simulation.glsl SOURCE on the CPU.
//
// TEST INFRASTRUCTURE ONLY (see oracle/barbu_hair_oracle.h). Built by oracle/Makefile into
// oracle/_ref/libbarbu_ref_glsl_N<n>.so, one library per control-point count REF_N.
//
// The shader text is pulled in from oracle/_ref/cs_simulation.gen.inc, which oracle/make_ref.py
// derives from /root/reference/src/shaders/hair/01_simulation/cs_simulation.glsl by lexical edits
// only; every GLSL built-in resolves to the reference's own vendored GLM (third_party/glm, 0.9.9.9).
// One workgroup = REF_N cooperative fibers; memoryBarrierShared() yields to the scheduler so that
// it behaves as barrier() — the documented patch (2) of SURVEY.md §8c. The host side mirrors
// Hair::update (src/fx/hair.cc:101-122): bind READ/WRITE planes, set the uniforms, dispatch one
// workgroup per strand, then PingPongBuffer::swap() (src/memory/pingpong_buffer.cc:73-84).
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#define GLM_FORCE_SWIZZLE
#include "glm/glm.hpp"

#include "shaders/hair/interop.h"                 // from the reference tree (-I<ref>/src)
#undef HAIR_MAX_PARTICLE_PER_STRAND               // documented patch (1): N override
#define HAIR_MAX_PARTICLE_PER_STRAND REF_N

namespace ref_glsl {
using namespace glm;

struct Invocation {
  ucontext_t ctx;
  struct { uint x; } global_id;
  uint local_index;
  bool done;
  std::vector<char> stack;
};
static thread_local Invocation* g_cur = nullptr;
static thread_local ucontext_t g_sched;

static inline void workgroup_barrier() { swapcontext(&g_cur->ctx, &g_sched); }
static inline void memoryBarrierShared() { workgroup_barrier(); }
static inline void groupMemoryBarrier() {}        // cs:160,194 — no effect on results (App. A note 4)

// GLSL fma() is component-wise; GLM (GLM_HAS_CXX11_STL) only imports the scalar std::fma.
static inline vec3 fma(vec3 const& a, vec3 const& b, vec3 const& c) {
  return vec3(std::fma(a.x, b.x, c.x), std::fma(a.y, b.y, c.y), std::fma(a.z, b.z, c.z));
}

#define uniform static thread_local
#define shared static thread_local
#define gl_GlobalInvocationID (g_cur->global_id)
#define gl_LocalInvocationIndex (g_cur->local_index)
#include "cs_simulation.gen.inc"
#undef uniform
#undef shared

static void fiber_entry() {
  shader_main();
  g_cur->done = true;
  swapcontext(&g_cur->ctx, &g_sched);
}

struct Workgroup {
  std::vector<Invocation> inv;
  Workgroup() : inv(REF_N) {
    for (auto& i : inv) i.stack.resize(64 * 1024);
  }
  void dispatch(uint group) {
    for (uint l = 0; l < REF_N; ++l) {
      Invocation& i = inv[l];
      getcontext(&i.ctx);
      i.ctx.uc_stack.ss_sp = i.stack.data();
      i.ctx.uc_stack.ss_size = i.stack.size();
      i.ctx.uc_link = &g_sched;
      i.global_id.x = group * REF_N + l;
      i.local_index = l;
      i.done = false;
      makecontext(&i.ctx, fiber_entry, 0);
    }
    // Round-robin: every invocation runs up to its next barrier before any proceeds past it.
    for (bool any = true; any;) {
      any = false;
      for (auto& i : inv) {
        if (i.done) continue;
        g_cur = &i;
        swapcontext(&g_sched, &i.ctx);
        any = any || !i.done;
      }
    }
  }
};
}  // namespace ref_glsl

extern "C" int ref_glsl_nverts() { return REF_N; }

// One Hair::update: pos4/vel4 are the READ buffer planes (S*REF_N float4 each), updated in place
// with the content of the WRITE planes after the dispatch (= swap()).
extern "C" void ref_glsl_update(float* pos4, float* vel4, int64_t nstrands, float dt, float scale,
                                const float* sphere4) {
  using namespace ref_glsl;
  const size_t V = (size_t)nstrands * REF_N;
  std::vector<vec4> wpos(V), wvel(V);
#pragma omp parallel
  {
    read_positions = reinterpret_cast<vec4*>(pos4);
    read_velocities = reinterpret_cast<vec4*>(vel4);
    write_positions = wpos.data();
    write_velocities = wvel.data();
    uTimeStep = dt;                                // hair.cc:107
    uScaleFactor = scale;                          // hair.cc:108
    uBoundingSphere = vec4(sphere4[0], sphere4[1], sphere4[2], sphere4[3]);  // hair.cc:110
    Workgroup wg;
#pragma omp for schedule(static)
    for (int64_t s = 0; s < nstrands; ++s) wg.dispatch((uint)s);  // glDispatchCompute(nroots_,1,1)
  }
  std::memcpy(pos4, wpos.data(), V * sizeof(vec4));  // PingPongBuffer::swap
  std::memcpy(vel4, wvel.data(), V * sizeof(vec4));
}
