// ref_marschner_harness.cpp — runs the reference's Marschner LUT shader SOURCES on the CPU.
//
// TEST INFRASTRUCTURE ONLY (see oracle/barbu_hair_oracle.h). Built by oracle/Makefile into
// oracle/_ref/libbarbu_ref_marschner.so from oracle/_ref/marschner.gen.inc, which oracle/make_ref.py derives from
// /root/reference/src/shaders/hair/marschner/{cs_marschner_m,cs_marschner_n,inc_marschner_n}.glsl and the shared
// includes they pull in, by lexical edits only. Every GLSL built-in resolves to the reference's vendored GLM
// (third_party/glm 0.9.9.9), i.e. to libm. The host side mirrors Marschner::generate (src/fx/marschner.cc:35-69): set
// the uniforms, dispatch one invocation per texel of a kTextureResolution^2 image; imageStore keeps the fp32 value
// (the rgba16f conversion of the image format is applied by the caller).
//
// GLSL implicit int -> float conversions the shader text relies on (pow(x, 2), min(L, 1)) are spelled as overloads
// here; GLSL pow(x, y) is undefined for x < 0 — GLM's (std::pow) value is what this harness pins.
#include <cmath>
#include <cstdint>

#define GLM_FORCE_SWIZZLE
#include "glm/glm.hpp"

namespace ref_marschner {
using namespace glm;

static inline float pow(float a, int b) { return std::pow(a, static_cast<float>(b)); }
static inline float pow(float a, float b) { return std::pow(a, b); }
static inline float min(float a, int b) { return glm::min(a, static_cast<float>(b)); }
static inline float min(float a, float b) { return glm::min(a, b); }
static inline float smoothstep(float a, float b, float x) { return glm::smoothstep(a, b, x); }
static inline float smoothstep(int a, float b, float x) { return glm::smoothstep(static_cast<float>(a), b, x); }
static inline vec3 fma(vec3 const& a, vec3 const& b, vec3 const& c) {
  return vec3(std::fma(a.x, b.x, c.x), std::fma(a.y, b.y, c.y), std::fma(a.z, b.z, c.z));
}

struct GlobalId {
  uint x, y;
  uvec2 xy() const { return uvec2(x, y); }
};
static thread_local GlobalId g_id;
static thread_local vec4* g_image;
static thread_local int g_width;
static const int uDstImg = 0;
static inline void imageStore(int, ivec2 const& at, vec4 const& v) { g_image[at.y * g_width + at.x] = v; }

#define uniform static thread_local
#define gl_GlobalInvocationID g_id
#include "marschner.gen.inc"
#undef uniform
}  // namespace ref_marschner

// params: eta, absorption, eccentricity, ar, br, glintScale, azimuthalWidth, deltaCaustic, deltaHm (marschner.h:38-52)
extern "C" void ref_marschner_luts(const float* params, int resolution, float* m_rgba, float* n_rgba) {
  using namespace ref_marschner;
  const float inv = 1.0f / static_cast<float>(resolution);                 // marschner.h:33
#pragma omp parallel
  {
    uEta = params[0]; uAbsorption = params[1]; uEccentricity = params[2];
    uLongitudinalShift = params[3]; uLongitudinalWidth = params[4];        // marschner.cc:47-48
    uGlintScale = params[5]; uAzimuthalWidth = params[6]; uDeltaCaustic = params[7]; uDeltaHm = params[8];
    uInvResolution = inv;
    g_width = resolution;
#pragma omp for schedule(static)
    for (int y = 0; y < resolution; ++y)
      for (int x = 0; x < resolution; ++x) {
        g_id.x = (uint)x; g_id.y = (uint)y;
        g_image = reinterpret_cast<vec4*>(m_rgba);
        marschner_m_main();
        g_image = reinterpret_cast<vec4*>(n_rgba);
        marschner_n_main();
      }
  }
}
