// ref_tess_skin_harness.cpp — runs the reference's tess-stream stages and dual-quaternion skinning SOURCES on the CPU.
//
// TEST INFRASTRUCTURE ONLY (see oracle/barbu_hair_oracle.h). Built by oracle/Makefile into
// oracle/_ref/libbarbu_ref_tess_skin.so from oracle/_ref/tess_skin.gen.inc, which oracle/make_ref.py derives from
//   /root/reference/src/shaders/shared/inc_maths.glsl:107-115,210-228,239-241,263-266   (sample_triangle2, hermite_mix,
//                                                                                         maprange, smoothstep2)
//   /root/reference/src/shaders/shared/inc_skinning.glsl:22-31,37-52,54-82              (apply_skinning,
//                                                                 get_dual_quaternions_matrices, skinning_DQBS)
//   /root/reference/src/shaders/hair/02_tess_stream/{vs,tcs,tes,gs}_stream_hair.glsl    (the four `main` bodies)
// by lexical edits only; every GLSL built-in and every vector/matrix operator resolves to the reference's vendored GLM
// (third_party/glm 0.9.9.9). This file supplies what the GL pipeline supplies: the stage interface variables, the
// fixed-function plumbing between the stages (vertex fetch through the element buffer, patch assembly, isoline
// tessellation coordinates, line assembly into the geometry shader, transform-feedback capture order) and the samplerBuffer.
//
// Defined here, not by the reference (the GL specification leaves them to the implementation / the reference fills them
// from std::random_device): isoline tess coordinates x = k / nsubsegments, y = line / nlines (equal_spacing), primitive
// order instance-major then patch, line, segment, and the CONTENT of the random table — the caller passes it. The index
// into the table is the reference's expression (tes_stream_hair.glsl:52-54).
#include <cmath>
#include <cstdint>
#include <vector>

#define GLM_FORCE_SWIZZLE
#include "glm/glm.hpp"

#include "shaders/hair/interop.h"                 // HAIR_TF_RANDOMBUFFER_SIZE, from the reference tree (-I<ref>/src)
#undef HAIR_MAX_PARTICLE_PER_STRAND               // documented patch (1) of SURVEY.md §8c: N at run time
#define HAIR_MAX_PARTICLE_PER_STRAND g_nverts

namespace ref_ts {
using namespace glm;
static thread_local int g_nverts = 4;

namespace shared_inc {
using namespace glm;
// layout(binding = 0) uniform samplerBuffer uSkinningDatas (inc_skinning.glsl:8): RGBA32F texels
static thread_local const vec4* uSkinningDatas = nullptr;
static inline vec4 texelFetch(const vec4* buffer, int texel) { return buffer[texel]; }
// subroutine uniform skinning_subroutine uSkinning (inc_skinning.glsl:16), bound to the DQBS subroutine
#define uSkinning skinning_DQBS
}  // namespace shared_inc

// ---- stage interfaces (the layout(...) declarations of each shader file) -----------------------------------------
#define STAGE_VS_INTERFACE                                                                                          \
  static thread_local vec4 inPosition, inTangent;                                                                  \
  static thread_local vec3 outPosition, outTangent;                                                                \
  static thread_local int outInstanceID, gl_InstanceID, gl_VertexID;                                               \
  static thread_local float outRelativePos;
#define STAGE_TCS_INTERFACE                                                                                         \
  static thread_local vec3 inPosition[6], inTangent[6], outPosition[6], outTangent[6];                             \
  static thread_local int inInstanceID[6], outInstanceID, gl_InvocationID;                                         \
  static thread_local float inRelativePos[6], outRelativePos[6], gl_TessLevelOuter[4];                             \
  static thread_local int uNumLines = 1, uNumSubSegments = 1;                                                      \
  static thread_local float uScaleFactor;
#define STAGE_TES_INTERFACE                                                                                         \
  static thread_local vec3 gl_TessCoord, inPosition[6], inTangent[6];                                              \
  static thread_local int inInstanceID, uNumInstances;                                                             \
  static thread_local float inRelativePos[6];                                                                      \
  static thread_local vec4 outPosition, outTangent;                                                                \
  static thread_local const vec3* randbuffer;
#define STAGE_GS_INTERFACE                                                                                          \
  static thread_local vec4 inPosition[2], position_xyz_coeff_w;                                                    \
  static thread_local vec4* g_capture;                                                                             \
  static inline void EmitVertex() { *g_capture++ = position_xyz_coeff_w; }                                         \
  static inline void EndPrimitive() {}

#include "tess_skin.gen.inc"
}  // namespace ref_ts

// The tess-stream pass of Hair::render (src/fx/hair.cc:141-173): glDrawElementsInstanced(GL_PATCHES, nelems, ..., ninstances)
// over the position/tangent planes with 6 control points per patch, captured by transform feedback as GL_LINES vertices.
// randtable: HAIR_TF_RANDOMBUFFER_SIZE x 2 floats (.xy of each `vec3 randbuffer[]` element). out4: npatches * ninstances *
// nlines * nsubsegments * 2 float4.
extern "C" void ref_tess_stream(const float* pos4, const float* tan4, const int32_t* patch, int64_t npatches, int nverts,
                                float scale, int ninstances, int nlines, int nsubsegments, const float* randtable,
                                float* out4) {
  using namespace ref_ts;
  g_nverts = nverts;
  std::vector<vec3> table(HAIR_TF_RANDOMBUFFER_SIZE);
  for (int i = 0; i < HAIR_TF_RANDOMBUFFER_SIZE; ++i) table[i] = vec3(randtable[2 * i], randtable[2 * i + 1], 0.0f);
  const vec4* P = reinterpret_cast<const vec4*>(pos4);
  const vec4* T = reinterpret_cast<const vec4*>(tan4);
  stage_gs::g_capture = reinterpret_cast<vec4*>(out4);
  for (int inst = 0; inst < ninstances; ++inst)
    for (int64_t pa = 0; pa < npatches; ++pa) {
      // vertex shader, once per control point of the patch
      for (int k = 0; k < 6; ++k) {
        const int e = patch[6 * pa + k];
        stage_vs::inPosition = P[e]; stage_vs::inTangent = T[e];
        stage_vs::gl_VertexID = e; stage_vs::gl_InstanceID = inst;
        stage_vs::vs_main();
        stage_tcs::inPosition[k] = stage_vs::outPosition; stage_tcs::inTangent[k] = stage_vs::outTangent;
        stage_tcs::inInstanceID[k] = stage_vs::outInstanceID; stage_tcs::inRelativePos[k] = stage_vs::outRelativePos;
      }
      // tessellation control shader, layout(vertices = 6) out: one invocation per output control point
      stage_tcs::uNumLines = nlines; stage_tcs::uNumSubSegments = nsubsegments; stage_tcs::uScaleFactor = scale;   // hair.cc:150-154
      for (int id = 0; id < 6; ++id) { stage_tcs::gl_InvocationID = id; stage_tcs::tcs_main(); }
      for (int k = 0; k < 6; ++k) {
        stage_tes::inPosition[k] = stage_tcs::outPosition[k]; stage_tes::inTangent[k] = stage_tcs::outTangent[k];
        stage_tes::inRelativePos[k] = stage_tcs::outRelativePos[k];
      }
      stage_tes::inInstanceID = stage_tcs::outInstanceID;
      stage_tes::uNumInstances = ninstances;
      stage_tes::randbuffer = table.data();
      // isolines: outer[0] lines of outer[1] segments; the evaluation shader runs once per isoline vertex
      const int lines = (int)stage_tcs::gl_TessLevelOuter[0], segs = (int)stage_tcs::gl_TessLevelOuter[1];
      for (int line = 0; line < lines; ++line) {
        vec4 prev(0.0f);
        for (int k = 0; k <= segs; ++k) {
          stage_tes::gl_TessCoord = vec3((float)k / (float)segs, (float)line / (float)lines, 0.0f);
          stage_tes::tes_main();
          if (k > 0) {                                                    // geometry shader: layout(lines) in, two vertices out
            stage_gs::inPosition[0] = prev; stage_gs::inPosition[1] = stage_tes::outPosition;
            stage_gs::gs_main();
          }
          prev = stage_tes::outPosition;
        }
      }
    }
}

// apply_skinning (inc_skinning.glsl:22-31) with uSkinning = skinning_DQBS on S vertices. dq palette: njoints * 8 floats =
// texels 2j (real part) and 2j + 1 (dual part) of uSkinningDatas, as SkeletonController uploads them
// (src/fx/animation/skeleton_controller.cc:248-265).
extern "C" void ref_skin_dq(const float* rest_pos3, const float* rest_nrm3, const int32_t* joints4, const float* weights3,
                            const float* dq_palette, int64_t S, float* out_pos3, float* out_nrm3) {
  using namespace ref_ts;
  shared_inc::uSkinningDatas = reinterpret_cast<const vec4*>(dq_palette);
  for (int64_t s = 0; s < S; ++s) {
    vec3 v(rest_pos3[3 * s], rest_pos3[3 * s + 1], rest_pos3[3 * s + 2]);
    vec3 n(rest_nrm3[3 * s], rest_nrm3[3 * s + 1], rest_nrm3[3 * s + 2]);
    const uvec4 idx((uint)joints4[4 * s], (uint)joints4[4 * s + 1], (uint)joints4[4 * s + 2], (uint)joints4[4 * s + 3]);
    const vec4 w(weights3[3 * s], weights3[3 * s + 1], weights3[3 * s + 2], 0.0f);
    shared_inc::apply_skinning(idx, w, v, n);
    for (int c = 0; c < 3; ++c) { out_pos3[3 * s + c] = v[c]; out_nrm3[3 * s + c] = n[c]; }
  }
}
