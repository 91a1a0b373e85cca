// ref_hair_host_harness.cpp — runs the reference's host-side strand generation code on the CPU.
//
// TEST INFRASTRUCTURE ONLY (see oracle/barbu_hair_oracle.h). Built by oracle/Makefile into
// oracle/_ref/libbarbu_ref_host_N<n>.so.
//
// src/fx/hair.cc cannot be compiled as a whole (it needs an OpenGL context and the asset managers),
// so oracle/make_ref.py slices the two pure-CPU pieces out of it at build time and this file
// provides the few members they touch:
//   _ref/hair_init_simulation.gen.inc   Hair::init_simulation body, hair.cc:237-328
//   _ref/hair_init_mesh.gen.inc         Hair::init_mesh element loop, hair.cc:397-409
//   _ref/raw_recalc_normals.gen.inc     RawMeshData::recalculateNormals, src/utils/raw_mesh_file.cc:11-50
//   _ref/generate_skinning_datas.gen.inc  SkeletonController::generate_skinning_datas, skeleton_controller.cc:248-265
// GLM (glm::simplex, vec types) is the reference's vendored copy.
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "glm/glm.hpp"
#include "glm/gtc/constants.hpp"
#include "glm/gtc/noise.hpp"
#define GLM_ENABLE_EXPERIMENTAL                    // as the reference does for gtx/ (src/fx/animation/common.h)
#include "glm/gtx/dual_quaternion.hpp"
#include <memory>

#include "shaders/hair/interop.h"                 // binding numbers, NUM_SSBO_HAIR_SIM_ATTRIBS
#undef HAIR_MAX_PARTICLE_PER_STRAND               // documented patch (1): N override
#define HAIR_MAX_PARTICLE_PER_STRAND REF_N

namespace {

// hair.cc:19-20
constexpr int kNumControlPoints = HAIR_MAX_PARTICLE_PER_STRAND;
constexpr int kNumControlSegments = kNumControlPoints - 1;

// The members of MeshData (src/memory/resources/mesh_data.h:62-92,118-125) the slices read.
struct MeshData {
  struct Vertex_t { glm::vec3 position; glm::vec3 normal; };
  std::vector<Vertex_t> vertices;
  std::vector<uint32_t> indices;
  int32_t nfaces() const { return static_cast<int32_t>(indices.size() / 3); }
};

// PingPongBuffer::attrib_index (src/memory/pingpong_buffer.h:57-59) with base_binding = 0.
struct PingPongStub {
  uint32_t attrib_index(uint32_t attrib_bind) const noexcept { return attrib_bind - SSBO_HAIR_SIM_FIRST_BINDING; }
};

struct HairRef {
  struct { struct { float maxlength = 0.50f; } sim; } params_;   // hair.h:27-30
  int nroots_ = 0;
  PingPongStub pbuffer_;
  std::vector<glm::vec3> normals_;
  struct { int nelems; int patchsize; } mesh_{};
  std::array<std::vector<glm::vec4>, NUM_SSBO_HAIR_SIM_ATTRIBS> captured_;
  std::vector<int> captured_elements_;

  void init_simulation(MeshData const& scalpMesh) {
#include "hair_init_simulation.gen.inc"
    for (size_t a = 0; a < host_attribs_.size(); ++a) captured_[a] = host_attribs_[a];
  }

  void init_mesh(MeshData const& scalpMesh) {
#include "hair_init_mesh.gen.inc"
    captured_elements_ = elements;
  }
};

MeshData make_mesh(const float* pos3, const float* nrm3, int64_t nverts, const int32_t* tri, int64_t nfaces) {
  MeshData m;
  m.vertices.resize((size_t)nverts);
  for (int64_t j = 0; j < nverts; ++j) {
    m.vertices[j].position = glm::vec3(pos3[3 * j], pos3[3 * j + 1], pos3[3 * j + 2]);
    m.vertices[j].normal = glm::vec3(nrm3[3 * j], nrm3[3 * j + 1], nrm3[3 * j + 2]);
  }
  if (tri) m.indices.assign(tri, tri + 3 * nfaces);
  return m;
}

}  // namespace

extern "C" int ref_host_nverts() { return REF_N; }

// Hair::setup -> init_simulation with srand(seed) standing in for app.cc:96-97.
// Outputs: 3 planes of S*REF_N float4 (pos, vel, tangent).
extern "C" void ref_host_init_simulation(const float* pos3, const float* nrm3, int64_t nstrands, unsigned seed,
                                         float maxlength, float* pos4, float* vel4, float* tan4) {
  MeshData m = make_mesh(pos3, nrm3, nstrands, nullptr, 0);
  HairRef h;
  h.params_.sim.maxlength = maxlength;
  h.nroots_ = (int)nstrands;                      // hair.cc:58
  srand(seed);
  h.init_simulation(m);
  const size_t bytes = (size_t)nstrands * REF_N * sizeof(glm::vec4);
  std::memcpy(pos4, h.captured_[0].data(), bytes);
  std::memcpy(vel4, h.captured_[1].data(), bytes);
  std::memcpy(tan4, h.captured_[2].data(), bytes);
}

// Hair::init_mesh element buffer: 6 * F * (REF_N - 1) ints.
extern "C" int64_t ref_host_patch_indices(const int32_t* tri, int64_t nfaces, int64_t nverts, int32_t* out) {
  std::vector<float> zeros((size_t)3 * nverts, 0.0f);
  MeshData m = make_mesh(zeros.data(), zeros.data(), nverts, tri, nfaces);
  HairRef h;
  h.nroots_ = (int)nverts;
  h.init_mesh(m);
  std::memcpy(out, h.captured_elements_.data(), h.captured_elements_.size() * sizeof(int));
  return (int64_t)h.captured_elements_.size();
}

extern "C" float ref_host_simplex2(float x, float y) { return glm::simplex(glm::vec2(x, y)); }

// ---- scalps without normals: RawMeshData::recalculateNormals (src/utils/raw_mesh_file.cc:11-50) -------------------------
// The members of RawMeshData (src/utils/raw_mesh_file.h:60-104) the function touches; its definition is the slice.
struct RawMeshData {
  std::vector<glm::vec3> vertices;
  std::vector<glm::vec3> normals;
  std::vector<glm::ivec3> elementsAttribs;
  int32_t nfaces() const { return static_cast<int32_t>(elementsAttribs.size() / 3); }
  void recalculateNormals();
};
#include "raw_recalc_normals.gen.inc"

// corner_v: 3 * nfaces zero-based position indices (what ParseOBJ leaves in elementsAttribs[].x after its [1, n] -> [0, n-1]
// pass, mesh_data_manager.cc:213-218). Out: one normal per corner, and the normal index the function gave each corner.
extern "C" void ref_host_recalc_normals(const float* pos3, int64_t nverts, const int32_t* corner_v, int64_t nfaces,
                                        float* nrm3_corner, int32_t* corner_n) {
  RawMeshData raw;
  raw.vertices.resize((size_t)nverts);
  for (int64_t i = 0; i < nverts; ++i) raw.vertices[i] = glm::vec3(pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2]);
  raw.elementsAttribs.resize((size_t)(3 * nfaces));
  for (int64_t q = 0; q < 3 * nfaces; ++q) raw.elementsAttribs[q] = glm::ivec3(corner_v[q], -1, -1);
  raw.recalculateNormals();
  for (int64_t q = 0; q < 3 * nfaces; ++q) {
    const int32_t n = raw.elementsAttribs[q].z;
    corner_n[q] = n;
    nrm3_corner[3 * q] = raw.normals[n].x; nrm3_corner[3 * q + 1] = raw.normals[n].y; nrm3_corner[3 * q + 2] = raw.normals[n].z;
  }
}

// ---- skinning palette: SkeletonController::generate_skinning_datas (src/fx/animation/skeleton_controller.cc:248-265) ------
// The members the function touches (skeleton_controller.h:50-56, skeleton.h:24-36, common.h:24); its definition is the slice.
#define LOOP_NTHREADS 4                            // skeleton_controller.cc:12
template<typename T> using JointBuffer_t = std::vector<T>;
enum class SkinningMode { LinearBlending, DualQuaternion };
struct Skeleton { JointBuffer_t<glm::mat4> inverse_bind_matrices; };
using SkeletonHandle = std::shared_ptr<Skeleton>;
struct SkeletonController {
  int32_t njoints_ = 0;
  JointBuffer_t<glm::mat4> global_pose_matrices_;
  JointBuffer_t<glm::mat3x4> skinning_matrices_;
  JointBuffer_t<glm::dualquat> dual_quaternions_;
  void generate_skinning_datas(SkinningMode const mode, SkeletonHandle skeleton);
};
#include "generate_skinning_datas.gen.inc"

// Matrices: GLM's memory layout (16 floats, column-major). Out: glm::dualquat per joint as it lies in memory — what
// the reference uploads into the skinning texture buffer: (real xyzw, dual xyzw).
extern "C" void ref_host_dq_palette(const float* global_pose16, const float* inverse_bind16, int njoints, float* dq8) {
  static_assert(sizeof(glm::dualquat) == 8 * sizeof(float), "dualquat is two quaternions");
  SkeletonController c;
  c.njoints_ = njoints;
  c.global_pose_matrices_.resize(njoints); c.skinning_matrices_.resize(njoints); c.dual_quaternions_.resize(njoints);
  auto skl = std::make_shared<Skeleton>();
  skl->inverse_bind_matrices.resize(njoints);
  std::memcpy(c.global_pose_matrices_.data(), global_pose16, sizeof(glm::mat4) * njoints);
  std::memcpy(skl->inverse_bind_matrices.data(), inverse_bind16, sizeof(glm::mat4) * njoints);
  c.generate_skinning_datas(SkinningMode::DualQuaternion, skl);
  std::memcpy(dq8, c.dual_quaternions_.data(), sizeof(glm::dualquat) * njoints);
}
