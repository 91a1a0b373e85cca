// barbu_hair.hpp — C++17 host adaptor with the call surface of the reference's `class Hair`
// (src/fx/hair.h:24-121) for the SIMULATION path, on top of the C ABI in barbu_hair.h.
//
// What a Barbü maintainer swaps in: `Hair::setup / update / set_bounding_sphere / initialized / deinit` keep their
// names, argument meaning and error behaviour (log + early return, never throw: hair.cc:45-48,90-93); the GL
// compute dispatch + PingPongBuffer::swap inside `update` (hair.cc:102-122) become one `bh_step`. `render` is not
// here: it stays the reference's GL code and keeps reading buffer 0 (position plane at byte offset 0, tangent plane
// at 2 * V * 16, stride 16, vertex = strand * N + i — hair.cc:371-389), either through `register_gl_buffer`
// (CUDA-GL interop, zero copy) or through `download`.
//
// Header-only; link with libbarbu_hair.so. No CUDA or GL headers needed by the including translation unit.
#ifndef BARBU_HAIR_HPP_
#define BARBU_HAIR_HPP_

#include <cstdint>
#include <cstdio>
#include <vector>

#include "barbu_hair.h"

namespace barbu {

// The members of MeshData the hair path reads (src/memory/resources/mesh_data.h:62-92): one root per scalp vertex
// (hair.cc:58), triangle LIST indices (hair.cc:401-409). Borrowed for the duration of setup(), like the resource.
struct ScalpMesh {
  const float* positions = nullptr;   // nvertices x 3
  const float* normals = nullptr;     // nvertices x 3
  std::int64_t nvertices = 0;
  const std::int32_t* indices = nullptr;  // nfaces x 3
  std::int64_t nfaces = 0;
  bool is_valid() const noexcept { return positions && normals && nvertices > 0; }
};

class Hair {
 public:
  // UI parameters of the reference (hair.h:27-51) that reach the simulation, plus what is a compile-time constant or
  // process state there and a run-time value here.
  struct Parameters_t {
    struct { float maxlength = 0.50f; } sim;                                  // hair.h:29
    struct { int ninstances = 3; int nlines = 2; int nsubsegments = 16; } tess;   // hair.h:33-35 (used by stream())
    struct { float lengthScale = 1.450f; } render;                            // hair.h:41 -> uScaleFactor (hair.cc:108)
    struct { int nroots = 0; int nControlPoints = 0; } readonly;              // hair.h:45-48
    struct {
      int ncontrol_points = 4;        // HAIR_MAX_PARTICLE_PER_STRAND (shaders/hair/interop.h:8)
      unsigned seed = 1234;           // stands in for srand(time(NULL)) (core/app.cc:96-97)
      int substeps = 1;               // 1 == reference (one dispatch per frame)
      int math = BH_MATH_EXACT;
      int device = 0;
      // More than one entry: the strand set is sharded contiguously over these CUDA devices (bh_group_*, one host thread,
      // no per-step exchange); the render GPU is devices[0]. Empty or one entry: the single-device path on `device`.
      std::vector<int> devices;
    } b200;
  };

  Hair() = default;
  Hair(const Hair&) = delete;
  Hair& operator=(const Hair&) = delete;
  ~Hair() { deinit(); }

  Parameters_t& params() noexcept { return params_; }
  const Parameters_t& params() const noexcept { return params_; }

  /* Initialize base resources (hair.cc:26-40). Nothing to compile or upload on this path: checks the library. */
  void init() { (void)bh_version(); }

  /* Release all allocated resources (hair.cc:66-87). */
  void deinit() {
    if (sim_) { bh_destroy(sim_); sim_ = nullptr; }
    if (group_) { bh_group_destroy(group_); group_ = nullptr; }
    patches_uploaded_ = false;
    nroots_ = 0;
    patch_indices_.clear();
  }

  /* Setup scalp specific resources (hair.cc:42-64): init_simulation + init_mesh. */
  void setup(ScalpMesh const& scalp) {
    if (!scalp.is_valid()) {
      log_error("The scalp mesh resource was not found.");                    // hair.cc:45-48
      return;
    }
    deinit();
    const int N = params_.b200.ncontrol_points;
    const std::int64_t S = scalp.nvertices;                                   // hair.cc:58
    if (params_.b200.devices.size() > 1) { setup_group(scalp, S, N); return; }
    if (params_.b200.devices.size() == 1) params_.b200.device = params_.b200.devices[0];
    if (!check(bh_create(&sim_, S, N, params_.b200.device), "bh_create")) { sim_ = nullptr; return; }
    bh_set_step_policy(sim_, BH_POLICY_AUTO);      // small scalps (the reference's own: 448 x 4) take the latency-oriented kernel
    // init_simulation (hair.cc:236-361): jitter on the host exactly as the reference, expansion on the device,
    // tangent plane on the host (libm + simplex noise), uploaded to plane 2.
    std::vector<float> rv(static_cast<size_t>(S));
    std::vector<float> tan(static_cast<size_t>(S) * N * 4);
    bool ok = check(bh_random_values(params_.b200.seed, 0, S, rv.data()), "bh_random_values") &&
              check(bh_init_strands(sim_, scalp.positions, scalp.normals, rv.data(), params_.sim.maxlength), "bh_init_strands") &&
              check(bh_init_tangents_host(scalp.normals, S, 0, S, N, params_.sim.maxlength, tan.data()), "bh_init_tangents_host") &&
              check(bh_upload(sim_, nullptr, nullptr, tan.data()), "bh_upload(tangents)");
    // init_mesh (hair.cc:363-417): the patch element buffer; the VAO stays with the GL side.
    if (ok && scalp.indices && scalp.nfaces > 0 && N > 1) {
      patch_indices_.resize(static_cast<size_t>(6) * scalp.nfaces * (N - 1));
      ok = check(bh_build_patch_indices(scalp.indices, scalp.nfaces, N, patch_indices_.data(), params_.b200.device),
                 "bh_build_patch_indices");
    }
    if (ok) ok = push_params();
    if (!ok) { deinit(); return; }
    nroots_ = static_cast<int>(S);
    params_.readonly.nroots = nroots_;                                        // hair.cc:60
    params_.readonly.nControlPoints = N;
  }

  /* Same from a scalp resource on disk (Application.cc:38-39 passes "models/InfiniteScan/Head_scalp.obj"): read by the
   * reference's OBJ rules; a missing or unusable file logs and leaves the module uninitialised. */
  void setup(const char* scalp_obj_path) {
    float *pos = nullptr, *nrm = nullptr; std::int32_t* tri = nullptr;
    ScalpMesh scalp;
    if (scalp_obj_path && bh_load_obj_scalp(scalp_obj_path, &pos, &nrm, &scalp.nvertices, &tri, &scalp.nfaces) == BH_OK) {
      scalp.positions = pos; scalp.normals = nrm; scalp.indices = tri;
    }
    setup(scalp);
    bh_free(pos); bh_free(nrm); bh_free(tri);
  }

  /* One simulation step (hair.cc:89-125). */
  void update(float const dt) {
    if (!initialized()) {
      log_debug("Calling Hair::update without initialization.");              // hair.cc:90-93
      return;
    }
    if (!push_params()) return;                                               // uniforms are re-sent every frame (hair.cc:107-110)
    if (group_) {
      if (check(bh_group_step(group_, dt, params_.b200.substeps), "bh_group_step") && group_gl_)
        check(bh_group_gather_to_gl(group_, 1u << BH_PLANE_POSITION), "bh_group_gather_to_gl");   // the render VAO reads positions (hair.cc:371-389)
      return;
    }
    check(bh_step(sim_, dt, params_.b200.substeps), "bh_step");
  }

  void set_bounding_sphere(const float (&bsphere)[4]) noexcept {              // hair.h:72-74
    for (int i = 0; i < 4; ++i) boundingsphere_[i] = bsphere[i];
    has_sphere_ = true;
  }

  bool initialized() const noexcept { return nroots_ != 0; }                  // hair.h:76-78

  // ---- buffer-binding contract of buffer 0 (pbuffer_.read_ssbo_id(), hair.cc:371-389) ----------------------------
  std::int64_t nvertices() const noexcept { return static_cast<std::int64_t>(nroots_) * params_.readonly.nControlPoints; }
  std::uint64_t position_plane_offset() const noexcept { return 0; }
  std::uint64_t tangent_plane_offset() const noexcept { return 2ull * static_cast<std::uint64_t>(nvertices()) * 16ull; }
  int patchsize() const noexcept { return 6; }                                // mesh_.patchsize, hair.cc:397
  const std::vector<std::int32_t>& patch_indices() const noexcept { return patch_indices_; }   // mesh_.ibo contents, nelems = size()

  /* Zero-copy: step straight into the GL buffer the render VAO is bound to (needs a current GL context). */
  bool register_gl_buffer(unsigned int gl_buffer) {
    if (group_) {                                                             // sharded: the GL buffer of the render GPU is the gather target
      group_gl_ = check(bh_group_register_gl_buffer(group_, gl_buffer, params_.b200.devices[0]), "bh_group_register_gl_buffer") &&
                  check(bh_group_gather_to_gl(group_, 7u), "bh_group_gather_to_gl");          // all three planes once; positions after every update
      return group_gl_;
    }
    return sim_ && check(bh_register_gl_buffer(sim_, gl_buffer), "bh_register_gl_buffer");
  }
  bool unregister_gl_buffer() {
    if (group_) { group_gl_ = false; return check(bh_group_unregister_gl_buffer(group_), "bh_group_unregister_gl_buffer"); }
    return sim_ && check(bh_unregister_gl_buffer(sim_), "bh_unregister_gl_buffer");
  }
  /* Copy path: planes to host memory (NULL = skip), V float4 each, global strand order whatever the sharding. */
  bool download(float* pos4, float* vel4, float* tan4) {
    if (group_) return check(bh_group_download(group_, pos4, vel4, tan4), "bh_group_download");
    return sim_ && check(bh_download(sim_, pos4, vel4, tan4), "bh_download");
  }
  /* Sharded setups: the position plane of the whole scalp on the render GPU (devices[0]), gathered by peer copies. */
  void* gather_positions(float* ms = nullptr) {
    void* ptr = nullptr;
    if (!group_ || !check(bh_group_gather_plane(group_, BH_PLANE_POSITION, params_.b200.devices[0], &ptr, ms), "bh_group_gather_plane")) return nullptr;
    return ptr;
  }

  /* The tess-stream half of Hair::render (hair.cc:141-173): the interpolated render strands as the GL_LINES vertex stream
   * (xyz, relPos; two float4 per sub-segment) that the reference captures by transform feedback and draws with
   * glDrawTransformFeedback, for params().tess. Returns the number of float4 (0 on failure); they stay on the device
   * (bh_tess_device_buffer) and are also copied to `out4_host` when it is not null (room for stream_count() float4). */
  std::int64_t stream_count() {
    if (group_) { log_error("Hair::stream: not available on a sharded setup (gather the planes to one device first)."); return 0; }
    if (!ensure_patches()) return 0;
    const bh_tess_params t = tess_params();
    const std::int64_t n = bh_tess_stream_count(sim_, &t);
    return n > 0 ? n : 0;
  }
  std::int64_t stream(float* out4_host = nullptr) {
    if (group_) { log_error("Hair::stream: not available on a sharded setup (gather the planes to one device first)."); return 0; }
    if (!ensure_patches()) return 0;
    const bh_tess_params t = tess_params();
    if (!check(bh_tess_stream(sim_, &t, out4_host), "bh_tess_stream")) return 0;
    return stream_count();
  }

  /* Strand state to / from a BARBUHS1 file (the reference has no persistence; include/barbu_hair.h describes the format). */
  bool save_state(const char* path) { return sim_ && check(bh_save_state(sim_, path, nullptr), "bh_save_state"); }
  bool load_state(const char* path) { return sim_ && check(bh_load_state(sim_, path, nullptr), "bh_load_state"); }

  bh_sim* handle() noexcept { return sim_; }
  bh_group* group_handle() noexcept { return group_; }

 private:
  void setup_group(ScalpMesh const& scalp, std::int64_t S, int N) {
    const std::vector<int>& devs = params_.b200.devices;
    if (!check(bh_group_create(&group_, devs.data(), static_cast<int>(devs.size()), S, N), "bh_group_create")) { group_ = nullptr; return; }
    std::vector<float> rv(static_cast<size_t>(S));
    std::vector<float> tan(static_cast<size_t>(S) * N * 4);
    bool ok = check(bh_random_values(params_.b200.seed, 0, S, rv.data()), "bh_random_values") &&
              check(bh_group_init_strands(group_, scalp.positions, scalp.normals, rv.data(), params_.sim.maxlength), "bh_group_init_strands") &&
              check(bh_init_tangents_host(scalp.normals, S, 0, S, N, params_.sim.maxlength, tan.data()), "bh_init_tangents_host") &&
              check(bh_group_upload(group_, nullptr, nullptr, tan.data()), "bh_group_upload(tangents)");
    if (ok && scalp.indices && scalp.nfaces > 0 && N > 1) {
      patch_indices_.resize(static_cast<size_t>(6) * scalp.nfaces * (N - 1));
      ok = check(bh_build_patch_indices(scalp.indices, scalp.nfaces, N, patch_indices_.data(), devs[0]), "bh_build_patch_indices");
    }
    if (ok) ok = push_params();
    if (!ok) { deinit(); return; }
    nroots_ = static_cast<int>(S);
    params_.readonly.nroots = nroots_;
    params_.readonly.nControlPoints = N;
  }
  bool push_params() {
    bh_params p;
    if (!check(bh_get_params(group_ ? bh_group_shard(group_, 0) : sim_, &p), "bh_get_params")) return false;
    p.scale = params_.render.lengthScale;
    p.math = params_.b200.math;
    // The reference sends its boundingsphere_ member every frame even when no collider exists (latent UB:
    // hair.h:99 is uninitialised); here the shader default (0,0,0,1) (cs_simulation.glsl:43) stands until one is set.
    if (has_sphere_) for (int i = 0; i < 4; ++i) p.sphere[i] = boundingsphere_[i];
    if (group_) return check(bh_group_set_params(group_, &p), "bh_group_set_params");
    return check(bh_set_params(sim_, &p), "bh_set_params");
  }
  bool ensure_patches() {                                                        // the element buffer of init_mesh, on the device
    if (!initialized()) { log_debug("Hair::stream called before setup()."); return false; }   // like render(), hair.cc:128-131
    if (patch_indices_.empty()) { log_error("Hair::stream: the scalp had no faces."); return false; }
    if (!patches_uploaded_) {
      if (!check(bh_tess_set_patches(sim_, patch_indices_.data(), static_cast<std::int64_t>(patch_indices_.size())), "bh_tess_set_patches")) return false;
      patches_uploaded_ = true;
    }
    return true;
  }
  bh_tess_params tess_params() const noexcept {
    bh_tess_params t;
    t.ninstances = params_.tess.ninstances; t.nlines = params_.tess.nlines; t.nsubsegments = params_.tess.nsubsegments;
    t.seed = params_.b200.seed;
    return t;
  }
  static bool check(int rc, const char* what) {
    if (rc == BH_OK) return true;
    std::fprintf(stderr, "[barbu::Hair] %s failed (%d): %s\n", what, rc, bh_last_error());
    return false;
  }
  static void log_error(const char* msg) { std::fprintf(stderr, "[barbu::Hair] ERROR %s\n", msg); }
  static void log_debug(const char* msg) {
#ifdef BARBU_ENABLE_DEBUG_LOG
    std::fprintf(stderr, "[barbu::Hair] %s\n", msg);
#else
    (void)msg;
#endif
  }

  Parameters_t params_;
  int nroots_ = 0;                         //< Number of strands / root vertices in the scalp.
  bh_sim* sim_ = nullptr;                  //< Replaces PingPongBuffer pbuffer_ + the cs_simulation program.
  bh_group* group_ = nullptr;              //< ... or its sharded form (params().b200.devices.size() > 1); exactly one of the two is set.
  bool group_gl_ = false;
  float boundingsphere_[4] = { 0.f, 0.f, 0.f, 1.f };
  bool has_sphere_ = false;
  std::vector<std::int32_t> patch_indices_;
  bool patches_uploaded_ = false;
};

}  // namespace barbu

#endif  // BARBU_HAIR_HPP_
