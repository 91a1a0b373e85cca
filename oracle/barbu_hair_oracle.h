/*
 * barbu_hair_oracle.h — CPU ORACLE for the hair-strand simulation hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke check in
 * __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product (barbu_b200/, include/barbu_hair.h) never links, imports or calls it.
 *
 * It restates, in scalar C with a fully specified IEEE-754 fp32 operation order, what the
 * reference computes on this path:
 *   - src/shaders/hair/01_simulation/cs_simulation.glsl:37-208   (one simulation step)
 *   - src/fx/hair.cc:236-328                                      (strand state generation)
 *   - src/fx/hair.cc:397-409                                      (tess patch indices)
 *   - third_party/glm (0.9.9.9) glm::simplex(vec2)                (tangent noise)
 *
 * Arithmetic profile "glm-strict": every GLSL built-in is evaluated the way the reference's
 * own vendored GLM evaluates it on the host (dot = (x*x + y*y) + z*z, normalize = v * (1/sqrt(dot)),
 * reflect = I - (N*dot(N,I))*2, fma = one fused rounding), nothing else is contracted.
 * Parity pin: oracle/_ref (the reference shader SOURCE compiled as C++ against the reference's
 * GLM, built by oracle/Makefile) is bit-identical to this file on the seeded cases in
 * tests/golden/ — see tests/test_oracle_vs_ref.py.
 */
#ifndef BARBU_HAIR_ORACLE_H_
#define BARBU_HAIR_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BHO_MAX_COLLIDERS 8

/* Collider kinds for the extension path (a sphere is a capsule with a == b). */
typedef struct bho_capsule {
  float a[3];   /* segment end A (sphere: centre) */
  float b[3];   /* segment end B (sphere: == a)   */
  float radius;
} bho_capsule;

typedef struct bho_params {
  /* --- reference uniforms / constants (cs_simulation.glsl:37-43, 71-74, 102, 197) --- */
  float dt;             /* uTimeStep */
  float scale;          /* uScaleFactor */
  float sphere[4];      /* uBoundingSphere: xyz centre, w radius */
  int   iterations;     /* kNumContraintsIteration = 8 */
  float gravity[3];     /* gravity = (0, -9.81, 0) */
  float force_coeff;    /* kForceCoeff = 20 */
  float damp;           /* dftl_damp_scale as evaluated by lane 0 = 0.80 */
  /* --- extensions, NO reference implementation; all off when zero ------------------- */
  float wind[3];        /* additive constant force (same units as force_coeff*gravity) */
  float drag;           /* velocity (= displacement) multiplier (1 - drag) at integration */
  int   ncapsules;      /* extra capsule colliders applied after the sphere, in order */
  bho_capsule capsules[BHO_MAX_COLLIDERS];
} bho_params;

void bho_default_params(bho_params* p);

/* One reference step, READ -> WRITE -> swap folded into an in-place update of the two planes.
 * pos4/vel4: S*N float4 each, vertex index = strand*N + i (hair.cc:257,279). */
void bho_step(float* pos4, float* vel4, int64_t nstrands, int nverts, const bho_params* p);
/* Same, OpenMP over strands with `nthreads` threads (cpu baseline). */
void bho_step_mt(float* pos4, float* vel4, int64_t nstrands, int nverts, const bho_params* p, int nthreads);

/* hair.cc:273-275 — srand(seed) then one rand() per strand, evaluated in double, cast to float. */
void bho_random_values(unsigned seed, int64_t nstrands, float* out);

/* hair.cc:255-287 — positions (xyz + rest length in w) and zero velocities. */
void bho_init_strands(const float* root_pos3, const float* root_nrm3, const float* random_value,
                      int64_t nstrands, int nverts, float maxlength, float* pos4, float* vel4);

/* hair.cc:290-328 — tangent plane (render-side data, untouched by the simulation). */
void bho_init_tangents(const float* root_nrm3, int64_t nstrands, int nverts, float maxlength, float* tan4);

/* glm::simplex(vec2) restated (third_party/glm/glm/gtc/noise.inl). */
float bho_simplex2(float x, float y);

/* hair.cc:397-409 — 6 * F * (N-1) int32 element indices. Returns 0, or -1 on int32 overflow. */
int bho_patch_indices(const int32_t* tri_indices, int64_t nfaces, int nverts, int32_t* out);

/* Synthetic pole-free lat-long unit-sphere scalp: R*C vertices, 2*(R-1)*C triangles. */
void bho_sphere_scalp(int rows, int cols, float* pos3, float* nrm3, int32_t* tri_indices);

/* Extension (no reference): dual-quaternion skinning of root positions, following the formula of
 * src/shaders/shared/inc_skinning.glsl:22-31,54-82. dq palette: njoints * 8 floats (real xyzw, dual xyzw). */
void bho_skin_roots_dq(const float* rest_pos3, const float* rest_nrm3, const int32_t* joints4,
                       const float* weights3, const float* dq_palette, int64_t nstrands,
                       float* out_pos3, float* out_nrm3);

/* FNV-1a 64 over raw bytes — "checksum of checksums" helper for full-size property tests. */
/* RawMeshData::recalculateNormals (raw_mesh_file.cc:11-50): one normal per face corner for a scalp that came without. */
void bho_recalc_normals(const float* pos3, int64_t nverts, const int32_t* corner_v, int64_t nfaces, float* nrm3_corner);
/* generate_skinning_datas (skeleton_controller.cc:248-265): global pose x inverse bind -> glm::dualquat(mat3x4) per joint. */
void bho_dq_palette_from_matrices(const float* global_pose16, const float* inverse_bind16, int njoints, float* dq8);
uint64_t bho_fnv1a64(const void* data, uint64_t nbytes);

/* Tess-stream stage (src/shaders/hair/02_tess_stream, hair.cc:141-173): interpolated render strands as GL_LINES
 * vertices (xyz, relPos). out4: bho_tess_stream_count(...) float4. See the .c file for what is reference text and what
 * is defined here (tess coordinates, primitive order, the seeded random pair). */
void bho_tess_random_pair(uint32_t seed, int index, float st[2]);
int64_t bho_tess_stream_count(int64_t npatches, int ninstances, int nlines, int nsubsegments);
void bho_tess_stream(const float* pos4, const float* tan4, const int32_t* patch_indices, int64_t npatches, int nverts,
                     float scale, int ninstances, int nlines, int nsubsegments, uint32_t seed, float* out4);

/* Marschner lookup tables (barbu_marschner_oracle.c; cs_marschner_m.glsl / cs_marschner_n.glsl): fp32 RGBA texels. */
void bho_marschner_luts(const float* params9, int resolution, float* m_rgba, float* n_rgba);
void bho_float_to_half(const float* in, int64_t count, uint16_t* out);     /* GL_RGBA16F store, round to nearest even */

#ifdef __cplusplus
}
#endif
#endif /* BARBU_HAIR_ORACLE_H_ */
