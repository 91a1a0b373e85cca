/*
 * barbu_hair.h — C ABI of the B200-native hair-strand simulation (libbarbu_hair.so).
 *
 * Drop-in boundary for ONE path of tcoppex/barbu: what `class Hair` (src/fx/hair.h:24-121) does
 * between `setup()` and the position buffer the tessellation/Marschner render path reads.
 * Plain pointers and sizes only; no C++/torch types. Every entry point returns 0 (BH_OK) or an
 * error code, never throws or aborts (the reference logs and returns, hair.cc:45-48,90-93).
 * The implementation is CUDA for sm_100a only: there is no CPU fallback, and every call fails with
 * BH_ERR_CUDA when no device is usable.
 *
 * Buffer contract (src/memory/pingpong_buffer.cc:16-17,44-48; src/fx/hair.cc:371-389):
 *   one device buffer ("buffer 0" == PingPongBuffer::read_ssbo_id()) of 3 SoA planes of float4,
 *   plane p at byte offset p * V * 16, V = nstrands * nverts, vertex index = strand * nverts + i:
 *     plane 0  position.xyz + rest length      (SSBO_HAIR_SIM_POSITION_READ, interop.h:19)
 *     plane 1  velocity.xyz + 0                (SSBO_HAIR_SIM_VELOCITY_READ, interop.h:20)
 *     plane 2  tangent.xyz  + 0                (SSBO_HAIR_SIM_TANGENT_READ,  interop.h:21)
 *   bh_step updates planes 0 and 1 IN PLACE (each strand is owned by one thread), so the
 *   reference's WRITE buffer and its per-frame swap() copy (pingpong_buffer.cc:73-84) do not exist.
 */
#ifndef BARBU_HAIR_H_
#define BARBU_HAIR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bh_sim bh_sim;   /* opaque; owns device memory, a stream and (optionally) a GL mapping */

enum {
  BH_OK = 0,
  BH_ERR_INVALID = 1,        /* bad argument */
  BH_ERR_CUDA = 2,           /* CUDA runtime error / no device; see bh_last_error() */
  BH_ERR_NOT_INITIALIZED = 3,/* step before any state was uploaded/generated (Hair::initialized(), hair.h:76-78) */
  BH_ERR_UNSUPPORTED = 4,
  BH_ERR_OVERFLOW = 5        /* an int32 output would overflow (hair.cc:398 `int nelems`) */
};

enum { BH_PLANE_POSITION = 0, BH_PLANE_VELOCITY = 1, BH_PLANE_TANGENT = 2, BH_NUM_PLANES = 3 };

/* Arithmetic profile of the step kernel.
 * BH_MATH_EXACT: IEEE-754 binary32 operation sequence of oracle/barbu_hair_oracle.c (bit-exact).
 * BH_MATH_FAST : FMA-contracted, MUFU.RSQ-based normalisation — the arithmetic a GPU GLSL compiler
 *                emits for the reference shader; <= 1e-5 relative per vertex after one step. */
enum { BH_MATH_EXACT = 0, BH_MATH_FAST = 1 };

#define BH_MAX_CAPSULES 8
typedef struct bh_capsule { float a[3]; float b[3]; float radius; } bh_capsule;

typedef struct bh_params {
  /* reference uniforms and shader constants (cs_simulation.glsl:37-43,71-74,102,197) */
  float scale;          /* uScaleFactor  <- Hair::Parameters_t::render.lengthScale = 1.45 (hair.cc:108) */
  float sphere[4];      /* uBoundingSphere (xyz centre, w radius), default (0,0,0,1) (cs:43) */
  int   iterations;     /* kNumContraintsIteration = 8 */
  float gravity[3];     /* (0, -9.81, 0) */
  float force_coeff;    /* kForceCoeff = 20 */
  float damp;           /* dftl_damp_scale as lane 0 evaluates it = 0.80 */
  int   math;           /* BH_MATH_EXACT (default) or BH_MATH_FAST */
  /* extensions with NO reference implementation; zero = off = reference behaviour */
  float wind[3];        /* constant force added to force_coeff * gravity */
  float drag;           /* integration uses velocity * (1 - drag) */
  int   ncapsules;      /* capsule colliders applied after the sphere, in order */
  bh_capsule capsules[BH_MAX_CAPSULES];
} bh_params;

/* ---- lifetime: PingPongBuffer::setup / destroy (pingpong_buffer.cc:6-33) -------------------- */
int  bh_create(bh_sim** out, int64_t nstrands, int nverts, int device);
int  bh_destroy(bh_sim* sim);
/* Launch on a caller-owned cudaStream_t (e.g. the host framework's current stream) instead of the
 * sim's own non-blocking stream. NULL is the CUDA legacy default stream, as everywhere in CUDA. */
int  bh_set_stream(bh_sim* sim, void* cuda_stream);
int  bh_reset_stream(bh_sim* sim);                    /* back to the sim's own stream */
int  bh_synchronize(bh_sim* sim);

/* ---- parameters: uniforms of Hair::update (hair.cc:107-110), Hair::set_bounding_sphere -------- */
void bh_default_params(bh_params* p);
int  bh_set_params(bh_sim* sim, const bh_params* p);
int  bh_get_params(const bh_sim* sim, bh_params* p);
int  bh_set_bounding_sphere(bh_sim* sim, const float sphere[4]);          /* hair.h:72-74 */

/* ---- state transfer: glNamedBufferSubData per plane (hair.cc:336-340); NULL planes are skipped -- */
int  bh_upload(bh_sim* sim, const float* pos4, const float* vel4, const float* tan4);
int  bh_download(bh_sim* sim, float* pos4, float* vel4, float* tan4);
/* Device address of plane p of buffer 0 (for zero-copy consumers: NCCL all-gather, CUDA renderers). */
int  bh_device_plane(bh_sim* sim, int plane, void** device_ptr, uint64_t* nbytes);

/* ---- strand generation: Hair::init_simulation (hair.cc:236-361) on the device --------------- */
/* hair.cc:273-275: srand(seed); value j = float(1.0 + 0.1*(1.0 - 2.0*rand()/RAND_MAX)) for global
 * strand j. Writes values [first, first+count). Host glibc rand(), exactly as the reference. */
int  bh_random_values(unsigned seed, int64_t first, int64_t count, float* out);
/* hair.cc:255-287 from caller-provided roots (scalp vertex position + normal, one per strand). */
int  bh_init_strands(bh_sim* sim, const float* root_pos3, const float* root_nrm3,
                     const float* random_value, float maxlength);
/* Synthetic unit-sphere scalp of rows*cols strands; this sim holds global strands
 * [first, first+nstrands). Roots are generated on the device, then expanded as bh_init_strands. */
int  bh_init_sphere_scalp(bh_sim* sim, int rows, int cols, int64_t first,
                          const float* random_value, float maxlength);
/* Same scalp with a chosen STRAND ORDER (= vertex numbering of the scalp mesh; the reference takes strands in mesh vertex
 * order, hair.cc:255-287). BH_SCALP_ROW_MAJOR: strand g = r * cols + c, one latitude circle after the other (the plain
 * calls above). BH_SCALP_COLUMN_MAJOR: strand g = c * rows + r, one meridian after the other — a contiguous strand range is
 * then a longitude wedge that holds every latitude, so contiguous shards carry the same collider-contact
 * load instead of one rank owning the pole where all hair lies on the collider. random_value[j] belongs to strand
 * first + j in the chosen order (bh_random_values(seed, first, count): one rand() per strand in strand order). */
enum { BH_SCALP_ROW_MAJOR = 0, BH_SCALP_COLUMN_MAJOR = 1 };
int  bh_init_sphere_scalp_ordered(bh_sim* sim, int rows, int cols, int order, int64_t first,
                                  const float* random_value, float maxlength);
int  bh_sphere_scalp_triangles_ordered(int rows, int cols, int order, int32_t* tri_indices);
/* hair.cc:290-328: tangent plane, host-evaluated (libm sinf/cosf + glm::simplex restated), for
 * global strands [first, first+count) out of `total`. Output: count*nverts float4. */
int  bh_init_tangents_host(const float* root_nrm3, int64_t total, int64_t first, int64_t count,
                           int nverts, float maxlength, float* tan4);
/* Triangle list of the synthetic sphere scalp: 2*(rows-1)*cols triangles, int32 x 3. */
int  bh_sphere_scalp_triangles(int rows, int cols, int32_t* tri_indices);

/* ---- scalp input: the reference's OBJ reading rules (mesh_data_manager.cc:69-223) and vertex re-indexing
 * (mesh_data.cc:384-406): one root per unique (v, vt, vn) corner triple in first-appearance order, quads split as
 * (x, y, z), (z, w, x). A file without `vn` lines gets the normals RawMeshData::recalculateNormals would make
 * (raw_mesh_file.cc:11-50): one per face corner, hence 3 roots per triangle. As in the reference a last line that does not
 * end in '\n' is not read. Outputs are malloc'ed: release each with bh_free. Host only, no device needed. */
int  bh_load_obj_scalp(const char* path, float** root_pos3, float** root_nrm3, int64_t* nvertices,
                       int32_t** tri_indices, int64_t* nfaces);
void bh_free(void* ptr);

/* ---- Hair::init_mesh element buffer (hair.cc:397-409), computed on the device --------------- */
/* out: 6 * nfaces * (nverts-1) int32, host memory. */
int  bh_build_patch_indices(const int32_t* tri_indices, int64_t nfaces, int nverts, int32_t* out, int device);

/* ---- Hair::update(dt) (hair.cc:89-125) ------------------------------------------------------- */
/* `substeps` launches of the fused step kernel, each with dt/substeps (substeps = 1: reference). */
int  bh_step(bh_sim* sim, float dt, int substeps);
/* Which step kernel serves bh_step. BH_POLICY_THROUGHPUT (default): the streaming kernel — fewest instructions and HBM bytes
 * per vertex, but a launch never takes less than 10-14 us (one warp's pipeline latency). BH_POLICY_LATENCY: the wavefront
 * kernel — eight lanes per strand, one constraint iteration each; about twice the instructions, a quarter of the latency;
 * the substeps of a bh_step run as passes of its one launch. BH_POLICY_AUTO: LATENCY up to 2^18 vertices (the reference's
 * own scalp is 448 x 4), THROUGHPUT above. Same arithmetic, bit-identical results (8 iterations; other counts keep the
 * generic kernel). The Hair adaptors (barbu_hair.hpp, barbu_b200.Hair) set AUTO. */
#define BH_POLICY_THROUGHPUT 0
#define BH_POLICY_LATENCY    1
#define BH_POLICY_AUTO       2
int  bh_set_step_policy(bh_sim* sim, int policy);
/* Frame-level substep fusion (off by default = one launch per substep, the reference's one dispatch per step). When on,
 * bh_step(dt, substeps > 1) runs the substeps of the frame as consecutive PASSES of one launch of the streaming kernel: a warp
 * takes a group of tiles through all passes before it asks for more, so a pass re-reads from L2 what the previous one
 * stored, and HBM carries 64 B per vertex per frame instead of per substep. Per strand the operations and their order are
 * those of `substeps` separate launches: results are bit-identical. Shapes the streaming kernel does not take (nverts = 1,
 * iteration counts other than 8, fewer tiles than one group) silently keep one launch per substep.
 * enabled: 0 off; 1 fuse where it pays (a group of tiles runs all its passes on one warp, so small shards — fewer than 1024
 * groups of more than one tile, e.g. the reference's own 448 x 4 scalp — keep their launches and their parallelism);
 * 2 fuse wherever the shape allows it (tests). */
int  bh_set_substep_fusion(bh_sim* sim, int enabled);
/* Same through HOST buffers: upload pos/vel, step, download pos/vel; copies are chunked and
 * overlapped with the kernels on internal streams. Buffers should be page-locked (bh_host_alloc). */
int  bh_step_host(bh_sim* sim, float dt, int substeps, float* pos4, float* vel4);
/* The reference's own frame with the state resident on the device: Hair::update(dt) takes no buffers (hair.cc:89-125), the
 * host sends uniforms (bh_set_bounding_sphere) and a host-side consumer reads the position plane of buffer 0. One call:
 * `substeps` steps, then plane 0 (V float4) in pos4; the shard is stepped slice by slice and the device->host copy of a
 * slice overlaps the step of the next. Returns when pos4 is complete. pos4 should be page-locked (bh_host_alloc). */
int  bh_step_readback(bh_sim* sim, float dt, int substeps, float* pos4);
int  bh_host_alloc(void** ptr, uint64_t nbytes);     /* cudaHostAlloc, portable across the devices of a group */
int  bh_host_free(void* ptr);
/* Kernel launches issued by this sim so far (bench.py's gpu_launches claim). */
int64_t bh_launch_count(const bh_sim* sim);
/* Which step kernel the current shape/parameters select: 0 streaming (TMA tiles + packed fp32x2; any nverts >= 2,
 * 8 iterations, sphere + capsules), 1 per-strand pipelined (nverts = 1, corner cases), 2 generic (any iteration count),
 * 3 latency-oriented wavefront (bh_set_step_policy). */
int  bh_step_kernel_kind(const bh_sim* sim);
/* Exhaustive device check of the exact profile's branch-free 1/sqrt(x) (scalar and packed fp32x2 forms) against
 * the IEEE-754 builtins over every finite binary32 >= 2^-102. *mismatches must come back 0. */
int  bh_selftest_math(int device, uint64_t* mismatches);

/* ---- extension: dual-quaternion skinned roots (formula of shared/inc_skinning.glsl:22-31,54-82) */
/* Stores rest roots + skin data once; bh_skin_roots rewrites vertex 0 of every strand in plane 0. Joint indices are
 * validated: negative ones are refused by bh_set_skin, ones >= njoints by bh_skin_roots (BH_ERR_INVALID). */
int  bh_set_skin(bh_sim* sim, const float* rest_root_pos3, const int32_t* joints4, const float* weights3);
int  bh_skin_roots(bh_sim* sim, const float* dq_palette, int njoints);
/* The palette as the reference's animation system makes it (SkeletonController::generate_skinning_datas,
 * src/fx/animation/skeleton_controller.cc:248-265): per joint global_pose * inverse_bind, its top three rows as the 3x4
 * skinning matrix, glm::dualquat of that. Matrices in GLM's layout: 16 floats, column-major. dq_palette: njoints x 8 floats
 * (real xyzw, dual xyzw) — the argument of bh_skin_roots. Host only, bit-identical to the reference's host code. */
int  bh_dq_palette_from_matrices(const float* global_pose16, const float* inverse_bind16, int njoints, float* dq_palette);

/* ---- next stage: the tess-stream pass of Hair::render (hair.cc:141-173) on the device ------------ */
/* Interpolated render strands as the GL_LINES vertex stream (xyz, relPos) that glDrawTransformFeedback consumes:
 * per patch (6 control points, bh_build_patch_indices order), instance, isoline and sub-segment two float4.
 * Formulas of shaders/hair/02_tess_stream/ (all four stages) + shared/inc_maths.glsl (hermite_mix, sample_triangle2, smoothstep2);
 * tess coordinates, primitive order and the (seeded, counter-based) random pair are defined by this library because
 * the reference leaves them to the GL implementation / to std::random_device: no reference parity is claimed there. */
typedef struct bh_tess_params { int ninstances; int nlines; int nsubsegments; unsigned seed; } bh_tess_params;   /* hair.h:32-36 */
int     bh_tess_set_patches(bh_sim* sim, const int32_t* patch_indices, int64_t nelems);   /* upload once per scalp */
int64_t bh_tess_stream_count(const bh_sim* sim, const bh_tess_params* t);                  /* float4 per call, -1 on bad params */
int     bh_tess_stream(bh_sim* sim, const bh_tess_params* t, float* out4_host /* NULL: leave it on the device */);
int     bh_tess_device_buffer(bh_sim* sim, void** device_ptr, int64_t* count);

/* ---- next stage: strand-state files and device checksums -------------------------------------------- */
/* The reference has no persistence (state lives in GL buffers, regenerated from rand() at start-up, hair.cc:236-361).
 * A state file is the planes of buffer 0, byte for byte in PingPongBuffer order (pingpong_buffer.cc:16-17,44-48), behind
 * a 512-byte little-endian header:
 *     0  char[8] "BARBUHS1"      8  u32 version (1)      12  u32 header_bytes (512)
 *    16  i64 nstrands (this file)                        24  i32 nverts       28  u32 plane_mask (bit p: plane p stored)
 *    32  i64 total_strands       40  i64 first_strand (shard coordinates)        48  i64 frame
 *    56  f32 dt                  60  u32 seed (srand seed of the length jitter, hair.cc:273-275)
 *    64  u64 checksum[2]         80  u32 params_bytes (292)   84  u32 reserved   88  bh_params   ...zero padding to 512
 *   512  the stored planes in ascending order, nstrands * nverts * 16 bytes each
 * checksum (computed on the device, one streaming pass): over the 32-bit words w of the stored planes, with
 * key = (plane << 60) + 4 * global_vertex + component and G = 0x9E3779B97F4A7C15:
 *     checksum[0] = sum w,    checksum[1] = sum w * ((key + 1) * G)          (both mod 2^64)
 * Keys use GLOBAL vertex indices ((first_strand + strand) * nverts + i), so the checksum of a scalp equals the wrapping
 * sum of the checksums of its shards. */
typedef struct bh_state_info {
  int64_t  nstrands;      /* out */
  int      nverts;        /* out */
  unsigned plane_mask;    /* in (0 = all three planes) / out */
  int64_t  total_strands; /* in (0 = nstrands) / out */
  int64_t  first_strand;  /* in / out */
  int64_t  frame;         /* in / out: caller's frame counter */
  float    dt;            /* in / out: caller's frame time step */
  unsigned seed;          /* in / out */
  uint64_t checksum[2];   /* out */
  bh_params params;       /* out: the sim's parameters when the file was written */
} bh_state_info;
int  bh_state_checksum(bh_sim* sim, unsigned plane_mask, int64_t first_strand, uint64_t out[2]);
int  bh_save_state(bh_sim* sim, const char* path, const bh_state_info* info /* NULL: unsharded, frame 0 */);
/* Header only; host only (no device needed): learn the shape before bh_create. */
int  bh_peek_state(const char* path, bh_state_info* info);
/* The sim must have the file's shape. Loads the stored planes, verifies the checksum on the device
 * (BH_ERR_INVALID on mismatch: the sim then holds no valid state) and adopts the stored parameters. */
int  bh_load_state(bh_sim* sim, const char* path, bh_state_info* info /* may be NULL */);

/* ---- next stage: the Marschner lookup tables of the render path ------------------------------------- */
/* Marschner::generate (src/fx/marschner.cc:35-69): cs_marschner_m.glsl + cs_marschner_n.glsl over a resolution^2 image
 * (kTextureResolution = 128, marschner.h:28), GL_RGBA16F texels, texel (x, y) at index y * resolution + x.
 *   M: (M_R, M_TT, M_TRT, cos theta_d) at (sin theta_i, sin theta_r) = 2 (x, y) / resolution - 1
 *   N: (N_R, N_TT, N_TRT, 1)          at (cos phi_d,  cos theta_d)  = 2 (x, y) / resolution - 1
 * Field defaults: Marschner::ShadingParameters_t (marschner.h:38-52). Any output pointer may be NULL; the rgba32f
 * outputs are the texels before the half-float store. Transcendental functions are CUDA's: a few ulp from a GL driver's. */
typedef struct bh_marschner_params {
  float eta, absorption, eccentricity;                 /* fiber properties  -> uEta, uAbsorption, uEccentricity */
  float ar, br;                                        /* surface           -> uLongitudinalShift, uLongitudinalWidth */
  float glint_scale, azimuthal_width, delta_caustic, delta_hm;   /* glints (unused by the shaders' active branch) */
} bh_marschner_params;
void bh_marschner_default_params(bh_marschner_params* p);
int  bh_marschner_generate(const bh_marschner_params* p, int resolution, int device, uint16_t* m_rgba16f,
                           uint16_t* n_rgba16f, float* m_rgba32f, float* n_rgba32f);

/* ---- one scalp over several GPUs, behind one handle, from ONE host thread (SURVEY.md §8b "Threading", §8e) ------------- */
/* The reference's host is one process and one thread (src/core/app.cc:60-81, core/renderer.cc:69-81): it cannot use ranks.
 * A group shards the strand set contiguously — shard g = global strands [S*g/G, S*(g+1)/G) on devices[g] (a device may be
 * listed more than once) — and steps all shards with asynchronous launches on their own streams; no step exchanges data
 * (strands are independent, cs_simulation.glsl:170-208). Global planes passed to upload / download are in global strand
 * order (vertex = strand * nverts + i), exactly the single-device layout. The ONE exchange is optional: gathering a plane
 * of buffer 0 onto the render GPU, pushed by every source GPU with peer copies over NVLink / NVSwitch, so that the render
 * VAO (hair.cc:371-389) sees the whole scalp. bh_group_shard() lends a shard's bh_sim for every per-shard call above. */
typedef struct bh_group bh_group;
int     bh_group_create(bh_group** out, const int* devices, int ndevices, int64_t nstrands, int nverts);
int     bh_group_destroy(bh_group* group);
int     bh_group_size(const bh_group* group);
bh_sim* bh_group_shard(bh_group* group, int shard);                                   /* borrowed; NULL when out of range */
int     bh_group_shard_range(const bh_group* group, int shard, int64_t* first, int64_t* count);
int     bh_group_set_params(bh_group* group, const bh_params* p);
int     bh_group_set_bounding_sphere(bh_group* group, const float sphere[4]);         /* Hair::set_bounding_sphere for every shard */
int     bh_group_init_sphere_scalp(bh_group* group, int rows, int cols, int order, unsigned seed, float maxlength);
int     bh_group_init_strands(bh_group* group, const float* root_pos3, const float* root_nrm3, const float* random_value,
                              float maxlength);                                       /* global arrays, one entry per strand */
int     bh_group_upload(bh_group* group, const float* pos4, const float* vel4, const float* tan4);   /* global planes */
int     bh_group_download(bh_group* group, float* pos4, float* vel4, float* tan4);
int     bh_group_step(bh_group* group, float dt, int substeps);                       /* Hair::update(dt); returns after the launches */
int     bh_group_set_substep_fusion(bh_group* group, int enabled);                    /* bh_set_substep_fusion on every shard */
int     bh_group_synchronize(bh_group* group);
/* `frames` x bh_group_step between per-device CUDA events: ms_per_shard[g] (may be NULL) and their maximum — the frame
 * time of the job, measured on the devices. */
int     bh_group_step_timed(bh_group* group, float dt, int substeps, int frames, float* ms_max, float* ms_per_shard);
int64_t bh_group_launch_count(const bh_group* group);
/* Gather plane `plane` of every shard into one buffer of V float4 on dst_device (owned by the group, PingPongBuffer layout:
 * *device_ptr = base of that plane). Ordered after the steps already queued; returns when the buffer is complete.
 * ms (may be NULL): max over shards of the copy's device time. */
int     bh_group_gather_plane(bh_group* group, int plane, int dst_device, void** device_ptr, float* ms);
/* The render GPU's GL buffer 0 (pbuffer_.read_ssbo_id(), 3 planes of the WHOLE scalp) as the gather target:
 * cudaGraphicsGLRegisterBuffer on render_device; bh_group_gather_to_gl maps it, gathers the planes in plane_mask
 * (0 = positions only) to their PingPongBuffer offsets and unmaps. Needs a current GL context on the calling thread. */
int     bh_group_register_gl_buffer(bh_group* group, unsigned int gl_buffer, int render_device);
int     bh_group_unregister_gl_buffer(bh_group* group);
int     bh_group_gather_to_gl(bh_group* group, unsigned plane_mask);

/* ---- CUDA-GL interop on buffer 0 (pbuffer_.read_ssbo_id(), hair.cc:371) --------------------- */
/* cudaGraphicsGLRegisterBuffer; while registered, bh_step maps the GL buffer, steps in place and
 * unmaps, so the render VAO (hair.cc:371-389) sees the new positions without a copy. Needs a
 * current GL context on the calling thread; returns BH_ERR_CUDA otherwise. */
int  bh_register_gl_buffer(bh_sim* sim, unsigned int gl_buffer);
int  bh_unregister_gl_buffer(bh_sim* sim);
/* The same protocol for a device allocation the caller owns (another CUDA allocation, a Vulkan / D3D buffer imported through
 * cudaImportExternalMemory, a slice of the renderer's arena): the state moves into it, every entry point brackets its work
 * with the map / unmap pair exactly as for a GL buffer (with a plain allocation the pair only keeps the books), and
 * bh_unregister_device_buffer moves the state back. `device_ptr`: >= 3 * V * 16 bytes on the sim's GPU, 128-byte aligned.
 * This is also how the GL path's state machine is exercised where no GL context exists (tests/test_shared_buffer.py). */
int  bh_register_device_buffer(bh_sim* sim, void* device_ptr, uint64_t nbytes);
int  bh_unregister_device_buffer(bh_sim* sim);
/* Map / unmap bookkeeping of the registered buffer (GL or caller-owned): calls so far and whether it is mapped right now —
 * after any entry point has returned, successfully or not, maps == unmaps and mapped_now == 0. */
int  bh_buffer_map_stats(const bh_sim* sim, int64_t* maps, int64_t* unmaps, int* mapped_now);

const char* bh_last_error(void);     /* thread-local message of the last failing call */
const char* bh_version(void);

#ifdef __cplusplus
}
#endif
#endif /* BARBU_HAIR_H_ */
