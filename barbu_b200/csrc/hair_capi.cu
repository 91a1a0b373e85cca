// hair_capi.cu — implementation of include/barbu_hair.h on top of the kernels.
//
// Host-side counterpart of Hair::{setup, update, set_bounding_sphere} (src/fx/hair.cc:42-125) and of
// PingPongBuffer (src/memory/pingpong_buffer.cc): one device allocation holding the three SoA float4
// planes of "buffer 0", updated in place. No CPU fallback: every path below ends in a CUDA call.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "hair_gen.cuh"
#include "hair_step.cuh"

// cuda_gl_interop.h needs <GL/gl.h>, which this image does not have; the two entry points we use are
// in libcudart and take a plain GLuint (unsigned int).
extern "C" cudaError_t cudaGraphicsGLRegisterBuffer(struct cudaGraphicsResource** resource, unsigned int buffer,
                                                    unsigned int flags);

#include "hair_sim.cuh"

namespace {
thread_local std::string g_last_error;
}  // namespace

namespace bh {
int fail(int code, const char* what, cudaError_t e) {
  char buf[512];
  if (e != cudaSuccess) std::snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  else std::snprintf(buf, sizeof buf, "%s", what);
  g_last_error = buf;
  return code;
}
}  // namespace bh

using bh::fail;
using bh::DeviceGuard;
using bh::kHostPipeStreams;

namespace {

constexpr int64_t kAutoLatencyVertices = 1 << 18;

bh::StepArgs make_args(const bh_sim* s, float dt, float4* pos, float4* vel, int64_t nstrands) {
  const bh_params& p = s->params;
  bh::StepArgs a;
  std::memset(&a, 0, sizeof a);
  a.pos = pos; a.vel = vel; a.nstrands = nstrands; a.nverts = s->nverts; a.iterations = p.iterations;
  a.dt = dt;
  a.dt2 = dt * dt;                                   // vec3 dt*dt, cs_simulation.glsl:182
  // force = kForceCoeff * gravity (cs:74), one fp32 product per component; wind extension added after
  volatile float fx = p.force_coeff * p.gravity[0], fy = p.force_coeff * p.gravity[1], fz = p.force_coeff * p.gravity[2];
  if (p.wind[0] != 0.0f || p.wind[1] != 0.0f || p.wind[2] != 0.0f) { fx = fx + p.wind[0]; fy = fy + p.wind[1]; fz = fz + p.wind[2]; }
  a.fx = fx; a.fy = fy; a.fz = fz;
  a.sf = p.scale; a.damp = p.damp;
  a.cx = p.sphere[0]; a.cy = p.sphere[1]; a.cz = p.sphere[2]; a.r = p.sphere[3];
  a.r2 = p.sphere[3] * p.sphere[3];                  // radius * radius, cs:133
  a.use_drag = (p.drag != 0.0f) ? 1 : 0;
  a.keep = 1.0f - p.drag;
  // small scalps: the latency-oriented kernel (hair_wave.cu). AUTO draws the line where the throughput kernels stop being
  // latency-bound themselves (tools/small_latency.py: up to 2^18 vertices the wavefront kernel is the faster one)
  a.prefer_latency = s->step_policy == BH_POLICY_LATENCY || (s->step_policy == BH_POLICY_AUTO && s->nvertices <= kAutoLatencyVertices);   // by the sim's size, not a slice's
  a.ncaps = p.ncapsules;
  for (int q = 0; q < p.ncapsules && q < bh::kMaxCapsules; ++q) {
    const bh_capsule& c = p.capsules[q];
    a.caps[q] = bh::Capsule{ c.a[0], c.a[1], c.a[2], c.b[0], c.b[1], c.b[2], c.radius };
  }
  return a;
}

int ensure_roots(bh_sim* s) {
  if (!s->root_pos3) BH_CUDA(cudaMalloc(&s->root_pos3, sizeof(float) * 3 * (size_t)s->nstrands));
  if (!s->root_nrm3) BH_CUDA(cudaMalloc(&s->root_nrm3, sizeof(float) * 3 * (size_t)s->nstrands));
  return BH_OK;
}

}  // namespace

namespace bh {
int unmap_gl(bh_sim* s);
bool shared_buffer(const bh_sim* s) { return s->gl_resource != nullptr || s->ext_buffer != nullptr; }

// Buffer 0 lives in a buffer the renderer owns (a GL buffer object, hair.cc:371-389) or, same protocol, in a device
// allocation of the caller: it is only OURS between map and unmap. Cleans up after itself: a failure leaves it unmapped.
int map_gl(bh_sim* s) {
  if (!shared_buffer(s)) return BH_OK;
  if (s->shared_mapped) return fail(BH_ERR_INVALID, "shared buffer mapped twice");
  void* ptr = s->ext_buffer; size_t bytes = s->ext_bytes;
  if (s->gl_resource) {
    BH_CUDA(cudaGraphicsMapResources(1, &s->gl_resource, s->stream));
    const cudaError_t e = cudaGraphicsResourceGetMappedPointer(&ptr, &bytes, s->gl_resource);
    if (e != cudaSuccess) { (void)cudaGraphicsUnmapResources(1, &s->gl_resource, s->stream); (void)cudaGetLastError(); return fail(BH_ERR_CUDA, "cudaGraphicsResourceGetMappedPointer", e); }
  }
  s->shared_maps += 1; s->shared_mapped = true;
  if (bytes < (size_t)BH_NUM_PLANES * s->nvertices * sizeof(float4)) { (void)unmap_gl(s); return fail(BH_ERR_INVALID, "shared buffer smaller than 3 planes"); }
  for (int p = 0; p < BH_NUM_PLANES; ++p) s->planes[p] = static_cast<float4*>(ptr) + (size_t)p * s->nvertices;
  return BH_OK;
}
int unmap_gl(bh_sim* s) {
  if (!shared_buffer(s) || !s->shared_mapped) return BH_OK;
  s->shared_unmaps += 1; s->shared_mapped = false;
  if (s->gl_resource) BH_CUDA(cudaGraphicsUnmapResources(1, &s->gl_resource, s->stream));
  return BH_OK;
}

}  // namespace bh
using bh::map_gl;
using bh::unmap_gl;
using bh::shared_buffer;

namespace {
// The sim's work moves to another stream: everything already queued on the old one comes first. All steps of a sim share one
// tile-scheduler slot and update the state in place, so two steps may never overlap — without this, back-to-back steps
// issued on two different streams could.
int switch_stream(bh_sim* s, cudaStream_t next) {
  if (next == s->stream) return BH_OK;
  DeviceGuard g(s->device);
  if (!s->order_event) BH_CUDA(cudaEventCreateWithFlags(&s->order_event, cudaEventDisableTiming));
  BH_CUDA(cudaEventRecord(s->order_event, s->stream));
  BH_CUDA(cudaStreamWaitEvent(next, s->order_event, 0));
  s->stream = next;
  return BH_OK;
}

// launch_step, and after a failure the scheduler words back to their idle state (a launch that died half-way leaves its
// tile counter wherever it was, and every later launch on this slot would skip those tiles).
cudaError_t launch_step_checked(bh_sim* s, const bh::StepArgs& a, cudaStream_t stream, unsigned int* words) {
  const cudaError_t e = bh::launch_step(a, s->params.math, stream, words);
  if (e != cudaSuccess) { (void)cudaGetLastError(); (void)cudaMemsetAsync(words, 0, 2 * sizeof(unsigned int), stream); (void)cudaGetLastError(); }
  return e;
}

// Maps the registered GL buffer for the lifetime of the object: every way out of an entry point, the early returns of
// BH_CUDA included, leaves the buffer unmapped (the renderer cannot touch a mapped buffer). done() unmaps explicitly and
// reports the unmap's own status on the regular path.
struct GlMapped {
  bh_sim* s; int rc; bool mapped;
  explicit GlMapped(bh_sim* sim) : s(sim), rc(map_gl(sim)), mapped(rc == BH_OK) {}
  int done() { if (!mapped) return rc; mapped = false; return unmap_gl(s); }
  ~GlMapped() { if (mapped) { (void)unmap_gl(s); } }
  GlMapped(const GlMapped&) = delete; GlMapped& operator=(const GlMapped&) = delete;
};
}  // namespace

extern "C" {

const char* bh_last_error(void) { return g_last_error.c_str(); }
int bh_host_fail_impl(int code, const char* what) { return fail(code, what); }      // for hair_host.cc (no CUDA headers there)
const char* bh_version(void) { return "barbu_hair 0.1 (sm_100a)"; }

void bh_default_params(bh_params* p) {
  if (!p) return;
  std::memset(p, 0, sizeof *p);
  p->scale = 1.45f;                                   // hair.h:41 via hair.cc:108
  p->sphere[3] = 1.0f;                                // cs:43
  p->iterations = 8;                                  // cs:197
  p->gravity[1] = -9.81f;                             // cs:71
  p->force_coeff = 20.0f;                             // cs:72
  p->damp = 0.80f;                                    // cs:102 as lane 0 evaluates it
  p->math = BH_MATH_EXACT;
}

int bh_create(bh_sim** out, int64_t nstrands, int nverts, int device) {
  if (!out) return fail(BH_ERR_INVALID, "bh_create: out is NULL");
  *out = nullptr;
  if (nstrands <= 0 || nverts <= 0) return fail(BH_ERR_INVALID, "bh_create: nstrands and nverts must be positive");
  int ndev = 0;
  BH_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(BH_ERR_INVALID, "bh_create: no such CUDA device");
  DeviceGuard g(device);
  if (!g.ok) return fail(BH_ERR_CUDA, "bh_create: cudaSetDevice failed");
  bh_sim* s = new (std::nothrow) bh_sim();
  if (!s) return fail(BH_ERR_INVALID, "bh_create: out of host memory");
  s->device = device; s->nstrands = nstrands; s->nverts = nverts; s->nvertices = nstrands * (int64_t)nverts;
  bh_default_params(&s->params);
  static const bool fuse_env = [] { const char* e = getenv("BH_SUBSTEP_FUSION"); return e && e[0] == '1'; }();   // tuning knob: default of bh_set_substep_fusion
  s->fuse_substeps = fuse_env ? 1 : 0;
  cudaError_t e = cudaMalloc(&s->buffer0, (size_t)BH_NUM_PLANES * s->nvertices * sizeof(float4));
  // defined contents from the start: a plane nobody uploaded (the tangents of a sim fed positions only) still travels in
  // whole-buffer copies — state files, the move into a shared buffer — and reads as zeros there, not as stale memory
  if (e == cudaSuccess) e = cudaMemset(s->buffer0, 0, (size_t)BH_NUM_PLANES * s->nvertices * sizeof(float4));
  if (e == cudaSuccess) e = cudaMalloc(&s->tile_counters, sizeof(unsigned int) * 32 * (kHostPipeStreams + 1));
  if (e == cudaSuccess) {
    unsigned int words[32 * (kHostPipeStreams + 1)] = { 0 };
    for (int i = 0; i <= kHostPipeStreams; ++i) bh::init_sched_words(words + 32 * i);
    e = cudaMemcpy(s->tile_counters, words, sizeof words, cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking);
  for (int i = 0; i < kHostPipeStreams && e == cudaSuccess; ++i) e = cudaStreamCreateWithFlags(&s->pipe[i], cudaStreamNonBlocking);
  if (e != cudaSuccess) { (void)cudaGetLastError(); bh_destroy(s); return fail(BH_ERR_CUDA, "bh_create: allocation", e); }
  s->stream = s->own_stream;
  for (int p = 0; p < BH_NUM_PLANES; ++p) s->planes[p] = s->buffer0 + (size_t)p * s->nvertices;
  *out = s;
  return BH_OK;
}

int bh_destroy(bh_sim* s) {
  if (!s) return BH_OK;
  DeviceGuard g(s->device);
  if (s->gl_resource) cudaGraphicsUnregisterResource(s->gl_resource);
  cudaFree(s->buffer0); cudaFree(s->tile_counters); cudaFree(s->root_pos3); cudaFree(s->root_nrm3);
  cudaFree(s->tess_patch); cudaFree(s->tess_out); cudaFree(s->checksum_words);
  cudaFree(s->skin_rest3); cudaFree(s->skin_joints4); cudaFree(s->skin_weights3); cudaFree(s->skin_dq);
  if (s->own_stream) cudaStreamDestroy(s->own_stream);
  for (auto& st : s->pipe) if (st) cudaStreamDestroy(st);
  for (cudaEvent_t ev : s->host_events) cudaEventDestroy(ev);
  if (s->order_event) cudaEventDestroy(s->order_event);
  delete s;
  return BH_OK;
}

int bh_set_stream(bh_sim* s, void* cuda_stream) {
  if (!s) return fail(BH_ERR_INVALID, "bh_set_stream: sim is NULL");
  return switch_stream(s, static_cast<cudaStream_t>(cuda_stream));
}

int bh_reset_stream(bh_sim* s) {
  if (!s) return fail(BH_ERR_INVALID, "bh_reset_stream: sim is NULL");
  return switch_stream(s, s->own_stream);
}

int bh_synchronize(bh_sim* s) {
  if (!s) return fail(BH_ERR_INVALID, "bh_synchronize: sim is NULL");
  DeviceGuard g(s->device);
  BH_CUDA(cudaStreamSynchronize(s->stream));
  return BH_OK;
}

int bh_set_params(bh_sim* s, const bh_params* p) {
  if (!s || !p) return fail(BH_ERR_INVALID, "bh_set_params: NULL argument");
  if (p->iterations < 0) return fail(BH_ERR_INVALID, "bh_set_params: iterations < 0");
  if (p->ncapsules < 0 || p->ncapsules > BH_MAX_CAPSULES) return fail(BH_ERR_INVALID, "bh_set_params: ncapsules out of range");
  if (p->math != BH_MATH_EXACT && p->math != BH_MATH_FAST) return fail(BH_ERR_INVALID, "bh_set_params: unknown math profile");
  s->params = *p;
  return BH_OK;
}

int bh_get_params(const bh_sim* s, bh_params* p) {
  if (!s || !p) return fail(BH_ERR_INVALID, "bh_get_params: NULL argument");
  *p = s->params;
  return BH_OK;
}

int bh_set_bounding_sphere(bh_sim* s, const float sphere[4]) {
  if (!s || !sphere) return fail(BH_ERR_INVALID, "bh_set_bounding_sphere: NULL argument");
  std::memcpy(s->params.sphere, sphere, 4 * sizeof(float));
  return BH_OK;
}

int bh_upload(bh_sim* s, const float* pos4, const float* vel4, const float* tan4) {
  if (!s) return fail(BH_ERR_INVALID, "bh_upload: sim is NULL");
  DeviceGuard g(s->device);
  const size_t bytes = (size_t)s->nvertices * sizeof(float4);
  const float* src[BH_NUM_PLANES] = { pos4, vel4, tan4 };
  GlMapped gl(s); if (gl.rc) return gl.rc;
  for (int p = 0; p < BH_NUM_PLANES; ++p)
    if (src[p]) BH_CUDA(cudaMemcpyAsync(s->planes[p], src[p], bytes, cudaMemcpyHostToDevice, s->stream));
  if (int rc = gl.done()) return rc;
  BH_CUDA(cudaStreamSynchronize(s->stream));
  if (pos4) s->initialized = true;
  return BH_OK;
}

int bh_download(bh_sim* s, float* pos4, float* vel4, float* tan4) {
  if (!s) return fail(BH_ERR_INVALID, "bh_download: sim is NULL");
  DeviceGuard g(s->device);
  const size_t bytes = (size_t)s->nvertices * sizeof(float4);
  float* dst[BH_NUM_PLANES] = { pos4, vel4, tan4 };
  GlMapped gl(s); if (gl.rc) return gl.rc;
  for (int p = 0; p < BH_NUM_PLANES; ++p)
    if (dst[p]) BH_CUDA(cudaMemcpyAsync(dst[p], s->planes[p], bytes, cudaMemcpyDeviceToHost, s->stream));
  if (int rc = gl.done()) return rc;
  BH_CUDA(cudaStreamSynchronize(s->stream));
  return BH_OK;
}

int bh_device_plane(bh_sim* s, int plane, void** device_ptr, uint64_t* nbytes) {
  if (!s || plane < 0 || plane >= BH_NUM_PLANES || !device_ptr) return fail(BH_ERR_INVALID, "bh_device_plane: bad argument");
  if (shared_buffer(s)) return fail(BH_ERR_UNSUPPORTED, "bh_device_plane: buffer 0 is a GL / caller-owned buffer (only ours during a call)");
  *device_ptr = s->planes[plane];
  if (nbytes) *nbytes = (uint64_t)s->nvertices * sizeof(float4);
  return BH_OK;
}

int bh_init_strands(bh_sim* s, const float* root_pos3, const float* root_nrm3, const float* random_value, float maxlength) {
  if (!s || !root_pos3 || !root_nrm3 || !random_value) return fail(BH_ERR_INVALID, "bh_init_strands: NULL argument");
  DeviceGuard g(s->device);
  int rc = ensure_roots(s); if (rc) return rc;
  float* d_rv = nullptr;
  const size_t S = (size_t)s->nstrands;
  BH_CUDA(cudaMalloc(&d_rv, sizeof(float) * S));
  cudaError_t e = cudaMemcpyAsync(s->root_pos3, root_pos3, sizeof(float) * 3 * S, cudaMemcpyHostToDevice, s->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(s->root_nrm3, root_nrm3, sizeof(float) * 3 * S, cudaMemcpyHostToDevice, s->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_rv, random_value, sizeof(float) * S, cudaMemcpyHostToDevice, s->stream);
  rc = BH_OK;
  if (e == cudaSuccess) rc = map_gl(s);
  if (e == cudaSuccess && rc == BH_OK) {
    const float scaleOffset = maxlength / static_cast<float>(s->nverts);       // hair.cc:265
    e = bh::launch_expand_strands(s->root_pos3, s->root_nrm3, d_rv, s->nstrands, s->nverts, scaleOffset,
                                  s->planes[BH_PLANE_POSITION], s->planes[BH_PLANE_VELOCITY], s->stream);
    s->launches += 1;
    if (e == cudaSuccess) rc = unmap_gl(s);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
  cudaFree(d_rv);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(BH_ERR_CUDA, "bh_init_strands", e); }
  if (rc) return rc;
  s->initialized = true;
  return BH_OK;
}

int bh_init_sphere_scalp(bh_sim* s, int rows, int cols, int64_t first, const float* random_value, float maxlength) {
  return bh_init_sphere_scalp_ordered(s, rows, cols, BH_SCALP_ROW_MAJOR, first, random_value, maxlength);
}

int bh_init_sphere_scalp_ordered(bh_sim* s, int rows, int cols, int order, int64_t first, const float* random_value, float maxlength) {
  if (!s || !random_value || rows <= 0 || cols <= 0) return fail(BH_ERR_INVALID, "bh_init_sphere_scalp: bad argument");
  if (order != BH_SCALP_ROW_MAJOR && order != BH_SCALP_COLUMN_MAJOR) return fail(BH_ERR_INVALID, "bh_init_sphere_scalp: unknown strand order");
  if (first < 0 || first + s->nstrands > (int64_t)rows * cols) return fail(BH_ERR_INVALID, "bh_init_sphere_scalp: strand range outside the rows*cols grid");
  DeviceGuard g(s->device);
  int rc = ensure_roots(s); if (rc) return rc;
  // Row / column trigonometry in double by libm, rounded to fp32 once; the device only multiplies.
  const double kPi = 3.14159265358979323846;
  std::vector<float> rowtab(2 * (size_t)rows), coltab(2 * (size_t)cols);
  for (int r = 0; r < rows; ++r) {
    const double th = kPi * ((double)r + 0.5) / (double)rows - kPi / 2.0;
    rowtab[2 * r] = (float)std::cos(th); rowtab[2 * r + 1] = (float)std::sin(th);
  }
  for (int c = 0; c < cols; ++c) {
    const double ph = 2.0 * kPi * (double)c / (double)cols;
    coltab[2 * c] = (float)std::cos(ph); coltab[2 * c + 1] = (float)std::sin(ph);
  }
  float *d_row = nullptr, *d_col = nullptr, *d_rv = nullptr;
  const size_t S = (size_t)s->nstrands;
  cudaError_t e = cudaMalloc(&d_row, sizeof(float) * rowtab.size());
  if (e == cudaSuccess) e = cudaMalloc(&d_col, sizeof(float) * coltab.size());
  if (e == cudaSuccess) e = cudaMalloc(&d_rv, sizeof(float) * S);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_row, rowtab.data(), sizeof(float) * rowtab.size(), cudaMemcpyHostToDevice, s->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_col, coltab.data(), sizeof(float) * coltab.size(), cudaMemcpyHostToDevice, s->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_rv, random_value, sizeof(float) * S, cudaMemcpyHostToDevice, s->stream);
  rc = BH_OK;
  if (e == cudaSuccess) rc = map_gl(s);
  if (e == cudaSuccess && rc == BH_OK) {
    e = bh::launch_sphere_roots(d_row, d_col, rows, cols, order == BH_SCALP_COLUMN_MAJOR, first, s->nstrands, s->root_pos3, s->root_nrm3, s->stream);
    const float scaleOffset = maxlength / static_cast<float>(s->nverts);
    if (e == cudaSuccess)
      e = bh::launch_expand_strands(s->root_pos3, s->root_nrm3, d_rv, s->nstrands, s->nverts, scaleOffset,
                                    s->planes[BH_PLANE_POSITION], s->planes[BH_PLANE_VELOCITY], s->stream);
    s->launches += 2;
    if (e == cudaSuccess) rc = unmap_gl(s);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
  cudaFree(d_row); cudaFree(d_col); cudaFree(d_rv);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(BH_ERR_CUDA, "bh_init_sphere_scalp", e); }
  if (rc) return rc;
  s->initialized = true;
  return BH_OK;
}

int bh_build_patch_indices(const int32_t* tri, int64_t nfaces, int nverts, int32_t* out, int device) {
  if (!tri || !out || nfaces < 0 || nverts < 1) return fail(BH_ERR_INVALID, "bh_build_patch_indices: bad argument");
  if (nfaces == 0 || nverts == 1) return BH_OK;
  // the reference keeps element values and counts in `int` (hair.cc:398-404): refuse what would wrap
  int32_t maxidx = 0;
  for (int64_t q = 0; q < 3 * nfaces; ++q) { if (tri[q] < 0) return fail(BH_ERR_INVALID, "negative scalp index"); if (tri[q] > maxidx) maxidx = tri[q]; }
  if ((int64_t)nverts * maxidx + (nverts - 1) > INT32_MAX) return fail(BH_ERR_OVERFLOW, "bh_build_patch_indices: element index exceeds int32");
  int ndev = 0;
  BH_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(BH_ERR_INVALID, "bh_build_patch_indices: no such CUDA device");
  DeviceGuard g(device);
  const size_t nin = 3 * (size_t)nfaces, nout = 6 * (size_t)nfaces * (size_t)(nverts - 1);
  int *d_in = nullptr, *d_out = nullptr;
  cudaError_t e = cudaMalloc(&d_in, nin * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&d_out, nout * sizeof(int));
  if (e == cudaSuccess) e = cudaMemcpy(d_in, tri, nin * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = bh::launch_patch_indices(d_in, nfaces, nverts, d_out, nullptr);
  if (e == cudaSuccess) e = cudaMemcpy(out, d_out, nout * sizeof(int), cudaMemcpyDeviceToHost);
  cudaFree(d_in); cudaFree(d_out);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(BH_ERR_CUDA, "bh_build_patch_indices", e); }
  return BH_OK;
}

int bh_step(bh_sim* s, float dt, int substeps) {
  if (!s) return fail(BH_ERR_INVALID, "bh_step: sim is NULL");
  if (!s->initialized) return fail(BH_ERR_NOT_INITIALIZED, "bh_step: no strand state (call bh_upload / bh_init_* first)");
  if (substeps < 1) return fail(BH_ERR_INVALID, "bh_step: substeps < 1");
  DeviceGuard g(s->device);
  GlMapped gl(s); if (gl.rc) return gl.rc;
  const float h = (substeps == 1) ? dt : dt / static_cast<float>(substeps);
  bh::StepArgs a = make_args(s, h, s->planes[BH_PLANE_POSITION], s->planes[BH_PLANE_VELOCITY], s->nstrands);
  static const bool zigzag = [] { const char* e = getenv("BH_NO_ZIGZAG"); return !(e && e[0] == '1'); }();
  if (a.prefer_latency && bh::wave_kernel_eligible(a)) {
    // latency-oriented kernel: a warp owns its strands outright, so the substeps of the frame are simply passes of one launch
    a.passes = substeps;
    BH_CUDA(launch_step_checked(s, a, s->stream, s->tile_counters + 32 * kHostPipeStreams));
    s->launches += 1;
    s->step_launches += 1;
    return gl.done();
  }
  if (s->fuse_substeps && bh::stream_fusion_eligible(a, substeps, s->fuse_substeps == 2)) {
    // Frame-level fusion: the substeps of this frame as the passes of ONE launch (StepArgs::passes). Same arithmetic, same
    // order per strand, so the result is bit-identical to `substeps` launches; HBM sees the state once per frame.
    a.passes = substeps;
    a.reverse = zigzag ? (int)(s->step_launches & 1) : 0;
    BH_CUDA(launch_step_checked(s, a, s->stream, s->tile_counters + 32 * kHostPipeStreams));
    s->launches += 1;
    s->step_launches += 1;
    return gl.done();
  }
  for (int q = 0; q < substeps; ++q) {
    // Consecutive launches walk the shard in opposite directions: a launch starts with the tiles the previous one wrote
    // last, which are still in the 126 MB L2 — those reads, and the write-backs they replace, never reach HBM.
    a.reverse = zigzag ? (int)(s->step_launches & 1) : 0;
    BH_CUDA(launch_step_checked(s, a, s->stream, s->tile_counters + 32 * kHostPipeStreams));
    s->launches += 1;
    s->step_launches += 1;
  }
  return gl.done();
}

int bh_set_step_policy(bh_sim* s, int policy) {
  if (!s) return fail(BH_ERR_INVALID, "bh_set_step_policy: sim is NULL");
  if (policy != BH_POLICY_THROUGHPUT && policy != BH_POLICY_LATENCY && policy != BH_POLICY_AUTO) return fail(BH_ERR_INVALID, "bh_set_step_policy: unknown policy");
  s->step_policy = policy;
  return BH_OK;
}

int bh_set_substep_fusion(bh_sim* s, int enabled) {
  if (!s) return fail(BH_ERR_INVALID, "bh_set_substep_fusion: sim is NULL");
  if (enabled < 0 || enabled > 2) return fail(BH_ERR_INVALID, "bh_set_substep_fusion: 0 (off), 1 (where it pays) or 2 (always)");
  s->fuse_substeps = enabled;
  return BH_OK;
}

int bh_step_host(bh_sim* s, float dt, int substeps, float* pos4, float* vel4) {
  if (!s || !pos4 || !vel4) return fail(BH_ERR_INVALID, "bh_step_host: NULL argument");
  if (substeps < 1) return fail(BH_ERR_INVALID, "bh_step_host: substeps < 1");
  if (shared_buffer(s)) return fail(BH_ERR_UNSUPPORTED, "bh_step_host: buffer 0 is a GL / caller-owned buffer");
  DeviceGuard g(s->device);
  const float h = (substeps == 1) ? dt : dt / static_cast<float>(substeps);
  // Strands are independent, so a slice of strands can be uploaded, stepped `substeps` times and downloaded while its
  // neighbours are still in flight. Three streams, one per engine — pipe[0] uploads, pipe[1] runs the kernels, pipe[2]
  // downloads — chained per slice by events, so no download ever waits behind an upload of a later slice (PCIe is full
  // duplex) and the only serial parts of a call are the first upload and the last download.
  // 16 slices measured best at configs[1] (profiles/r01_e2e_slices.txt): fewer lengthen the serial first upload / last
  // download, more add per-slice costs.
  const int64_t S = s->nstrands;
  static const int nslices = [] { const char* e = getenv("BH_HOST_SLICES"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 16; }();   // tuning knob
  std::vector<int64_t> bounds;                                              // slice k = [bounds[k], bounds[k + 1])
  {
    int64_t slice = (S + nslices - 1) / nslices;
    if (slice < 4096) slice = 4096;
    slice = (slice + 127) / 128 * 128;                                      // whole 32-strand tiles, even strand counts
    for (int64_t x = 0; x < S; x += slice) bounds.push_back(x);
    bounds.push_back(S);
  }
  const int64_t count_slices = (int64_t)bounds.size() - 1;
  while ((int64_t)s->host_events.size() < 2 * count_slices + 1) {
    cudaEvent_t ev;
    BH_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    s->host_events.push_back(ev);
  }
  cudaStream_t up = s->pipe[0], run = s->pipe[1], down = s->pipe[2];
  // order after whatever is queued on the sim's stream
  cudaEvent_t ev0 = s->host_events[2 * count_slices];
  cudaError_t e = cudaEventRecord(ev0, s->stream);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(up, ev0, 0);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(run, ev0, 0);
  for (int64_t k = 0; k < count_slices && e == cudaSuccess; ++k) {
    const int64_t first = bounds[k], count = bounds[k + 1] - bounds[k];
    const size_t off = (size_t)first * s->nverts, bytes = (size_t)count * s->nverts * sizeof(float4);
    float4* dP = s->planes[BH_PLANE_POSITION] + off;
    float4* dV = s->planes[BH_PLANE_VELOCITY] + off;
    cudaEvent_t uploaded = s->host_events[2 * k], stepped = s->host_events[2 * k + 1];
    e = cudaMemcpyAsync(dP, pos4 + 4 * off, bytes, cudaMemcpyHostToDevice, up);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dV, vel4 + 4 * off, bytes, cudaMemcpyHostToDevice, up);
    if (e == cudaSuccess) e = cudaEventRecord(uploaded, up);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(run, uploaded, 0);
    bh::StepArgs a = make_args(s, h, dP, dV, count);
    if (e == cudaSuccess && s->fuse_substeps && bh::stream_fusion_eligible(a, substeps, s->fuse_substeps == 2)) {
      a.passes = substeps;                                                  // the substeps of the slice as passes of one launch
      e = launch_step_checked(s, a, run, s->tile_counters); s->launches += 1;
    } else {
      for (int q = 0; q < substeps && e == cudaSuccess; ++q) { e = launch_step_checked(s, a, run, s->tile_counters); s->launches += 1; }
    }
    if (e == cudaSuccess) e = cudaEventRecord(stepped, run);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(down, stepped, 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(pos4 + 4 * off, dP, bytes, cudaMemcpyDeviceToHost, down);
    if (e == cudaSuccess) e = cudaMemcpyAsync(vel4 + 4 * off, dV, bytes, cudaMemcpyDeviceToHost, down);
  }
  for (int i = 0; i < 3; ++i) { cudaError_t e2 = cudaStreamSynchronize(s->pipe[i]); if (e == cudaSuccess) e = e2; }
  if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(BH_ERR_CUDA, "bh_step_host", e); }
  s->initialized = true;
  return BH_OK;
}

int bh_step_readback(bh_sim* s, float dt, int substeps, float* pos4) {
  if (!s || !pos4) return fail(BH_ERR_INVALID, "bh_step_readback: NULL argument");
  if (!s->initialized) return fail(BH_ERR_NOT_INITIALIZED, "bh_step_readback: no strand state (call bh_upload / bh_init_* first)");
  if (substeps < 1) return fail(BH_ERR_INVALID, "bh_step_readback: substeps < 1");
  if (shared_buffer(s)) return fail(BH_ERR_UNSUPPORTED, "bh_step_readback: buffer 0 is a GL / caller-owned buffer (the renderer reads it in place)");
  DeviceGuard g(s->device);
  const float h = (substeps == 1) ? dt : dt / static_cast<float>(substeps);
  // The reference's frame: uniforms in, Hair::update, positions out (hair.cc:89-125 + the consumer of buffer 0). Strands are
  // independent, so the shard is stepped slice by slice on one stream while the position plane of the slices already done
  // leaves on another: the device->host copy (the long pole: 16 B per vertex over PCIe) starts after 1/nslices of the
  // compute instead of after all of it, and the step of slice k+1 hides behind the copy of slice k.
  const int64_t S = s->nstrands;
  static const int nslices = [] { const char* e = getenv("BH_READBACK_SLICES"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 8; }();   // tuning knob
  int64_t slice = (S + nslices - 1) / nslices;
  if (slice < 8192) slice = 8192;
  slice = (slice + 127) / 128 * 128;                                        // whole 32-strand tiles, even strand counts
  const int64_t count_slices = (S + slice - 1) / slice;
  while ((int64_t)s->host_events.size() < 2 * count_slices + 1) {
    cudaEvent_t ev;
    BH_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    s->host_events.push_back(ev);
  }
  cudaStream_t run = s->pipe[1], down = s->pipe[2];
  cudaEvent_t ev0 = s->host_events[2 * count_slices];
  cudaError_t e = cudaEventRecord(ev0, s->stream);                          // order after whatever is queued on the sim's stream
  if (e == cudaSuccess) e = cudaStreamWaitEvent(run, ev0, 0);
  for (int64_t k = 0; k < count_slices && e == cudaSuccess; ++k) {
    const int64_t first = k * slice, count = (first + slice <= S ? slice : S - first);
    const size_t off = (size_t)first * s->nverts, bytes = (size_t)count * s->nverts * sizeof(float4);
    float4* dP = s->planes[BH_PLANE_POSITION] + off;
    bh::StepArgs a = make_args(s, h, dP, s->planes[BH_PLANE_VELOCITY] + off, count);
    if (s->fuse_substeps && bh::stream_fusion_eligible(a, substeps, s->fuse_substeps == 2)) {
      a.passes = substeps;
      e = launch_step_checked(s, a, run, s->tile_counters); s->launches += 1;
    } else {
      for (int q = 0; q < substeps && e == cudaSuccess; ++q) { e = launch_step_checked(s, a, run, s->tile_counters); s->launches += 1; }
    }
    cudaEvent_t stepped = s->host_events[2 * k];
    if (e == cudaSuccess) e = cudaEventRecord(stepped, run);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(down, stepped, 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(pos4 + 4 * off, dP, bytes, cudaMemcpyDeviceToHost, down);
  }
  for (int i = 1; i < 3; ++i) { cudaError_t e2 = cudaStreamSynchronize(s->pipe[i]); if (e == cudaSuccess) e = e2; }
  if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(BH_ERR_CUDA, "bh_step_readback", e); }
  return BH_OK;
}

int bh_host_alloc(void** ptr, uint64_t nbytes) {
  if (!ptr) return fail(BH_ERR_INVALID, "bh_host_alloc: ptr is NULL");
  // portable: page-locked for every CUDA context of the process (the shards of a bh_group live on several devices)
  BH_CUDA(cudaHostAlloc(ptr, (size_t)nbytes, cudaHostAllocPortable));
  return BH_OK;
}
int bh_host_free(void* ptr) {
  if (ptr) BH_CUDA(cudaFreeHost(ptr));
  return BH_OK;
}

int64_t bh_launch_count(const bh_sim* s) { return s ? s->launches : 0; }

int bh_step_kernel_kind(const bh_sim* s) {
  if (!s) return -1;
  return bh::step_kernel_kind(make_args(s, 0.0f, s->planes[BH_PLANE_POSITION], s->planes[BH_PLANE_VELOCITY], s->nstrands));
}

int bh_selftest_math(int device, uint64_t* mismatches) {
  if (!mismatches) return fail(BH_ERR_INVALID, "bh_selftest_math: NULL argument");
  int ndev = 0;
  BH_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(BH_ERR_INVALID, "bh_selftest_math: no such CUDA device");
  DeviceGuard g(device);
  unsigned long long bad = 0;
  BH_CUDA(bh::selftest_inversesqrt(&bad));
  *mismatches = bad;
  return BH_OK;
}

int bh_set_skin(bh_sim* s, const float* rest_root_pos3, const int32_t* joints4, const float* weights3) {
  if (!s || !rest_root_pos3 || !joints4 || !weights3) return fail(BH_ERR_INVALID, "bh_set_skin: NULL argument");
  const size_t S = (size_t)s->nstrands;
  // the kernel reads dq + 8 * joint: remember the range so that bh_skin_roots can refuse a palette that is too short
  int32_t jmin = 0, jmax = 0;
  for (size_t q = 0; q < 4 * S; ++q) { if (joints4[q] < jmin) jmin = joints4[q]; if (joints4[q] > jmax) jmax = joints4[q]; }
  if (jmin < 0) return fail(BH_ERR_INVALID, "bh_set_skin: negative joint index");
  DeviceGuard g(s->device);
  if (!s->skin_rest3) BH_CUDA(cudaMalloc(&s->skin_rest3, sizeof(float) * 3 * S));
  if (!s->skin_joints4) BH_CUDA(cudaMalloc(&s->skin_joints4, sizeof(int) * 4 * S));
  if (!s->skin_weights3) BH_CUDA(cudaMalloc(&s->skin_weights3, sizeof(float) * 3 * S));
  BH_CUDA(cudaMemcpyAsync(s->skin_rest3, rest_root_pos3, sizeof(float) * 3 * S, cudaMemcpyHostToDevice, s->stream));
  BH_CUDA(cudaMemcpyAsync(s->skin_joints4, joints4, sizeof(int) * 4 * S, cudaMemcpyHostToDevice, s->stream));
  BH_CUDA(cudaMemcpyAsync(s->skin_weights3, weights3, sizeof(float) * 3 * S, cudaMemcpyHostToDevice, s->stream));
  BH_CUDA(cudaStreamSynchronize(s->stream));
  s->skin_max_joint = jmax;
  return BH_OK;
}

int bh_skin_roots(bh_sim* s, const float* dq_palette, int njoints) {
  if (!s || !dq_palette || njoints <= 0) return fail(BH_ERR_INVALID, "bh_skin_roots: bad argument");
  if (!s->skin_rest3) return fail(BH_ERR_NOT_INITIALIZED, "bh_skin_roots: call bh_set_skin first");
  if (!s->initialized) return fail(BH_ERR_NOT_INITIALIZED, "bh_skin_roots: no strand state");
  if (s->skin_max_joint >= njoints) return fail(BH_ERR_INVALID, "bh_skin_roots: a joint index of bh_set_skin is outside the palette (njoints too small)");
  DeviceGuard g(s->device);
  if (njoints > s->skin_dq_cap) {
    cudaFree(s->skin_dq); s->skin_dq = nullptr; s->skin_dq_cap = 0;
    BH_CUDA(cudaMalloc(&s->skin_dq, sizeof(float) * 8 * (size_t)njoints));
    s->skin_dq_cap = njoints;
  }
  BH_CUDA(cudaMemcpyAsync(s->skin_dq, dq_palette, sizeof(float) * 8 * (size_t)njoints, cudaMemcpyHostToDevice, s->stream));
  GlMapped gl(s); if (gl.rc) return gl.rc;
  BH_CUDA(bh::launch_skin_roots_dq(s->skin_rest3, s->skin_joints4, s->skin_weights3, s->skin_dq, s->nstrands, s->nverts,
                                   s->planes[BH_PLANE_POSITION], s->stream));
  s->launches += 1;
  if (int rc = gl.done()) return rc;
  // the palette is pageable host memory owned by the caller: do not return before it was consumed
  BH_CUDA(cudaStreamSynchronize(s->stream));
  return BH_OK;
}

int bh_tess_set_patches(bh_sim* s, const int32_t* patch_indices, int64_t nelems) {
  if (!s || !patch_indices || nelems <= 0 || nelems % 6 != 0) return fail(BH_ERR_INVALID, "bh_tess_set_patches: need 6 indices per patch");
  for (int64_t q = 0; q < nelems; ++q)
    if (patch_indices[q] < 0 || patch_indices[q] >= s->nvertices) return fail(BH_ERR_INVALID, "bh_tess_set_patches: index outside the vertex range");
  DeviceGuard g(s->device);
  cudaFree(s->tess_patch); s->tess_patch = nullptr; s->tess_npatches = 0;
  BH_CUDA(cudaMalloc(&s->tess_patch, sizeof(int) * (size_t)nelems));
  BH_CUDA(cudaMemcpyAsync(s->tess_patch, patch_indices, sizeof(int) * (size_t)nelems, cudaMemcpyHostToDevice, s->stream));
  BH_CUDA(cudaStreamSynchronize(s->stream));
  s->tess_npatches = nelems / 6;
  return BH_OK;
}

int64_t bh_tess_stream_count(const bh_sim* s, const bh_tess_params* t) {
  if (!s || !t || t->ninstances < 1 || t->nlines < 1 || t->nsubsegments < 1) return -1;
  return s->tess_npatches * t->ninstances * t->nlines * t->nsubsegments * 2;
}

int bh_tess_stream(bh_sim* s, const bh_tess_params* t, float* out4_host) {
  if (!s || !t) return fail(BH_ERR_INVALID, "bh_tess_stream: NULL argument");
  if (!s->initialized) return fail(BH_ERR_NOT_INITIALIZED, "bh_tess_stream: no strand state");
  if (!s->tess_patch) return fail(BH_ERR_NOT_INITIALIZED, "bh_tess_stream: call bh_tess_set_patches first");
  const int64_t count = bh_tess_stream_count(s, t);
  if (count < 0) return fail(BH_ERR_INVALID, "bh_tess_stream: ninstances, nlines and nsubsegments must be >= 1");
  DeviceGuard g(s->device);
  if (count > s->tess_out_cap) {
    cudaFree(s->tess_out); s->tess_out = nullptr; s->tess_out_cap = 0;
    BH_CUDA(cudaMalloc(&s->tess_out, sizeof(float4) * (size_t)count));
    s->tess_out_cap = count;
  }
  GlMapped gl(s); if (gl.rc) return gl.rc;
  BH_CUDA(bh::launch_tess_stream(s->planes[BH_PLANE_POSITION], s->planes[BH_PLANE_TANGENT], s->tess_patch, s->tess_npatches, s->nverts,
                                 s->params.scale, t->ninstances, t->nlines, t->nsubsegments, t->seed, s->tess_out, s->stream));
  s->launches += 1;
  s->tess_out_count = count;
  if (int rc = gl.done()) return rc;
  if (out4_host) {
    BH_CUDA(cudaMemcpyAsync(out4_host, s->tess_out, sizeof(float4) * (size_t)count, cudaMemcpyDeviceToHost, s->stream));
    BH_CUDA(cudaStreamSynchronize(s->stream));
  }
  return BH_OK;
}

int bh_tess_device_buffer(bh_sim* s, void** device_ptr, int64_t* count) {
  if (!s || !device_ptr) return fail(BH_ERR_INVALID, "bh_tess_device_buffer: NULL argument");
  *device_ptr = s->tess_out;
  if (count) *count = s->tess_out_count;
  return BH_OK;
}

namespace {
// Second half of a registration, common to both kinds of shared buffer: move the current state into it so that the VAO of
// hair.cc:371-389 (or the caller) sees it; on any failure give the buffer back and keep stepping our own.
int adopt_shared_buffer(bh_sim* s, const char* what) {
  float4* own[BH_NUM_PLANES] = { s->buffer0, s->buffer0 + s->nvertices, s->buffer0 + 2 * s->nvertices };
  auto give_back = [&] {
    if (s->gl_resource) { cudaGraphicsUnregisterResource(s->gl_resource); s->gl_resource = nullptr; }
    s->ext_buffer = nullptr; s->ext_bytes = 0;
    for (int p = 0; p < BH_NUM_PLANES; ++p) s->planes[p] = own[p];
  };
  int rc = map_gl(s);
  if (rc) { give_back(); return rc; }
  const cudaError_t e = cudaMemcpyAsync(s->planes[0], s->buffer0, (size_t)BH_NUM_PLANES * s->nvertices * sizeof(float4), cudaMemcpyDeviceToDevice, s->stream);
  rc = unmap_gl(s);
  if (e != cudaSuccess || rc) { (void)cudaGetLastError(); give_back(); return e != cudaSuccess ? fail(BH_ERR_CUDA, what, e) : rc; }
  return BH_OK;
}
}  // namespace

int bh_register_gl_buffer(bh_sim* s, unsigned int gl_buffer) {
  if (!s) return fail(BH_ERR_INVALID, "bh_register_gl_buffer: sim is NULL");
  if (shared_buffer(s)) return fail(BH_ERR_INVALID, "bh_register_gl_buffer: a buffer is already registered");
  DeviceGuard g(s->device);
  cudaGraphicsResource* res = nullptr;
  BH_CUDA(cudaGraphicsGLRegisterBuffer(&res, gl_buffer, 0 /* cudaGraphicsRegisterFlagsNone: read + write */));
  s->gl_resource = res;
  return adopt_shared_buffer(s, "bh_register_gl_buffer: copy into the GL buffer");
}

int bh_register_device_buffer(bh_sim* s, void* device_ptr, uint64_t nbytes) {
  if (!s || !device_ptr) return fail(BH_ERR_INVALID, "bh_register_device_buffer: NULL argument");
  if (shared_buffer(s)) return fail(BH_ERR_INVALID, "bh_register_device_buffer: a buffer is already registered");
  if (nbytes < (uint64_t)BH_NUM_PLANES * s->nvertices * sizeof(float4)) return fail(BH_ERR_INVALID, "bh_register_device_buffer: buffer smaller than 3 planes");
  if (reinterpret_cast<uintptr_t>(device_ptr) % 128) return fail(BH_ERR_INVALID, "bh_register_device_buffer: the buffer must be 128-byte aligned (TMA tiles)");
  DeviceGuard g(s->device);
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, device_ptr) != cudaSuccess || at.type != cudaMemoryTypeDevice || at.device != s->device) {
    (void)cudaGetLastError();
    return fail(BH_ERR_INVALID, "bh_register_device_buffer: not a device allocation of the sim's GPU");
  }
  s->ext_buffer = device_ptr; s->ext_bytes = (size_t)nbytes;
  return adopt_shared_buffer(s, "bh_register_device_buffer: copy into the buffer");
}

int bh_unregister_gl_buffer(bh_sim* s) {
  if (!s) return fail(BH_ERR_INVALID, "bh_unregister_gl_buffer: sim is NULL");
  if (!shared_buffer(s)) return BH_OK;
  DeviceGuard g(s->device);
  int rc = map_gl(s);
  if (rc == BH_OK) {
    cudaMemcpyAsync(s->buffer0, s->planes[0], (size_t)BH_NUM_PLANES * s->nvertices * sizeof(float4), cudaMemcpyDeviceToDevice, s->stream);
    unmap_gl(s);
    cudaStreamSynchronize(s->stream);
  }
  if (s->gl_resource) cudaGraphicsUnregisterResource(s->gl_resource);
  s->gl_resource = nullptr; s->ext_buffer = nullptr; s->ext_bytes = 0;
  for (int p = 0; p < BH_NUM_PLANES; ++p) s->planes[p] = s->buffer0 + (size_t)p * s->nvertices;
  return BH_OK;
}
int bh_unregister_device_buffer(bh_sim* s) { return bh_unregister_gl_buffer(s); }

int bh_buffer_map_stats(const bh_sim* s, int64_t* maps, int64_t* unmaps, int* mapped_now) {
  if (!s) return fail(BH_ERR_INVALID, "bh_buffer_map_stats: sim is NULL");
  if (maps) *maps = s->shared_maps;
  if (unmaps) *unmaps = s->shared_unmaps;
  if (mapped_now) *mapped_now = s->shared_mapped ? 1 : 0;
  return BH_OK;
}

}  // extern "C"
