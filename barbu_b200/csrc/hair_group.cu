// hair_group.cu — one scalp sharded over several GPUs behind ONE handle, driven from ONE host thread (SURVEY.md §8b
// "Threading", §8e): what the reference's single-threaded frame loop (src/core/app.cc:60-81, core/renderer.cc:69-81) can
// call without processes or torch.
//
// Strands never interact (cs_simulation.glsl reads and writes only its own workgroup's strand), so shard g owns the
// contiguous global strand range [S*g/G, S*(g+1)/G) on devices[g], is stepped by asynchronous launches on its own stream,
// and no step ever exchanges data. The one exchange the path knows is the optional gather of a plane of buffer 0 to the
// render GPU (the VAO of hair.cc:371-389 reads positions at offset 0 and tangents at 2*V*16): peer copies pushed by every
// source GPU over NVLink/NVSwitch (cudaMemcpyPeerAsync on the source's stream, so a shard's copy starts the moment its
// own step is done and overlaps the steps still running on the other GPUs), into a device buffer or a registered GL buffer.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "hair_sim.cuh"

extern "C" cudaError_t cudaGraphicsGLRegisterBuffer(struct cudaGraphicsResource** resource, unsigned int buffer, unsigned int flags);

using bh::DeviceGuard;
using bh::fail;

struct bh_group {
  std::vector<bh_sim*> shard;
  std::vector<int64_t> first;            // G + 1 entries: shard g = strands [first[g], first[g + 1])
  int64_t nstrands = 0;
  int nverts = 0;
  std::vector<cudaEvent_t> ev0, ev1;     // per shard, on its device: timed regions
  // gather target on the render device
  int gather_device = -1;
  float4* gather_buf = nullptr;          // up to 3 planes, PingPongBuffer layout: plane p at p * V float4
  cudaStream_t gather_stream = nullptr;
  cudaGraphicsResource* gl_resource = nullptr;
  int gl_device = -1;
};

namespace {

int64_t V_of(const bh_group* g) { return g->nstrands * (int64_t)g->nverts; }

void enable_peers(const std::vector<int>& devs) {
  for (int a : devs)
    for (int b : devs) {
      if (a == b) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, a, b) != cudaSuccess || !can) { (void)cudaGetLastError(); continue; }
      DeviceGuard g(a);
      cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
      if (e != cudaSuccess) (void)cudaGetLastError();                       // already enabled: fine; otherwise copies are staged by the driver
    }
}

int ensure_gather_target(bh_group* g, int device) {
  if (g->gather_buf && g->gather_device == device) return BH_OK;
  if (g->gather_buf) { DeviceGuard d(g->gather_device); cudaFree(g->gather_buf); cudaStreamDestroy(g->gather_stream); g->gather_buf = nullptr; g->gather_stream = nullptr; }
  DeviceGuard d(device);
  if (!d.ok) return fail(BH_ERR_INVALID, "bh_group_gather_plane: no such CUDA device");
  BH_CUDA(cudaMalloc(&g->gather_buf, (size_t)BH_NUM_PLANES * V_of(g) * sizeof(float4)));
  BH_CUDA(cudaStreamCreateWithFlags(&g->gather_stream, cudaStreamNonBlocking));
  g->gather_device = device;
  return BH_OK;
}

// Pushes plane `plane` of every shard into dst (plane base on dst_device) and makes `after` (a stream of dst_device) wait
// for the copies. ms_max: max over shards of the copy's device time, or NULL.
int gather_into(bh_group* g, int plane, float4* dst_plane, int dst_device, cudaStream_t after, float* ms_max) {
  const int G = (int)g->shard.size();
  for (int q = 0; q < G; ++q) {
    bh_sim* s = g->shard[q];
    if (!s->initialized) return fail(BH_ERR_NOT_INITIALIZED, "bh_group_gather_plane: a shard has no strand state");
    if (s->gl_resource || s->ext_buffer) return fail(BH_ERR_UNSUPPORTED, "bh_group_gather_plane: shard buffers registered with GL one by one cannot be gathered");
    DeviceGuard d(s->device);
    const size_t off = (size_t)g->first[q] * g->nverts, bytes = (size_t)(g->first[q + 1] - g->first[q]) * g->nverts * sizeof(float4);
    BH_CUDA(cudaEventRecord(g->ev0[q], s->stream));
    if (bytes) BH_CUDA(cudaMemcpyPeerAsync(dst_plane + off, dst_device, s->planes[plane], s->device, bytes, s->stream));
    BH_CUDA(cudaEventRecord(g->ev1[q], s->stream));
  }
  {
    DeviceGuard d(dst_device);
    for (int q = 0; q < G; ++q) BH_CUDA(cudaStreamWaitEvent(after, g->ev1[q], 0));
  }
  if (ms_max) {
    *ms_max = 0.0f;
    for (int q = 0; q < G; ++q) {
      DeviceGuard d(g->shard[q]->device);
      BH_CUDA(cudaEventSynchronize(g->ev1[q]));
      float ms = 0.0f;
      BH_CUDA(cudaEventElapsedTime(&ms, g->ev0[q], g->ev1[q]));
      if (ms > *ms_max) *ms_max = ms;
    }
  }
  return BH_OK;
}

}  // namespace

extern "C" {

int bh_group_create(bh_group** out, const int* devices, int ndevices, int64_t nstrands, int nverts) {
  if (!out) return fail(BH_ERR_INVALID, "bh_group_create: out is NULL");
  *out = nullptr;
  if (!devices || ndevices < 1 || ndevices > 64) return fail(BH_ERR_INVALID, "bh_group_create: need 1..64 devices");
  if (nstrands < ndevices || nverts <= 0) return fail(BH_ERR_INVALID, "bh_group_create: need at least one strand per shard and nverts > 0");
  bh_group* g = new (std::nothrow) bh_group();
  if (!g) return fail(BH_ERR_INVALID, "bh_group_create: out of host memory");
  g->nstrands = nstrands; g->nverts = nverts;
  g->first.resize(ndevices + 1);
  for (int q = 0; q <= ndevices; ++q) g->first[q] = (int64_t)(((__int128)nstrands * q) / ndevices);   // barbu_b200/shard.py:shard_range
  for (int q = 0; q < ndevices; ++q) {
    bh_sim* s = nullptr;
    int rc = bh_create(&s, g->first[q + 1] - g->first[q], nverts, devices[q]);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (rc == BH_OK) {
      DeviceGuard d(devices[q]);
      if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { (void)cudaGetLastError(); rc = fail(BH_ERR_CUDA, "bh_group_create: cudaEventCreate"); }
    }
    if (rc != BH_OK) { if (s) bh_destroy(s); bh_group_destroy(g); return rc; }
    g->shard.push_back(s); g->ev0.push_back(e0); g->ev1.push_back(e1);
  }
  enable_peers(std::vector<int>(devices, devices + ndevices));
  *out = g;
  return BH_OK;
}

int bh_group_destroy(bh_group* g) {
  if (!g) return BH_OK;
  if (g->gl_resource) { DeviceGuard d(g->gl_device); cudaGraphicsUnregisterResource(g->gl_resource); }
  if (g->gather_buf) { DeviceGuard d(g->gather_device); cudaFree(g->gather_buf); if (g->gather_stream) cudaStreamDestroy(g->gather_stream); }
  for (size_t q = 0; q < g->shard.size(); ++q) {
    { DeviceGuard d(g->shard[q]->device); if (g->ev0[q]) cudaEventDestroy(g->ev0[q]); if (g->ev1[q]) cudaEventDestroy(g->ev1[q]); }
    bh_destroy(g->shard[q]);
  }
  delete g;
  return BH_OK;
}

int bh_group_size(const bh_group* g) { return g ? (int)g->shard.size() : 0; }

bh_sim* bh_group_shard(bh_group* g, int q) { return (g && q >= 0 && q < (int)g->shard.size()) ? g->shard[q] : nullptr; }

int bh_group_shard_range(const bh_group* g, int q, int64_t* first, int64_t* count) {
  if (!g || q < 0 || q >= (int)g->shard.size()) return fail(BH_ERR_INVALID, "bh_group_shard_range: bad argument");
  if (first) *first = g->first[q];
  if (count) *count = g->first[q + 1] - g->first[q];
  return BH_OK;
}

int bh_group_set_params(bh_group* g, const bh_params* p) {
  if (!g) return fail(BH_ERR_INVALID, "bh_group_set_params: group is NULL");
  for (bh_sim* s : g->shard) { int rc = bh_set_params(s, p); if (rc) return rc; }
  return BH_OK;
}

int bh_group_set_bounding_sphere(bh_group* g, const float sphere[4]) {
  if (!g) return fail(BH_ERR_INVALID, "bh_group_set_bounding_sphere: group is NULL");
  for (bh_sim* s : g->shard) { int rc = bh_set_bounding_sphere(s, sphere); if (rc) return rc; }
  return BH_OK;
}

int bh_group_init_sphere_scalp(bh_group* g, int rows, int cols, int order, unsigned seed, float maxlength) {
  if (!g) return fail(BH_ERR_INVALID, "bh_group_init_sphere_scalp: group is NULL");
  if ((int64_t)rows * cols != g->nstrands) return fail(BH_ERR_INVALID, "bh_group_init_sphere_scalp: rows * cols must equal the group's strand count");
  std::vector<float> rv;
  for (size_t q = 0; q < g->shard.size(); ++q) {
    const int64_t count = g->first[q + 1] - g->first[q];
    rv.resize((size_t)count);
    int rc = bh_random_values(seed, g->first[q], count, rv.data());
    if (rc == BH_OK) rc = bh_init_sphere_scalp_ordered(g->shard[q], rows, cols, order, g->first[q], rv.data(), maxlength);
    if (rc) return rc;
  }
  return BH_OK;
}

int bh_group_init_strands(bh_group* g, const float* root_pos3, const float* root_nrm3, const float* random_value, float maxlength) {
  if (!g || !root_pos3 || !root_nrm3 || !random_value) return fail(BH_ERR_INVALID, "bh_group_init_strands: NULL argument");
  for (size_t q = 0; q < g->shard.size(); ++q) {
    const size_t f = (size_t)g->first[q];
    int rc = bh_init_strands(g->shard[q], root_pos3 + 3 * f, root_nrm3 + 3 * f, random_value + f, maxlength);
    if (rc) return rc;
  }
  return BH_OK;
}

int bh_group_upload(bh_group* g, const float* pos4, const float* vel4, const float* tan4) {
  if (!g) return fail(BH_ERR_INVALID, "bh_group_upload: group is NULL");
  for (size_t q = 0; q < g->shard.size(); ++q) {
    const size_t off = 4 * (size_t)g->first[q] * g->nverts;
    int rc = bh_upload(g->shard[q], pos4 ? pos4 + off : nullptr, vel4 ? vel4 + off : nullptr, tan4 ? tan4 + off : nullptr);
    if (rc) return rc;
  }
  return BH_OK;
}

int bh_group_download(bh_group* g, float* pos4, float* vel4, float* tan4) {
  if (!g) return fail(BH_ERR_INVALID, "bh_group_download: group is NULL");
  for (size_t q = 0; q < g->shard.size(); ++q) {
    const size_t off = 4 * (size_t)g->first[q] * g->nverts;
    int rc = bh_download(g->shard[q], pos4 ? pos4 + off : nullptr, vel4 ? vel4 + off : nullptr, tan4 ? tan4 + off : nullptr);
    if (rc) return rc;
  }
  return BH_OK;
}

int bh_group_step(bh_group* g, float dt, int substeps) {
  if (!g) return fail(BH_ERR_INVALID, "bh_group_step: group is NULL");
  for (bh_sim* s : g->shard) { int rc = bh_step(s, dt, substeps); if (rc) return rc; }   // asynchronous: G streams run side by side
  return BH_OK;
}

int bh_group_set_substep_fusion(bh_group* g, int enabled) {
  if (!g) return fail(BH_ERR_INVALID, "bh_group_set_substep_fusion: group is NULL");
  for (bh_sim* s : g->shard) { int rc = bh_set_substep_fusion(s, enabled); if (rc) return rc; }
  return BH_OK;
}

int bh_group_synchronize(bh_group* g) {
  if (!g) return fail(BH_ERR_INVALID, "bh_group_synchronize: group is NULL");
  for (bh_sim* s : g->shard) { int rc = bh_synchronize(s); if (rc) return rc; }
  return BH_OK;
}

int bh_group_step_timed(bh_group* g, float dt, int substeps, int frames, float* ms_max, float* ms_per_shard) {
  if (!g || frames < 1) return fail(BH_ERR_INVALID, "bh_group_step_timed: bad argument");
  int rc = bh_group_synchronize(g); if (rc) return rc;                        // the barrier in front of the timed region
  const int G = (int)g->shard.size();
  for (int q = 0; q < G; ++q) { DeviceGuard d(g->shard[q]->device); BH_CUDA(cudaEventRecord(g->ev0[q], g->shard[q]->stream)); }
  for (int f = 0; f < frames; ++f) { rc = bh_group_step(g, dt, substeps); if (rc) return rc; }   // frame by frame, shard by shard: no GPU waits for its launches
  for (int q = 0; q < G; ++q) { DeviceGuard d(g->shard[q]->device); BH_CUDA(cudaEventRecord(g->ev1[q], g->shard[q]->stream)); }
  float mx = 0.0f;
  for (int q = 0; q < G; ++q) {
    DeviceGuard d(g->shard[q]->device);
    BH_CUDA(cudaEventSynchronize(g->ev1[q]));
    float ms = 0.0f;
    BH_CUDA(cudaEventElapsedTime(&ms, g->ev0[q], g->ev1[q]));
    if (ms_per_shard) ms_per_shard[q] = ms;
    if (ms > mx) mx = ms;
  }
  if (ms_max) *ms_max = mx;
  return BH_OK;
}

int64_t bh_group_launch_count(const bh_group* g) {
  int64_t n = 0;
  if (g) for (const bh_sim* s : g->shard) n += bh_launch_count(s);
  return n;
}

int bh_group_gather_plane(bh_group* g, int plane, int dst_device, void** device_ptr, float* ms) {
  if (!g || plane < 0 || plane >= BH_NUM_PLANES) return fail(BH_ERR_INVALID, "bh_group_gather_plane: bad argument");
  int ndev = 0;
  BH_CUDA(cudaGetDeviceCount(&ndev));
  if (dst_device < 0 || dst_device >= ndev) return fail(BH_ERR_INVALID, "bh_group_gather_plane: no such CUDA device");
  int rc = ensure_gather_target(g, dst_device); if (rc) return rc;
  float4* dst = g->gather_buf + (size_t)plane * V_of(g);
  rc = gather_into(g, plane, dst, dst_device, g->gather_stream, ms); if (rc) return rc;
  { DeviceGuard d(dst_device); BH_CUDA(cudaStreamSynchronize(g->gather_stream)); }
  if (device_ptr) *device_ptr = dst;
  return BH_OK;
}

int bh_group_register_gl_buffer(bh_group* g, unsigned int gl_buffer, int render_device) {
  if (!g) return fail(BH_ERR_INVALID, "bh_group_register_gl_buffer: group is NULL");
  if (g->gl_resource) return fail(BH_ERR_INVALID, "bh_group_register_gl_buffer: already registered");
  int ndev = 0;
  BH_CUDA(cudaGetDeviceCount(&ndev));
  if (render_device < 0 || render_device >= ndev) return fail(BH_ERR_INVALID, "bh_group_register_gl_buffer: no such CUDA device");
  DeviceGuard d(render_device);
  cudaGraphicsResource* res = nullptr;
  BH_CUDA(cudaGraphicsGLRegisterBuffer(&res, gl_buffer, 0));
  g->gl_resource = res; g->gl_device = render_device;
  return BH_OK;
}

int bh_group_unregister_gl_buffer(bh_group* g) {
  if (!g) return fail(BH_ERR_INVALID, "bh_group_unregister_gl_buffer: group is NULL");
  if (!g->gl_resource) return BH_OK;
  DeviceGuard d(g->gl_device);
  cudaGraphicsUnregisterResource(g->gl_resource);
  g->gl_resource = nullptr; g->gl_device = -1;
  return BH_OK;
}

int bh_group_gather_to_gl(bh_group* g, unsigned plane_mask) {
  if (!g) return fail(BH_ERR_INVALID, "bh_group_gather_to_gl: group is NULL");
  if (!g->gl_resource) return fail(BH_ERR_NOT_INITIALIZED, "bh_group_gather_to_gl: call bh_group_register_gl_buffer first");
  if (plane_mask == 0) plane_mask = 1u << BH_PLANE_POSITION;
  int rc = ensure_gather_target(g, g->gl_device); if (rc) return rc;       // only for its stream
  DeviceGuard d(g->gl_device);
  BH_CUDA(cudaGraphicsMapResources(1, &g->gl_resource, g->gather_stream));
  void* ptr = nullptr; size_t bytes = 0;
  cudaError_t e = cudaGraphicsResourceGetMappedPointer(&ptr, &bytes, g->gl_resource);
  rc = BH_OK;
  if (e != cudaSuccess) { (void)cudaGetLastError(); rc = fail(BH_ERR_CUDA, "bh_group_gather_to_gl: mapped pointer", e); }
  else if (bytes < (size_t)BH_NUM_PLANES * V_of(g) * sizeof(float4)) rc = fail(BH_ERR_INVALID, "GL buffer smaller than 3 planes of the whole scalp");
  for (int p = 0; p < BH_NUM_PLANES && rc == BH_OK; ++p)
    if (plane_mask >> p & 1u) rc = gather_into(g, p, static_cast<float4*>(ptr) + (size_t)p * V_of(g), g->gl_device, g->gather_stream, nullptr);
  // always unmap (the unmap is ordered after the copies on gather_stream, which waits for every shard's copy)
  cudaError_t e2 = cudaGraphicsUnmapResources(1, &g->gl_resource, g->gather_stream);
  if (e2 != cudaSuccess && rc == BH_OK) { (void)cudaGetLastError(); rc = fail(BH_ERR_CUDA, "bh_group_gather_to_gl: unmap", e2); }
  cudaStreamSynchronize(g->gather_stream);
  return rc;
}

}  // extern "C"
