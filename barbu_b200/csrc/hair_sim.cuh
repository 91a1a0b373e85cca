// hair_sim.cuh — the object behind the opaque `bh_sim` handle of include/barbu_hair.h, shared by the translation
// units that implement the C ABI (hair_capi.cu, hair_state.cu). Internal: never installed, never seen by callers.
#pragma once
#include "../../include/barbu_hair.h"

#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace bh {

constexpr int kHostPipeStreams = 4;

// Records `what` (+ the CUDA error text) as this thread's bh_last_error() and returns `code`.
int fail(int code, const char* what, cudaError_t e = cudaSuccess);

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; }
    if (cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace bh

#define BH_CUDA(expr)                                                       \
  do {                                                                      \
    cudaError_t e__ = (expr);                                               \
    if (e__ != cudaSuccess) { (void)cudaGetLastError(); return bh::fail(BH_ERR_CUDA, #expr, e__); } \
  } while (0)

struct bh_sim {
  int device = 0;
  int64_t nstrands = 0;
  int nverts = 0;
  int64_t nvertices = 0;                 // V = S * N
  float4* buffer0 = nullptr;             // 3 planes, layout of PingPongBuffer buffer 0
  float4* planes[BH_NUM_PLANES] = { nullptr, nullptr, nullptr };
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaStream_t pipe[bh::kHostPipeStreams] = { nullptr, nullptr, nullptr, nullptr };
  bh_params params;
  bool initialized = false;              // Hair::initialized(): state present
  int step_policy = 0;                   // bh_set_step_policy: BH_POLICY_THROUGHPUT / LATENCY / AUTO
  int fuse_substeps = 0;                 // bh_set_substep_fusion: 0 off, 1 where it pays, 2 wherever the shape allows it
  int64_t launches = 0;
  int64_t step_launches = 0;             // launches of bh_step alone: parity = tile direction of the next one
  cudaEvent_t order_event = nullptr;     // bh_set_stream / bh_reset_stream: the new stream waits for what the old one holds
  std::vector<cudaEvent_t> host_events;  // bh_step_host: two per slice + one, created on first use
  unsigned int* tile_counters = nullptr;  // kHostPipeStreams + 1 words: one tile scheduler per stream that may be in flight
  // roots kept for re-generation / skinning ("base normals are kept for potential future uses", hair.cc:262)
  float* root_pos3 = nullptr;
  float* root_nrm3 = nullptr;
  // skinning extension
  float* skin_rest3 = nullptr;
  int* skin_joints4 = nullptr;
  float* skin_weights3 = nullptr;
  float* skin_dq = nullptr;
  int skin_dq_cap = 0;
  int skin_max_joint = 0;                // largest joint index bh_set_skin stored (bh_skin_roots checks njoints against it)
  // tess-stream stage
  int* tess_patch = nullptr;            // device copy of the patch element buffer
  int64_t tess_npatches = 0;
  float4* tess_out = nullptr;           // GL_LINES vertex stream (xyz, relPos)
  int64_t tess_out_cap = 0, tess_out_count = 0;
  // GL interop
  cudaGraphicsResource* gl_resource = nullptr;
  // bh_register_device_buffer: a caller-owned device allocation in the GL buffer's role (same state machine: planes live
  // there, every entry point brackets its work with map / unmap); and the bracket's bookkeeping for both kinds
  void* ext_buffer = nullptr;
  size_t ext_bytes = 0;
  int64_t shared_maps = 0, shared_unmaps = 0;
  bool shared_mapped = false;
  // state files / checksums (hair_state.cu)
  unsigned long long* checksum_words = nullptr;   // 2 device words
};

namespace bh {
// While buffer 0 is a registered GL buffer it is only addressable between these two calls (no-ops otherwise).
int map_gl(bh_sim* s);
int unmap_gl(bh_sim* s);
}  // namespace bh
