// hair_state.cu — strand-state files and device checksums.
//
// The reference has no persistence: its state lives in two GL buffers for the lifetime of the process
// (src/memory/pingpong_buffer.cc:6-33) and is regenerated from rand() at every start (hair.cc:236-361). A state file
// is the same bytes the render path reads — the SoA float4 planes of buffer 0 in PingPongBuffer order
// (pingpong_buffer.cc:16-17,44-48) — behind a fixed little-endian header that also carries the uniforms of
// Hair::update (hair.cc:107-110) and the shard coordinates of §8e, so that a run can be resumed bit for bit, a shard
// written on one GPU can be checked on another, and oracle and device states can be exchanged as files.
//
// The checksum is computed ON THE DEVICE by one streaming pass (16 B per vertex per plane, HBM-bound) and is additive
// over disjoint strand ranges when words are keyed by their GLOBAL index: the checksum of a scalp is the wrapping sum
// of the checksums of its shards, whatever the partition.
#include "hair_sim.cuh"

#include <cstdio>
#include <cstring>

namespace bh {

namespace {

constexpr unsigned long long kGolden = 0x9E3779B97F4A7C15ull;

// c0 = sum of the 32-bit words; c1 = sum of word * ((key + 1) * kGolden), key = plane << 60 | global word index;
// both mod 2^64.
__global__ void __launch_bounds__(256) checksum_kernel(const uint4* __restrict__ plane, long long nvertices,
                                                       unsigned long long key0, unsigned long long* __restrict__ out) {
  unsigned long long c0 = 0, c1 = 0;
  auto add = [&](const uint4 w, const long long v) {
    const unsigned long long k = (key0 + 4ull * (unsigned long long)v + 1ull) * kGolden;   // key of .x, already multiplied
    c0 += (unsigned long long)w.x + w.y + w.z + w.w;
    c1 += w.x * k + w.y * (k + kGolden) + w.z * (k + 2 * kGolden) + w.w * (k + 3 * kGolden);
  };
  // four independent 16-byte loads in flight per thread (a read-only stream: latency is hidden by loads, not by warps)
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; v + 3 * stride < nvertices; v += 4 * stride) {
    const uint4 w0 = __ldcs(plane + v), w1 = __ldcs(plane + v + stride), w2 = __ldcs(plane + v + 2 * stride), w3 = __ldcs(plane + v + 3 * stride);
    add(w0, v); add(w1, v + stride); add(w2, v + 2 * stride); add(w3, v + 3 * stride);
  }
  for (; v < nvertices; v += stride) add(__ldcs(plane + v), v);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    c0 += __shfl_xor_sync(0xffffffffu, c0, d);
    c1 += __shfl_xor_sync(0xffffffffu, c1, d);
  }
  __shared__ unsigned long long s0[8], s1[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s0[warp] = c0; s1[warp] = c1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { c0 += s0[w]; c1 += s1[w]; }
    atomicAdd(out, c0);
    atomicAdd(out + 1, c1);
  }
}

// Checksum of the planes in `mask` of the sim's CURRENT buffer 0 (already mapped), strand 0 being global strand `first`.
int checksum_planes(bh_sim* s, unsigned mask, int64_t first, uint64_t out[2]) {
  if (!s->checksum_words) BH_CUDA(cudaMalloc(&s->checksum_words, 2 * sizeof(unsigned long long)));
  BH_CUDA(cudaMemsetAsync(s->checksum_words, 0, 2 * sizeof(unsigned long long), s->stream));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device);
  long long blocks = (s->nvertices + 255) / 256;
  if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
  for (int p = 0; p < BH_NUM_PLANES; ++p) {
    if (!(mask >> p & 1u)) continue;
    const unsigned long long key0 = ((unsigned long long)p << 60) + 4ull * (unsigned long long)first * (unsigned long long)s->nverts;
    checksum_kernel<<<(unsigned)blocks, 256, 0, s->stream>>>(reinterpret_cast<const uint4*>(s->planes[p]), s->nvertices, key0, s->checksum_words);
    BH_CUDA(cudaGetLastError());
    s->launches += 1;
  }
  unsigned long long h[2];
  BH_CUDA(cudaMemcpyAsync(h, s->checksum_words, sizeof h, cudaMemcpyDeviceToHost, s->stream));
  BH_CUDA(cudaStreamSynchronize(s->stream));
  out[0] = h[0]; out[1] = h[1];
  return BH_OK;
}

// ---- the file ----------------------------------------------------------------------------------------------------
constexpr char kMagic[8] = { 'B', 'A', 'R', 'B', 'U', 'H', 'S', '1' };
constexpr uint32_t kVersion = 1, kHeaderBytes = 512;

#pragma pack(push, 1)
struct FileHeader {                 // little-endian, offsets in include/barbu_hair.h
  char magic[8];
  uint32_t version, header_bytes;
  int64_t nstrands;
  int32_t nverts;
  uint32_t plane_mask;
  int64_t total_strands, first_strand, frame;
  float dt;
  uint32_t seed;
  uint64_t checksum[2];
  uint32_t params_bytes, reserved;
  bh_params params;
};
#pragma pack(pop)
static_assert(sizeof(bh_params) == 292, "bh_params is part of the state-file format: bump kVersion when it changes");
static_assert(sizeof(FileHeader) == 88 + 292 && sizeof(FileHeader) <= kHeaderBytes, "header layout");

int popcount3(unsigned m) { return (int)(m & 1u) + (int)(m >> 1 & 1u) + (int)(m >> 2 & 1u); }

// Reads and validates the header; on success the stream is positioned at the first plane.
int read_header(FILE* f, const char* path, FileHeader* h) {
  unsigned char raw[kHeaderBytes];
  if (std::fread(raw, 1, kHeaderBytes, f) != kHeaderBytes) return fail(BH_ERR_INVALID, "state file: shorter than its header");
  std::memcpy(h, raw, sizeof *h);
  if (std::memcmp(h->magic, kMagic, 8) != 0) return fail(BH_ERR_INVALID, "state file: bad magic (not a BARBUHS1 file)");
  if (h->version != kVersion) return fail(BH_ERR_UNSUPPORTED, "state file: unknown version");
  if (h->header_bytes != kHeaderBytes || h->params_bytes != sizeof(bh_params)) return fail(BH_ERR_INVALID, "state file: header size mismatch");
  if (h->nstrands <= 0 || h->nverts <= 0 || (h->plane_mask & ~7u) || h->plane_mask == 0) return fail(BH_ERR_INVALID, "state file: bad shape");
  if (h->first_strand < 0 || h->total_strands < h->first_strand + h->nstrands) return fail(BH_ERR_INVALID, "state file: shard outside the scalp");
  if (h->params.ncapsules < 0 || h->params.ncapsules > BH_MAX_CAPSULES || h->params.iterations < 0 ||
      (h->params.math != BH_MATH_EXACT && h->params.math != BH_MATH_FAST)) return fail(BH_ERR_INVALID, "state file: bad parameters");
  // the payload must be there in full
  const long pos = std::ftell(f);
  if (std::fseek(f, 0, SEEK_END) != 0) return fail(BH_ERR_INVALID, "state file: not seekable");
  const long long size = std::ftell(f);
  const long long want = (long long)kHeaderBytes + (long long)popcount3(h->plane_mask) * h->nstrands * h->nverts * 16;
  if (size != want) return fail(BH_ERR_INVALID, "state file: truncated or trailing bytes (size does not match the header)");
  std::fseek(f, pos, SEEK_SET);
  (void)path;
  return BH_OK;
}

void fill_info(const FileHeader& h, bh_state_info* info) {
  info->nstrands = h.nstrands; info->nverts = h.nverts; info->plane_mask = h.plane_mask;
  info->total_strands = h.total_strands; info->first_strand = h.first_strand; info->frame = h.frame;
  info->dt = h.dt; info->seed = h.seed; info->checksum[0] = h.checksum[0]; info->checksum[1] = h.checksum[1];
  info->params = h.params;
}

constexpr size_t kStageBytes = 32u << 20;   // two pinned staging buffers: the copy of chunk i+1 overlaps the file I/O of chunk i

struct Staging {
  void* buf[2] = { nullptr, nullptr };
  cudaEvent_t ev[2] = { nullptr, nullptr };
  cudaError_t init() {
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
      e = cudaMallocHost(&buf[i], kStageBytes);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
    }
    return e;
  }
  ~Staging() { for (int i = 0; i < 2; ++i) { if (buf[i]) cudaFreeHost(buf[i]); if (ev[i]) cudaEventDestroy(ev[i]); } }
};

}  // namespace
}  // namespace bh

using bh::fail;
using bh::DeviceGuard;

extern "C" {

int bh_state_checksum(bh_sim* s, unsigned plane_mask, int64_t first_strand, uint64_t out[2]) {
  if (!s || !out) return fail(BH_ERR_INVALID, "bh_state_checksum: NULL argument");
  if (plane_mask == 0) plane_mask = 7u;
  if ((plane_mask & ~7u) || first_strand < 0) return fail(BH_ERR_INVALID, "bh_state_checksum: bad plane mask or first strand");
  if (!s->initialized) return fail(BH_ERR_NOT_INITIALIZED, "bh_state_checksum: no strand state");
  DeviceGuard g(s->device);
  int rc = bh::map_gl(s); if (rc) return rc;
  rc = bh::checksum_planes(s, plane_mask, first_strand, out);
  const int rc2 = bh::unmap_gl(s);
  return rc ? rc : rc2;
}

int bh_peek_state(const char* path, bh_state_info* info) {
  if (!path || !info) return fail(BH_ERR_INVALID, "bh_peek_state: NULL argument");
  FILE* f = std::fopen(path, "rb");
  if (!f) return fail(BH_ERR_INVALID, "bh_peek_state: cannot open the file");
  bh::FileHeader h;
  const int rc = bh::read_header(f, path, &h);
  std::fclose(f);
  if (rc) return rc;
  bh::fill_info(h, info);
  return BH_OK;
}

int bh_save_state(bh_sim* s, const char* path, const bh_state_info* info) {
  if (!s || !path) return fail(BH_ERR_INVALID, "bh_save_state: NULL argument");
  if (!s->initialized) return fail(BH_ERR_NOT_INITIALIZED, "bh_save_state: no strand state");
  bh::FileHeader h;
  std::memset(&h, 0, sizeof h);
  std::memcpy(h.magic, bh::kMagic, 8);
  h.version = bh::kVersion; h.header_bytes = bh::kHeaderBytes;
  h.nstrands = s->nstrands; h.nverts = s->nverts;
  h.plane_mask = (info && info->plane_mask) ? info->plane_mask : 7u;
  h.total_strands = (info && info->total_strands) ? info->total_strands : s->nstrands;
  h.first_strand = info ? info->first_strand : 0;
  h.frame = info ? info->frame : 0;
  h.dt = info ? info->dt : 0.0f;
  h.seed = info ? info->seed : 0u;
  h.params_bytes = sizeof(bh_params);
  h.params = s->params;
  if ((h.plane_mask & ~7u) || h.first_strand < 0 || h.total_strands < h.first_strand + h.nstrands)
    return fail(BH_ERR_INVALID, "bh_save_state: bad plane mask or shard coordinates");
  DeviceGuard g(s->device);
  bh::Staging st;
  BH_CUDA(st.init());
  int rc = bh::map_gl(s); if (rc) return rc;
  rc = bh::checksum_planes(s, h.plane_mask, h.first_strand, h.checksum);
  if (rc) { bh::unmap_gl(s); return rc; }
  FILE* f = std::fopen(path, "wb");
  if (!f) { bh::unmap_gl(s); return fail(BH_ERR_INVALID, "bh_save_state: cannot create the file"); }
  unsigned char raw[bh::kHeaderBytes];
  std::memset(raw, 0, sizeof raw);
  std::memcpy(raw, &h, sizeof h);
  bool io_ok = std::fwrite(raw, 1, sizeof raw, f) == sizeof raw;
  cudaError_t e = cudaSuccess;
  const size_t plane_bytes = (size_t)s->nvertices * sizeof(float4);
  for (int p = 0; p < BH_NUM_PLANES && io_ok && e == cudaSuccess; ++p) {
    if (!(h.plane_mask >> p & 1u)) continue;
    const char* src = reinterpret_cast<const char*>(s->planes[p]);
    const size_t nchunks = (plane_bytes + bh::kStageBytes - 1) / bh::kStageBytes;
    auto chunk_bytes = [&](size_t c) { return (c + 1 == nchunks) ? plane_bytes - c * bh::kStageBytes : bh::kStageBytes; };
    e = cudaMemcpyAsync(st.buf[0], src, chunk_bytes(0), cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaEventRecord(st.ev[0], s->stream);
    for (size_t c = 0; c < nchunks && io_ok && e == cudaSuccess; ++c) {
      if (c + 1 < nchunks) {
        e = cudaMemcpyAsync(st.buf[(c + 1) & 1], src + (c + 1) * bh::kStageBytes, chunk_bytes(c + 1), cudaMemcpyDeviceToHost, s->stream);
        if (e == cudaSuccess) e = cudaEventRecord(st.ev[(c + 1) & 1], s->stream);
      }
      if (e == cudaSuccess) e = cudaEventSynchronize(st.ev[c & 1]);
      if (e == cudaSuccess) io_ok = std::fwrite(st.buf[c & 1], 1, chunk_bytes(c), f) == chunk_bytes(c);
    }
  }
  cudaStreamSynchronize(s->stream);
  io_ok = (std::fclose(f) == 0) && io_ok;
  rc = bh::unmap_gl(s);
  if (e != cudaSuccess) { (void)cudaGetLastError(); std::remove(path); return fail(BH_ERR_CUDA, "bh_save_state: device to host copy", e); }
  if (!io_ok) { std::remove(path); return fail(BH_ERR_INVALID, "bh_save_state: write failed"); }
  return rc;
}

int bh_load_state(bh_sim* s, const char* path, bh_state_info* info) {
  if (!s || !path) return fail(BH_ERR_INVALID, "bh_load_state: NULL argument");
  FILE* f = std::fopen(path, "rb");
  if (!f) return fail(BH_ERR_INVALID, "bh_load_state: cannot open the file");
  bh::FileHeader h;
  int rc = bh::read_header(f, path, &h);
  if (rc == BH_OK && (h.nstrands != s->nstrands || h.nverts != s->nverts))
    rc = fail(BH_ERR_INVALID, "bh_load_state: the file holds a different shape than this sim (see bh_peek_state)");
  if (rc) { std::fclose(f); return rc; }
  DeviceGuard g(s->device);
  bh::Staging st;
  cudaError_t e = st.init();
  if (e != cudaSuccess) { std::fclose(f); (void)cudaGetLastError(); return fail(BH_ERR_CUDA, "bh_load_state: staging buffers", e); }
  rc = bh::map_gl(s);
  if (rc) { std::fclose(f); return rc; }
  const bool had_state = s->initialized;
  s->initialized = false;                       // until the payload is in and verified
  bool io_ok = true;
  const size_t plane_bytes = (size_t)s->nvertices * sizeof(float4);
  size_t k = 0;                                 // running chunk number: picks the staging buffer
  for (int p = 0; p < BH_NUM_PLANES && io_ok && e == cudaSuccess; ++p) {
    if (!(h.plane_mask >> p & 1u)) continue;
    char* dst = reinterpret_cast<char*>(s->planes[p]);
    for (size_t off = 0; off < plane_bytes && io_ok && e == cudaSuccess; off += bh::kStageBytes, ++k) {
      const size_t n = (plane_bytes - off < bh::kStageBytes) ? plane_bytes - off : bh::kStageBytes;
      if (k >= 2) e = cudaEventSynchronize(st.ev[k & 1]);          // the copy that last used this buffer is done
      if (e != cudaSuccess) break;
      io_ok = std::fread(st.buf[k & 1], 1, n, f) == n;
      if (!io_ok) break;
      e = cudaMemcpyAsync(dst + off, st.buf[k & 1], n, cudaMemcpyHostToDevice, s->stream);
      if (e == cudaSuccess) e = cudaEventRecord(st.ev[k & 1], s->stream);
    }
  }
  std::fclose(f);
  cudaStreamSynchronize(s->stream);
  if (e != cudaSuccess) { bh::unmap_gl(s); (void)cudaGetLastError(); return fail(BH_ERR_CUDA, "bh_load_state: host to device copy", e); }
  if (!io_ok) { bh::unmap_gl(s); return fail(BH_ERR_INVALID, "bh_load_state: read failed"); }
  uint64_t got[2];
  rc = bh::checksum_planes(s, h.plane_mask, h.first_strand, got);
  const int rc2 = bh::unmap_gl(s);
  if (rc) return rc;
  if (rc2) return rc2;
  if (got[0] != h.checksum[0] || got[1] != h.checksum[1])
    return fail(BH_ERR_INVALID, "bh_load_state: checksum mismatch (payload corrupted); the sim holds no valid state");
  s->params = h.params;
  s->initialized = had_state || (h.plane_mask & 1u) != 0;
  if (info) bh::fill_info(h, info);
  return BH_OK;
}

}  // extern "C"
