// hair_host.cc — the host-evaluated parts of Hair::init_simulation (src/fx/hair.cc:236-361).
//
// The reference generates strand state on the host. Positions/velocities moved to the device
// (hair_gen.cu); what stays here is what is defined by host libraries and therefore cannot be
// reproduced bit-for-bit by device code: the glibc rand() length jitter (hair.cc:273-275) and the
// tangent plane (hair.cc:290-328), which goes through libm sinf/cosf and glm::simplex.
// Build with -ffp-contract=off: one rounding per written operation, like the reference's
// -O2 -msse4.1 build (CMakeLists.txt:263-264).
#include "../../include/barbu_hair.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

namespace {

// --- 2-D simplex noise as computed by glm::simplex(vec2) (GLM 0.9.9.9, gtc/noise.inl:591-645) ----
// Ashima/McEwan "textureless" simplex noise; operation order follows GLM's vector expressions.
inline float mod289(float x) { return x - std::floor(x * (1.0f / 289.0f)) * 289.0f; }      // detail/_noise.hpp:9-12
inline float permute(float x) { return mod289(((x * 34.0f) + 1.0f) * x); }                  // detail/_noise.hpp:15-18

float simplex2(float vx, float vy) {
  // GLM writes T(<double literal>): the decimal is rounded to double first, then to float
  const float Cx = static_cast<float>(0.211324865405187), Cy = static_cast<float>(0.366025403784439);
  const float Cz = static_cast<float>(-0.577350269189626), Cw = static_cast<float>(0.024390243902439);
  const float skew = vx * Cy + vy * Cy;                                   // dot(v, C.yy)
  float ix = std::floor(vx + skew), iy = std::floor(vy + skew);
  const float unskew = ix * Cx + iy * Cx;                                 // dot(i, C.xx)
  const float x0x = vx - ix + unskew, x0y = vy - iy + unskew;
  const bool lower = x0x > x0y;
  const float i1x = lower ? 1.0f : 0.0f, i1y = lower ? 0.0f : 1.0f;
  const float x1x = (x0x + Cx) - i1x, x1y = (x0y + Cx) - i1y;             // x12.xy
  const float x2x = x0x + Cz, x2y = x0y + Cz;                             // x12.zw
  ix = ix - 289.0f * std::floor(ix / 289.0f);                             // mod(i, 289)
  iy = iy - 289.0f * std::floor(iy / 289.0f);
  const float py[3] = { iy + 0.0f, iy + i1y, iy + 1.0f };
  const float px[3] = { 0.0f, i1x, 1.0f };
  const float cornerx[3] = { x0x, x1x, x2x }, cornery[3] = { x0y, x1y, x2y };
  float m[3], g[3];
  for (int c = 0; c < 3; ++c) {
    const float p = permute(permute(py[c]) + ix + px[c]);
    float mc = 0.5f - (cornerx[c] * cornerx[c] + cornery[c] * cornery[c]);
    mc = mc < 0.0f ? 0.0f : mc;
    mc = mc * mc;
    mc = mc * mc;
    const float pc = p * Cw;
    const float x = 2.0f * (pc - std::floor(pc)) - 1.0f;                  // 2 * fract(p * C.w) - 1
    const float h = std::fabs(x) - 0.5f;
    const float a0 = x - std::floor(x + 0.5f);
    m[c] = mc * (static_cast<float>(1.79284291400159) - static_cast<float>(0.85373472095314) * (a0 * a0 + h * h));
    g[c] = a0 * cornerx[c] + h * cornery[c];
  }
  return 130.0f * ((m[0] * g[0] + m[1] * g[1]) + m[2] * g[2]);
}

}  // namespace

// errors of the host-only entry points reach bh_last_error() like those of the CUDA ones (hair_capi.cu)
extern "C" int bh_host_fail_impl(int code, const char* what);
static int bh_host_fail(int code, const char* what) { return bh_host_fail_impl(code, what); }

extern "C" {

// glibc's rand() is random() on the TYPE_3 additive-feedback generator (128 bytes of state). The reentrant form of the
// same generator (initstate_r / random_r, stdlib.h) yields the same sequence from a PRIVATE state: the host application's
// own rand() stream — the reference seeds it once with time() and its other modules draw from it (core/app.cc:96-97) — is
// neither reseeded nor consumed, and concurrent callers do not race. The state is kept per thread together with its
// position in the sequence, so shard after shard of one scalp (ascending `first`, same seed) costs O(count), not O(first).
int bh_random_values(unsigned seed, int64_t first, int64_t count, float* out) {
  if (first < 0 || count < 0 || (count > 0 && !out)) return bh_host_fail(BH_ERR_INVALID, "bh_random_values: bad argument");
  struct Gen { struct random_data rd; char state[128]; unsigned seed; int64_t pos; bool live; };
  static thread_local Gen g = {};
  if (!g.live || g.seed != seed || g.pos > first) {
    std::memset(&g.rd, 0, sizeof g.rd);
    if (initstate_r(seed, g.state, sizeof g.state, &g.rd) != 0) return bh_host_fail(BH_ERR_INVALID, "bh_random_values: initstate_r failed");
    g.seed = seed; g.pos = 0; g.live = true;
  }
  int32_t r = 0;
  for (; g.pos < first; ++g.pos) (void)random_r(&g.rd, &r);
  for (int64_t j = 0; j < count; ++j, ++g.pos) {
    (void)random_r(&g.rd, &r);                                            // == rand() after srand(seed), hair.cc:273-275
    out[j] = static_cast<float>(1.0 + 0.1 * (1.0 - 2.0 * static_cast<double>(r) / static_cast<double>(RAND_MAX)));
  }
  return BH_OK;
}

int bh_init_tangents_host(const float* root_nrm3, int64_t total, int64_t first, int64_t count, int nverts,
                          float maxlength, float* tan4) {
  if (!root_nrm3 || !tan4 || total <= 0 || first < 0 || count < 0 || first + count > total || nverts < 1) return bh_host_fail(BH_ERR_INVALID, "bh_init_tangents_host: bad argument");
  const int N = nverts;
  const float inv_nroots = 1.0f / static_cast<float>(total);              // hair.cc:292
  const float kPi = static_cast<float>(3.14159265358979323846264338327950288);             // glm::pi<float>()
  const float scaleMaxLength = 0.125f * sqrtf(maxlength);                 // hair.cc:294
  for (int64_t q = 0; q < count; ++q) {
    const int64_t j = first + q;                                          // global strand index
    const float* nr = root_nrm3 + 3 * q;
    float* T = tan4 + 4 * (size_t)q * N;
    const float dj = static_cast<float>(j + 1) * inv_nroots;              // hair.cc:302
    const float n0 = 1.25f * simplex2(sinf(3.0f * dj), cosf(5.0f));       // hair.cc:305
    float curly[3] = { cosf(n0 * 4.0f * kPi), -0.71f * n0, sinf(n0 * 2.7f * kPi) };
    const int B = N - 1;
    for (int c = 0; c < 3; ++c) T[c] = .15f * nr[c];                      // outer tangents, hair.cc:310-311
    T[3] = .15f * 0.0f;
    for (int c = 0; c < 3; ++c) T[4 * B + c] = .2f * (-nr[c] + curly[c]);
    T[4 * B + 3] = .2f * 0.0f;
    const float dist_AB = static_cast<float>(B);
    const float inv_dist = 1.0f / dist_AB;
    for (int i = 1; i < B; ++i) {                                         // inner tangents, hair.cc:316-326
      const float di = 10.0f * static_cast<float>(B - i) / (dist_AB - 1.0f);
      const float n = di * simplex2(sinf(43.0f * dj), cosf(5.0f * di));
      curly[0] = -5.8f * (10.7f * cosf(n * kPi));
      curly[1] = -5.8f * (-2.3f * n);
      curly[2] = -5.8f * (20.5f * sinf(n * kPi));
      const float s = 0.1f * static_cast<float>(i) * inv_dist * scaleMaxLength;
      for (int c = 0; c < 3; ++c) T[4 * i + c] = s * curly[c];
      T[4 * i + 3] = s * 0.0f;
    }
  }
  return BH_OK;
}

int bh_sphere_scalp_triangles(int rows, int cols, int32_t* tri) { return bh_sphere_scalp_triangles_ordered(rows, cols, BH_SCALP_ROW_MAJOR, tri); }

int bh_sphere_scalp_triangles_ordered(int rows, int cols, int order, int32_t* tri) {
  if (rows < 1 || cols < 1 || !tri || (order != BH_SCALP_ROW_MAJOR && order != BH_SCALP_COLUMN_MAJOR)) return bh_host_fail(BH_ERR_INVALID, "bh_sphere_scalp_triangles: bad argument");
  if ((int64_t)rows * cols > INT32_MAX) return bh_host_fail(BH_ERR_OVERFLOW, "bh_sphere_scalp_triangles: vertex ids exceed int32");
  const bool cm = order == BH_SCALP_COLUMN_MAJOR;
  auto id = [&](int r, int c) -> int32_t { return cm ? c * rows + r : r * cols + c; };   // the same faces in the same order; only the vertex numbering differs
  size_t t = 0;
  for (int r = 0; r + 1 < rows; ++r)
    for (int c = 0; c < cols; ++c) {
      const int c1 = (c + 1) % cols;
      const int32_t v00 = id(r, c), v10 = id(r + 1, c), v01 = id(r, c1), v11 = id(r + 1, c1);
      tri[t++] = v00; tri[t++] = v10; tri[t++] = v01;
      tri[t++] = v01; tri[t++] = v10; tri[t++] = v11;
    }
  return BH_OK;
}

// ---- scalp input: Wavefront OBJ as the reference's mesh loader reads it -------------------------------------------------
// ParseOBJ (src/memory/resources/mesh_data_manager.cc:69-223): one pass over '\n'-terminated lines (a last line without
// newline is ignored); `v x y z`, `vn x y z`, `vt u v` (v flipped, unused here); `f` with v/vt/vn corner triples, a 4th
// corner splits the quad into (x, y, z), (z, ext, x) (l.201-208); indices are 1-based in the file.
// MeshData::setup (src/memory/resources/mesh_data.cc:384-406): vertices are re-indexed in first-appearance order of the
// unique (v, vt, vn) triples over the corner list; vertex j takes position[v] and normal[vn]. One root per vertex
// (src/fx/hair.cc:58). A scalp without normals gets per-corner normals as recalculateNormals makes them
// (src/utils/raw_mesh_file.cc:11-50) and so 3 strands per face.
int bh_load_obj_scalp(const char* path, float** pos3, float** nrm3, int64_t* nvertices, int32_t** tri, int64_t* nfaces) {
  if (!path || !pos3 || !nrm3 || !nvertices || !tri || !nfaces) return bh_host_fail(BH_ERR_INVALID, "bh_load_obj_scalp: NULL argument");
  *pos3 = *nrm3 = nullptr; *tri = nullptr; *nvertices = *nfaces = 0;
  FILE* f = std::fopen(path, "rb");
  if (!f) return bh_host_fail(BH_ERR_INVALID, "bh_load_obj_scalp: cannot open the file");                                           // "The scalp mesh resource was not found." (hair.cc:45-48)
  std::string text;
  char buf[1 << 16];
  for (size_t n; (n = std::fread(buf, 1, sizeof buf, f)) > 0;) text.append(buf, n);
  std::fclose(f);

  struct V3 { float x, y, z; };
  struct I3 { int v, t, n; };
  std::vector<V3> positions, normals;
  size_t ntex = 0;
  std::vector<I3> corners;
  size_t start = 0;
  for (size_t end; (end = text.find('\n', start)) != std::string::npos; start = end + 1) {
    text[end] = '\0';
    const char* s = text.c_str() + start;
    if (s[0] & 1) continue;                                                // '#', 'o', 'g', 's', 'u'semtl, 'm'tllib: first char odd
    if (s[0] == 'v') {
      V3 v{ 0.f, 0.f, 0.f };
      if (s[1] == ' ') { std::sscanf(s + 2, "%f %f %f", &v.x, &v.y, &v.z); positions.push_back(v); }
      else if (s[1] == 't') { ++ntex; }
      else { std::sscanf(s + 3, "%f %f %f", &v.x, &v.y, &v.z); normals.push_back(v); }
    } else if (s[0] == 'f') {
      I3 a{ 0, 0, 0 }, b{ 0, 0, 0 }, c{ 0, 0, 0 }, e{ 0, 0, 0 };
      const bool has_t = ntex != 0, has_n = !normals.empty();
      if (has_t && has_n) std::sscanf(s + 2, "%d/%d/%d %d/%d/%d %d/%d/%d %d/%d/%d", &a.v, &a.t, &a.n, &b.v, &b.t, &b.n, &c.v, &c.t, &c.n, &e.v, &e.t, &e.n);
      else if (has_t) std::sscanf(s + 2, "%d/%d %d/%d %d/%d %d/%d", &a.v, &a.t, &b.v, &b.t, &c.v, &c.t, &e.v, &e.t);
      else if (has_n) std::sscanf(s + 2, "%d//%d %d//%d %d//%d %d//%d", &a.v, &a.n, &b.v, &b.n, &c.v, &c.n, &e.v, &e.n);
      else std::sscanf(s + 2, "%d %d %d %d", &a.v, &b.v, &c.v, &e.v);
      corners.push_back(a); corners.push_back(b); corners.push_back(c);
      if (e.v > 0) { corners.push_back(c); corners.push_back(e); corners.push_back(a); }
    }
  }
  if (corners.empty() || positions.empty()) return bh_host_fail(BH_ERR_INVALID, "bh_load_obj_scalp: no faces or no positions in the file");
  if (normals.empty()) {
    // A scalp without `vn` lines: MeshData::setup recalculates (mesh_data.cc:366-372 -> RawMeshData::recalculateNormals,
    // src/utils/raw_mesh_file.cc:11-50). Per face the unit normal of (v2 - v1) x (v3 - v2) is added, unweighted, to its three
    // vertices; the sums are normalised; then every CORNER receives a normal entry of its own (a copy of its vertex's), so
    // the re-indexing below sees 3 * F unique triples — one root per face corner. GLM arithmetic: cross as written in
    // func_geometric.inl:74-77, dot = (x*x + y*y) + z*z, normalize = v * (1 / sqrt(dot)).
    for (const I3& c : corners)
      if (c.v < 1 || (size_t)c.v > positions.size()) return bh_host_fail(BH_ERR_INVALID, "bh_load_obj_scalp: face index outside the v list");
    std::vector<V3> sum(positions.size(), V3{ 0.f, 0.f, 0.f });
    auto unit = [](V3 a) { const float inv = 1.0f / std::sqrt((a.x * a.x + a.y * a.y) + a.z * a.z); return V3{ a.x * inv, a.y * inv, a.z * inv }; };
    for (size_t f3 = 0; f3 + 2 < corners.size(); f3 += 3) {
      const V3 &p1 = positions[corners[f3].v - 1], &p2 = positions[corners[f3 + 1].v - 1], &p3 = positions[corners[f3 + 2].v - 1];
      const V3 u{ p2.x - p1.x, p2.y - p1.y, p2.z - p1.z }, w{ p3.x - p2.x, p3.y - p2.y, p3.z - p2.z };
      const V3 n = unit(V3{ u.y * w.z - w.y * u.z, u.z * w.x - w.z * u.x, u.x * w.y - w.x * u.y });
      for (int k = 0; k < 3; ++k) { V3& a = sum[corners[f3 + k].v - 1]; a.x += n.x; a.y += n.y; a.z += n.z; }
    }
    for (V3& a : sum) a = unit(a);
    normals.reserve(corners.size());
    for (I3& c : corners) { normals.push_back(sum[c.v - 1]); c.n = (int)normals.size(); }   // 1-based like the file's own indices
  }
  std::map<std::tuple<int, int, int>, int32_t> seen;
  std::vector<I3> unique;
  std::vector<int32_t> indices;
  indices.reserve(corners.size());
  for (const I3& c0 : corners) {
    const I3 c{ c0.v - 1, c0.t - 1, c0.n - 1 };                            // [1, n] -> [0, n-1]; absent attributes become -1
    if (c.v < 0 || (size_t)c.v >= positions.size() || c.n < 0 || (size_t)c.n >= normals.size()) return bh_host_fail(BH_ERR_INVALID, "bh_load_obj_scalp: face index outside the v / vn lists");
    auto key = std::make_tuple(c.v, c.t, c.n);
    auto it = seen.find(key);
    if (it == seen.end()) { it = seen.emplace(key, (int32_t)unique.size()).first; unique.push_back(c); }
    indices.push_back(it->second);
  }
  const size_t nv = unique.size(), nf = indices.size() / 3;
  float* P = static_cast<float*>(std::malloc(sizeof(float) * 3 * nv));
  float* Nn = static_cast<float*>(std::malloc(sizeof(float) * 3 * nv));
  int32_t* T = static_cast<int32_t*>(std::malloc(sizeof(int32_t) * 3 * nf));
  if (!P || !Nn || !T) { std::free(P); std::free(Nn); std::free(T); return bh_host_fail(BH_ERR_INVALID, "bh_load_obj_scalp: out of host memory"); }
  for (size_t j = 0; j < nv; ++j) {
    const V3 &p = positions[unique[j].v], &n = normals[unique[j].n];
    P[3 * j] = p.x; P[3 * j + 1] = p.y; P[3 * j + 2] = p.z;
    Nn[3 * j] = n.x; Nn[3 * j + 1] = n.y; Nn[3 * j + 2] = n.z;
  }
  std::memcpy(T, indices.data(), sizeof(int32_t) * 3 * nf);
  *pos3 = P; *nrm3 = Nn; *tri = T; *nvertices = (int64_t)nv; *nfaces = (int64_t)nf;
  return BH_OK;
}

// ---- skinning palette from joint matrices --------------------------------------------------------------------------------
// What SkeletonController::generate_skinning_datas (src/fx/animation/skeleton_controller.cc:248-265) hands the skinning
// shader per joint, from the two matrices it has: S = global_pose * inverse_bind; the 3 x 4 skinning matrix is the top three
// ROWS of S (l.255: first three columns of the transpose); glm::dualquat(mat3x4) (gtx/dual_quaternion.inl:303-351) takes the
// rotation by the usual largest-diagonal case split and the dual part as 0.5 * translation * rotation. Matrices arrive in
// GLM's layout (16 floats, column-major: element (row r, column c) at [4 * c + r]).
int bh_dq_palette_from_matrices(const float* global_pose16, const float* inverse_bind16, int njoints, float* dq_palette) {
  if (!global_pose16 || !inverse_bind16 || !dq_palette || njoints < 0) return bh_host_fail(BH_ERR_INVALID, "bh_dq_palette_from_matrices: bad argument");
  for (int j = 0; j < njoints; ++j) {
    const float* G = global_pose16 + 16 * (size_t)j;
    const float* I = inverse_bind16 + 16 * (size_t)j;
    float row[3][4];                                                         // row[r][c] = S(r, c), r < 3; the product in GLM's order of operations
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c)
        row[r][c] = ((G[r] * I[4 * c] + G[4 + r] * I[4 * c + 1]) + G[8 + r] * I[4 * c + 2]) + G[12 + r] * I[4 * c + 3];
    const float m00 = row[0][0], m11 = row[1][1], m22 = row[2][2];
    float qx, qy, qz, qw;
    const float trace = m00 + m11 + m22;
    if (trace > 0.0f) {
      const float r = std::sqrt(1.0f + trace), k = 0.5f / r;
      qw = 0.5f * r; qx = (row[2][1] - row[1][2]) * k; qy = (row[0][2] - row[2][0]) * k; qz = (row[1][0] - row[0][1]) * k;
    } else if (m00 > m11 && m00 > m22) {
      const float r = std::sqrt(1.0f + m00 - m11 - m22), k = 0.5f / r;
      qx = 0.5f * r; qy = (row[1][0] + row[0][1]) * k; qz = (row[0][2] + row[2][0]) * k; qw = (row[2][1] - row[1][2]) * k;
    } else if (m11 > m22) {
      const float r = std::sqrt(1.0f + m11 - m00 - m22), k = 0.5f / r;
      qx = (row[1][0] + row[0][1]) * k; qy = 0.5f * r; qz = (row[2][1] + row[1][2]) * k; qw = (row[0][2] - row[2][0]) * k;
    } else {
      const float r = std::sqrt(1.0f + m22 - m00 - m11), k = 0.5f / r;
      qx = (row[0][2] + row[2][0]) * k; qy = (row[2][1] + row[1][2]) * k; qz = 0.5f * r; qw = (row[1][0] - row[0][1]) * k;
    }
    const float tx = row[0][3], ty = row[1][3], tz = row[2][3];
    float* o = dq_palette + 8 * (size_t)j;
    o[0] = qx; o[1] = qy; o[2] = qz; o[3] = qw;
    o[4] = 0.5f * (tx * qw + ty * qz - tz * qy);
    o[5] = 0.5f * (-tx * qz + ty * qw + tz * qx);
    o[6] = 0.5f * (tx * qy - ty * qx + tz * qw);
    o[7] = -0.5f * (tx * qx + ty * qy + tz * qz);
  }
  return BH_OK;
}

void bh_free(void* ptr) { std::free(ptr); }

}  // extern "C"
