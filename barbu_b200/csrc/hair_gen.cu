// hair_gen.cu — device-side strand generation, patch-index generation and root skinning.
//
// These are the one-off / per-frame producers either side of the step kernel:
//   expand_strands   Hair::init_simulation positions + velocities     (src/fx/hair.cc:255-287)
//   sphere_roots     synthetic pole-free lat-long scalp               (no reference code)
//   patch_indices    Hair::init_mesh element buffer                   (src/fx/hair.cc:397-409)
//   skin_roots_dq    extension: dual-quaternion skinned roots         (src/shaders/shared/inc_skinning.glsl:22-31,54-82)
// All fp32 arithmetic uses explicit round-to-nearest intrinsics in the reference's operation order
// (the reference host code is built -O2 -msse4.1: no FMA, nothing contracted), so results are
// bit-identical to the CPU oracle.
#include "hair_gen.cuh"

namespace bh {

namespace {

__global__ void __launch_bounds__(256) expand_strands_kernel(const float* __restrict__ root_pos3,
                                                             const float* __restrict__ root_nrm3,
                                                             const float* __restrict__ random_value,
                                                             long long nstrands, int N, float scaleOffset,
                                                             float4* __restrict__ pos, float4* __restrict__ vel) {
  const long long V = nstrands * N;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < V; idx += (long long)gridDim.x * blockDim.x) {
    const long long j = idx / N;
    const int i = (int)(idx - j * N);
    const float rv = random_value[j];
    // offset = float(i) * scaleOffset * random_value   (hair.cc:281-283); lastOffset = previous i's offset
    const float offset = __fmul_rn(__fmul_rn((float)i, scaleOffset), rv);
    const float last = (i > 0) ? __fmul_rn(__fmul_rn((float)(i - 1), scaleOffset), rv) : 0.0f;
    const float vx = root_pos3[3 * j], vy = root_pos3[3 * j + 1], vz = root_pos3[3 * j + 2];
    const float nx = root_nrm3[3 * j], ny = root_nrm3[3 * j + 1], nz = root_nrm3[3 * j + 2];
    // Positions[idx] = vec4(v + offset * n, offset - lastOffset)   (hair.cc:284)
    pos[idx] = make_float4(__fadd_rn(vx, __fmul_rn(offset, nx)), __fadd_rn(vy, __fmul_rn(offset, ny)),
                           __fadd_rn(vz, __fmul_rn(offset, nz)), __fsub_rn(offset, last));
    vel[idx] = make_float4(0.f, 0.f, 0.f, 0.f);                   // hair.cc:285
  }
}

// rowtab[2r] = cos(theta_r), rowtab[2r+1] = sin(theta_r); coltab likewise for phi_c; both rounded
// from libm double on the host, so the only device arithmetic is two exact-rounded products.
__global__ void __launch_bounds__(256) sphere_roots_kernel(const float* __restrict__ rowtab, const float* __restrict__ coltab,
                                                           int rows, int cols, int column_major, long long first, long long count,
                                                           float* __restrict__ root_pos3, float* __restrict__ root_nrm3) {
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += (long long)gridDim.x * blockDim.x) {
    const long long g = first + j;
    // strand order: row-major g = r * cols + c (a latitude circle after the other), or column-major g = c * rows + r
    // (a meridian after the other: a contiguous strand range is then a longitude wedge that holds every latitude)
    const long long r = column_major ? g % rows : g / cols;
    const int c = (int)(column_major ? g / rows : g - r * cols);
    const float ct = rowtab[2 * r], st = rowtab[2 * r + 1];
    const float cp = coltab[2 * c], sp = coltab[2 * c + 1];
    const float nx = __fmul_rn(ct, cp), ny = st, nz = __fmul_rn(ct, sp);
    root_nrm3[3 * j] = nx; root_nrm3[3 * j + 1] = ny; root_nrm3[3 * j + 2] = nz;
    root_pos3[3 * j] = nx; root_pos3[3 * j + 1] = ny; root_pos3[3 * j + 2] = nz;
  }
}

// One thread per (face, segment): 6 ints  e_k, e_k + 1 with e_k = N * indices[3f + k] + j.
__global__ void __launch_bounds__(256) patch_indices_kernel(const int* __restrict__ tri, long long nfaces, int N,
                                                            int* __restrict__ out) {
  const int nseg = N - 1;
  const long long total = nfaces * nseg;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const long long f = q / nseg;
    const int j = (int)(q - f * nseg);
    int e[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) e[k] = N * tri[3 * f + k] + j;
    int2* o = reinterpret_cast<int2*>(out + 6 * q);
    o[0] = make_int2(e[0], e[0] + 1);
    o[1] = make_int2(e[1], e[1] + 1);
    o[2] = make_int2(e[2], e[2] + 1);
  }
}

__device__ __forceinline__ void cross3(const float a[3], const float b[3], float o[3]) {
  o[0] = __fsub_rn(__fmul_rn(a[1], b[2]), __fmul_rn(b[1], a[2]));
  o[1] = __fsub_rn(__fmul_rn(a[2], b[0]), __fmul_rn(b[2], a[0]));
  o[2] = __fsub_rn(__fmul_rn(a[0], b[1]), __fmul_rn(b[0], a[1]));
}
// vec4 * mat3x4, one column: left to right (glm/detail/type_mat3x4.inl:454-464)
__device__ __forceinline__ float dot4(const float* a, const float* b) {
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a[0], b[0]), __fmul_rn(a[1], b[1])), __fmul_rn(a[2], b[2])), __fmul_rn(a[3], b[3]));
}

// Root of strand s -> position plane vertex s*N (rest length in .w preserved).
__global__ void __launch_bounds__(256) skin_roots_dq_kernel(const float* __restrict__ rest_pos3, const int* __restrict__ joints4,
                                                            const float* __restrict__ weights3, const float* __restrict__ dq,
                                                            long long nstrands, int N, float4* __restrict__ pos) {
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < nstrands; s += (long long)gridDim.x * blockDim.x) {
    float v[3] = { rest_pos3[3 * s], rest_pos3[3 * s + 1], rest_pos3[3 * s + 2] };
    float w[4] = { weights3[3 * s], weights3[3 * s + 1], weights3[3 * s + 2], 0.f };
    if (!(w[0] <= 1e-6f)) {                                      // apply_skinning early-out, l.23-25
      w[3] = __fsub_rn(1.0f, __fadd_rn(__fadd_rn(w[0], w[1]), w[2]));
      const float* q[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) q[k] = dq + 8 * (long long)joints4[4 * s + k];
#pragma unroll
      for (int k = 0; k < 3; ++k) {                              // weights.xyz *= sign(Ma[3] * mat3x4(Ma)), l.65
        const float d = dot4(q[3], q[k]);
        const float sg = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
        w[k] = __fmul_rn(w[k], sg);
      }
      float A[4], B[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {                              // A = Ma * w, B = Mb * w (GLM mat4*vec4 order)
        A[c] = __fadd_rn(__fadd_rn(__fmul_rn(q[0][c], w[0]), __fmul_rn(q[1][c], w[1])),
                         __fadd_rn(__fmul_rn(q[2][c], w[2]), __fmul_rn(q[3][c], w[3])));
        B[c] = __fadd_rn(__fadd_rn(__fmul_rn(q[0][4 + c], w[0]), __fmul_rn(q[1][4 + c], w[1])),
                         __fadd_rn(__fmul_rn(q[2][4 + c], w[2]), __fmul_rn(q[3][4 + c], w[3])));
      }
      // glm::dot(vec4, vec4) pairs the products: (x*x + y*y) + (z*z + w*w)   (glm/detail/func_geometric.inl:58-65)
      const float inv = __frcp_rn(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(A[0], A[0]), __fmul_rn(A[1], A[1])),
                                                       __fadd_rn(__fmul_rn(A[2], A[2]), __fmul_rn(A[3], A[3])))));
#pragma unroll
      for (int c = 0; c < 4; ++c) { A[c] = __fmul_rn(A[c], inv); B[c] = __fmul_rn(B[c], inv); }
      float c1[3], c2[3], cab[3];
      cross3(A, v, c1);                                          // v += 2 * cross(A.xyz, cross(A.xyz, v) + A.w*v)
#pragma unroll
      for (int c = 0; c < 3; ++c) c1[c] = __fadd_rn(c1[c], __fmul_rn(A[3], v[c]));
      cross3(A, c1, c2);
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = __fadd_rn(v[c], __fmul_rn(2.0f, c2[c]));
      cross3(A, B, cab);                                         // v += 2 * (A.w*B.xyz - B.w*A.xyz + cross(A.xyz, B.xyz))
#pragma unroll
      for (int c = 0; c < 3; ++c)
        v[c] = __fadd_rn(v[c], __fmul_rn(2.0f, __fadd_rn(__fsub_rn(__fmul_rn(A[3], B[c]), __fmul_rn(B[3], A[c])), cab[c])));
    }
    float4 p = pos[s * N];
    p.x = v[0]; p.y = v[1]; p.z = v[2];
    pos[s * N] = p;
  }
}

inline unsigned grid_for(long long n, int block) {
  long long b = (n + block - 1) / block;
  const long long cap = 148LL * 32;   // grid-stride loops: a few waves of the 148 SMs is enough
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace

cudaError_t launch_expand_strands(const float* root_pos3, const float* root_nrm3, const float* random_value,
                                  long long nstrands, int nverts, float scaleOffset, float4* pos, float4* vel,
                                  cudaStream_t stream) {
  if (nstrands <= 0) return cudaSuccess;
  expand_strands_kernel<<<grid_for(nstrands * nverts, 256), 256, 0, stream>>>(root_pos3, root_nrm3, random_value, nstrands,
                                                                              nverts, scaleOffset, pos, vel);
  return cudaGetLastError();
}

cudaError_t launch_sphere_roots(const float* rowtab, const float* coltab, int rows, int cols, int column_major, long long first,
                                long long count, float* root_pos3, float* root_nrm3, cudaStream_t stream) {
  if (count <= 0) return cudaSuccess;
  sphere_roots_kernel<<<grid_for(count, 256), 256, 0, stream>>>(rowtab, coltab, rows, cols, column_major, first, count, root_pos3, root_nrm3);
  return cudaGetLastError();
}

cudaError_t launch_patch_indices(const int* tri, long long nfaces, int nverts, int* out, cudaStream_t stream) {
  if (nfaces <= 0 || nverts < 2) return cudaSuccess;
  patch_indices_kernel<<<grid_for(nfaces * (nverts - 1), 256), 256, 0, stream>>>(tri, nfaces, nverts, out);
  return cudaGetLastError();
}

cudaError_t launch_skin_roots_dq(const float* rest_pos3, const int* joints4, const float* weights3, const float* dq,
                                 long long nstrands, int nverts, float4* pos, cudaStream_t stream) {
  if (nstrands <= 0) return cudaSuccess;
  skin_roots_dq_kernel<<<grid_for(nstrands, 256), 256, 0, stream>>>(rest_pos3, joints4, weights3, dq, nstrands, nverts, pos);
  return cudaGetLastError();
}

}  // namespace bh
