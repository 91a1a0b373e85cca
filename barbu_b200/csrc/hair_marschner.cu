// hair_marschner.cu — the Marschner lookup tables on the device.
//
// What Marschner::generate (src/fx/marschner.cc:35-69) dispatches: cs_marschner_m.glsl (longitudinal lobes + cos theta_d)
// and cs_marschner_n.glsl (azimuthal terms through the generic Np, `#if 1` branch) over a kTextureResolution^2 image of
// GL_RGBA16F texels, one invocation per texel in 16 x 16 groups (marschner.h:28-30). One kernel writes both tables.
// The operation order is the one oracle/barbu_marschner_oracle.c spells out (no contraction: explicit .rn products and
// sums); sinf/cosf/asinf/acosf/expf/powf are CUDA's, so results agree with the libm-based oracle to a few ulp except
// next to the solver's branch points (tests/test_marschner.py states the tolerance). 16,384 texels: a launch-latency
// kernel by construction, run once per parameter change (marschner.cc:27-32) — no roofline claim is made for it.
#include "hair_sim.cuh"

#include <cuda_fp16.h>

namespace bh {
namespace {

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float dvd(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float glm_min(float x, float y) { return (y < x) ? y : x; }

constexpr float kEps = 1e-6f;          // inc_constants.glsl:6
constexpr float kPi = 3.141564f;       // inc_constants.glsl:7 (the reference's value)

__device__ float gaussian(float sigma, float x_mu) {                         // inc_maths.glsl:270-272
  return dvd(expf(dvd(-mul(x_mu, x_mu), mul(mul(2.0f, sigma), sigma))), mul(2.5066282f, fabsf(sigma)));
}

__device__ float fresnel_power_ratio(float etaRatio, float nA, float nB, float cosI, float sinI) {   // inc_fresnel.glsl:14-24
  const float sinTSquared = powf(mul(etaRatio, sinI), 2.0f);
  if (sinTSquared > 1.0f) return 1.0f;
  const float cosT = __fsqrt_rn(sub(1.0f, sinTSquared));
  const float A = mul(nA, cosI), B = mul(nB, cosT);
  const float R = dvd(sub(A, B), add(A, B));
  return glm_min(1.0f, mul(R, R));
}
__device__ float fresnel(float etaOrigin, float etaPerp, float etaPar, float cosA, float sinA) {     // inc_fresnel.glsl:29-57
  const float r = fresnel_power_ratio(dvd(etaOrigin, etaPerp), etaOrigin, etaPerp, cosA, sinA);
  const float t = fresnel_power_ratio(dvd(etaOrigin, etaPar), etaPar, etaOrigin, cosA, sinA);
  return add(mul(r, sub(1.0f, 0.5f)), mul(t, 0.5f));
}

struct Roots { float v[3]; int n; };

__device__ Roots solver_linear(float a, float b) {                           // inc_solver.glsl:16-18
  Roots r = { { 0.f, 0.f, 0.f }, 0 };
  if (fabsf(a) > kEps) { r.v[0] = dvd(-b, a); r.n = 1; }
  return r;
}
__device__ Roots solver_quadratic(float a, float b, float c) {               // inc_solver.glsl:22-35
  Roots r = { { 0.f, 0.f, 0.f }, 0 };
  if (fabsf(a) < kEps) return solver_linear(b, c);
  float delta = sub(mul(b, b), mul(mul(4.0f, a), c));
  if (delta < 0.0f) return r;
  delta = __fsqrt_rn(delta);
  r.v[0] = dvd(add(-b, delta), mul(2.0f, a));
  r.v[1] = dvd(sub(-b, delta), mul(2.0f, a));
  r.n = (delta < kEps) ? 1 : 2;                                              // int(1 + step(Epsilon(), delta))
  return r;
}
__device__ float glm_sign(float x) { return (float)((0.0f < x) - (x < 0.0f)); }
__device__ Roots solver_cubic_normalized(float a, float b, float c) {        // inc_solver.glsl:51-90
  Roots roots = { { 0.f, 0.f, 0.f }, 0 };
  if (fabsf(c) < kEps) {
    roots = solver_quadratic(1.0f, a, b);
    roots.v[roots.n] = 0.0f;
    roots.n += 1;
  } else {
    const float Q = dvd(sub(mul(3.0f, b), mul(a, a)), 9.0f);
    const float R = dvd(sub(sub(mul(mul(9.0f, a), b), mul(27.0f, c)), mul(mul(mul(2.0f, a), a), a)), 54.0f);
    const float Q3 = mul(mul(Q, Q), Q);
    const float D = add(Q3, mul(R, R));
    const float third_a = dvd(a, 3.0f);
    if (D > 0.0f) {
      const float sqrtD = __fsqrt_rn(D);
      const float s = mul(glm_sign(add(R, sqrtD)), powf(fabsf(add(R, sqrtD)), 0.333f));
      const float t = mul(glm_sign(sub(R, sqrtD)), powf(fabsf(sub(R, sqrtD)), 0.333f));
      roots.v[0] = sub(add(s, t), third_a);
      roots.n = 1;
    } else {
      const float theta = acosf(mul(R, dvd(1.0f, __fsqrt_rn(-Q3))));
      const float twoSqrtQ = mul(2.0f, __fsqrt_rn(-Q));
      roots.v[0] = __fmaf_rn(twoSqrtQ, cosf(dvd(theta, 3.0f)), -third_a);
      roots.v[1] = __fmaf_rn(twoSqrtQ, cosf(dvd(add(theta, mul(2.0f, kPi)), 3.0f)), -third_a);
      roots.v[2] = __fmaf_rn(twoSqrtQ, cosf(dvd(add(theta, mul(4.0f, kPi)), 3.0f)), -third_a);
      roots.n = 3;
    }
  }
  return roots;
}
__device__ Roots solver_cubic(float a, float b, float c, float d) {          // inc_solver.glsl:42-47
  return (fabsf(a) < kEps) ? solver_quadratic(b, c, d) : solver_cubic_normalized(dvd(b, a), dvd(c, a), dvd(d, a));
}

__device__ float calculate_absorption(int p, float absorption, float etaPerp, float etaPar, float cosGammaI, float sinGammaI) {
  if (p == 0) return fresnel(1.0f, etaPerp, etaPar, cosGammaI, sinGammaI);   // inc_marschner_n.glsl:66-95
  const float sinGammaT = dvd(sinGammaI, etaPerp);
  const float gammaT = asinf(sinGammaT);
  const float cosGammaT = cosf(gammaT);
  const float fi = fresnel(1.0f, etaPerp, etaPar, cosGammaI, sinGammaI);
  const float ft = fresnel(1.0f, dvd(1.0f, etaPerp), dvd(1.0f, etaPar), cosGammaT, sinGammaT);
  const float t = expf(mul(mul(-4.0f, absorption), powf(cosGammaT, 2.0f)));
  return mul(mul(powf(sub(1.0f, fi), 2.0f), powf(ft, (float)(p - 1))), powf(t, (float)p));
}

__device__ float Np(int p, float absorption, float etaPerp, float etaPar, float c, float phi) {      // inc_marschner_n.glsl:100-131
  const float pc = mul((float)p, c);
  const float kx = mul(pc, -0.25801227547f), kz = sub(mul(pc, 1.90985931710f), 2.0f), kw = sub(mul((float)p, kPi), phi);
  const Roots roots = solver_cubic(kx, 0.0f, kz, kw);
  float L = 0.0f;
  for (int i = 0; i < roots.n; ++i) {
    const float gammaI = roots.v[i];
    const float sinGammaI = sinf(gammaI), cosGammaI = cosf(gammaI);
    const float a = calculate_absorption(p, absorption, etaPerp, etaPar, cosGammaI, sinGammaI);
    const float d = dvd(add(mul(mul(3.0f, kx), powf(gammaI, 2.0f)), kz), cosGammaI);                 // first_derivative_result
    L = add(L, mul(a, dvd(1.0f, fabsf(mul(2.0f, d)))));
  }
  return glm_min(L, 1.0f);
}

__device__ __forceinline__ void store_texel(float4 v, size_t at, uint2* half4, float4* full4) {
  if (full4) full4[at] = v;
  if (half4) {                                                               // the GL_RGBA16F store (marschner.h:30)
    const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
    half4[at] = make_uint2(*reinterpret_cast<const unsigned int*>(&lo), *reinterpret_cast<const unsigned int*>(&hi));
  }
}

__global__ void __launch_bounds__(256) marschner_luts_kernel(const bh_marschner_params p, const int res, uint2* __restrict__ m16,
                                                             uint2* __restrict__ n16, float4* __restrict__ m32, float4* __restrict__ n32) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= res || y >= res) return;
  const size_t at = (size_t)y * res + x;
  const float inv = dvd(1.0f, (float)res);
  const float u = sub(mul(mul(2.0f, (float)x), inv), 1.0f), v = sub(mul(mul(2.0f, (float)y), inv), 1.0f);
  {                                                                          // cs_marschner_m.glsl:35-62
    const float thetaI = asinf(u), thetaR = asinf(v);
    const float thetaH = dvd(add(thetaI, thetaR), 2.0f), thetaD = dvd(sub(thetaI, thetaR), 2.0f);
    const float degH = mul(thetaH, 57.295779513082320876798154814105f);
    const float m0 = gaussian(mul(1.0f, p.br), sub(degH, mul(1.0f, p.ar)));
    const float m1 = gaussian(mul(0.5f, p.br), sub(degH, mul(-0.5f, p.ar)));
    const float m2 = gaussian(mul(2.0f, p.br), sub(degH, mul(-1.5f, p.ar)));
    store_texel(make_float4(m0, m1, m2, cosf(thetaD)), at, m16, m32);
  }
  {                                                                          // cs_marschner_n.glsl:34-83
    const float cosPhiD = u, cosThetaD = v;
    const float sinThetaDSquared = sub(1.0f, powf(cosThetaD, 2.0f));
    const float refractionSquared = mul(p.eta, p.eta);
    const float etaPerp = dvd(__fsqrt_rn(sub(refractionSquared, sinThetaDSquared)), cosThetaD);
    const float etaPar = dvd(refractionSquared, etaPerp);
    const float phiD = acosf(cosPhiD);
    const float c = asinf(dvd(1.0f, etaPerp));
    const float R = Np(0, p.absorption, etaPerp, etaPar, c, phiD);
    const float TT = Np(1, p.absorption, etaPerp, etaPar, c, phiD);
    const float TRT = Np(2, p.absorption, etaPerp, etaPar, c, phiD);
    store_texel(make_float4(R, TT, TRT, 1.0f), at, n16, n32);
  }
}

}  // namespace
}  // namespace bh

using bh::fail;

extern "C" {

void bh_marschner_default_params(bh_marschner_params* p) {                   // marschner.h:38-52
  if (!p) return;
  p->eta = 1.55f; p->absorption = 0.20f; p->eccentricity = 0.85f;
  p->ar = -5.0f; p->br = 5.0f;
  p->glint_scale = 0.5f; p->azimuthal_width = 10.0f; p->delta_caustic = 0.2f; p->delta_hm = 0.5f;
}

int bh_marschner_generate(const bh_marschner_params* p, int resolution, int device, uint16_t* m_rgba16f, uint16_t* n_rgba16f,
                          float* m_rgba32f, float* n_rgba32f) {
  if (!p || resolution < 1 || resolution > 8192) return fail(BH_ERR_INVALID, "bh_marschner_generate: bad argument");
  int ndev = 0;
  BH_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(BH_ERR_INVALID, "bh_marschner_generate: no such CUDA device");
  bh::DeviceGuard g(device);
  const size_t texels = (size_t)resolution * resolution;
  uint2 *d16 = nullptr; float4* d32 = nullptr;
  cudaError_t e = cudaMalloc(&d16, 2 * texels * sizeof(uint2));
  if (e == cudaSuccess) e = cudaMalloc(&d32, 2 * texels * sizeof(float4));
  if (e == cudaSuccess) {
    const dim3 block(16, 16), grid((resolution + 15) / 16, (resolution + 15) / 16);   // kComputeBlockSize, marschner.h:29
    bh::marschner_luts_kernel<<<grid, block>>>(*p, resolution, d16, d16 + texels, d32, d32 + texels);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess && m_rgba16f) e = cudaMemcpy(m_rgba16f, d16, texels * sizeof(uint2), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && n_rgba16f) e = cudaMemcpy(n_rgba16f, d16 + texels, texels * sizeof(uint2), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && m_rgba32f) e = cudaMemcpy(m_rgba32f, d32, texels * sizeof(float4), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && n_rgba32f) e = cudaMemcpy(n_rgba32f, d32 + texels, texels * sizeof(float4), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaFree(d16); cudaFree(d32);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(BH_ERR_CUDA, "bh_marschner_generate", e); }
  return BH_OK;
}

}  // extern "C"
