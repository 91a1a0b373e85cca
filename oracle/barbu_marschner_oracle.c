/*
 * barbu_marschner_oracle.c — CPU ORACLE for the Marschner lookup tables.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE (see barbu_hair_oracle.h): only tests/ may load it.
 *
 * Scalar C restatement, in a fully specified fp32 operation order, of what Marschner::generate
 * (src/fx/marschner.cc:35-69) dispatches:
 *   src/shaders/hair/marschner/cs_marschner_m.glsl:33-63      longitudinal lobes  M_R, M_TT, M_TRT + cos(theta_d)
 *   src/shaders/hair/marschner/cs_marschner_n.glsl:24-84      azimuthal terms     N_R, N_TT, N_TRT (generic Np branch, `#if 1`)
 *   src/shaders/hair/marschner/inc_marschner_n.glsl:32-126    polynomial coefficients, absorption, Np
 *   src/shaders/shared/inc_solver.glsl:16-95                  linear / quadratic / cubic solvers
 *   src/shaders/shared/inc_fresnel.glsl:14-58                 Fresnel terms
 *   src/shaders/shared/inc_maths.glsl:270-276                 gaussian
 *   src/shaders/shared/inc_constants.glsl:6-8                 Epsilon() = 1e-6, Pi() = 3.141564 (sic)
 * Built-ins follow the reference's vendored GLM (min(x,y) = y<x ? y : x, mix = x*(1-a) + y*a, degrees = x*57.29577951...,
 * inversesqrt = 1/sqrt, sign = (0<x)-(x<0), dot(vec4) = (x+y)+(z+w)), transcendental functions are libm's. GLSL's implicit
 * int -> float conversions are written out. pow(x, y) with x < 0 is undefined in GLSL; this file takes libm's powf, as GLM does.
 * Parity pin: oracle/_ref/libbarbu_ref_marschner.so (the shader SOURCES compiled as C++ against the reference's GLM by
 * oracle/Makefile) is bit-identical to this file — tests/test_marschner.py, tests/golden/marschner.npz.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "barbu_hair_oracle.h"

static const float kEps = 1e-6f;          /* inc_constants.glsl:6 */
static const float kPi = 3.141564f;       /* inc_constants.glsl:7 — the reference's value, not pi */

typedef struct { float x, y, z, w; } v4;

static float glm_min(float x, float y) { return (y < x) ? y : x; }

/* inc_maths.glsl:270-272 */
static float gaussian(float sigma, float x_mu) {
  return expf(-(x_mu * x_mu) / ((2.0f * sigma) * sigma)) / (2.5066282f * fabsf(sigma));
}

/* inc_fresnel.glsl:14-24 */
static float fresnel_power_ratio(float etaRatio, float nA, float nB, float cosI, float sinI) {
  const float sinTSquared = powf(etaRatio * sinI, 2.0f);
  if (sinTSquared > 1.0f) return 1.0f;
  const float cosT = sqrtf(1.0f - sinTSquared);
  const float A = nA * cosI;
  const float B = nB * cosT;
  const float R = (A - B) / (A + B);
  return glm_min(1.0f, R * R);
}
/* inc_fresnel.glsl:29-31, 40-42, 51-57 */
static float fresnel(float etaOrigin, float etaPerp, float etaPar, float cosA, float sinA) {
  const float r = fresnel_power_ratio(etaOrigin / etaPerp, etaOrigin, etaPerp, cosA, sinA);
  const float t = fresnel_power_ratio(etaOrigin / etaPar, etaPar, etaOrigin, cosA, sinA);
  return r * (1.0f - 0.5f) + t * 0.5f;
}

/* inc_solver.glsl:16-18 */
static v4 solver_linear(float a, float b) {
  v4 r = { 0.f, 0.f, 0.f, 0.f };
  if (fabsf(a) > kEps) { r.x = -b / a; r.w = 1.0f; }
  return r;
}
/* inc_solver.glsl:22-35 */
static v4 solver_quadratic(float a, float b, float c) {
  v4 r = { 0.f, 0.f, 0.f, 0.f };
  if (fabsf(a) < kEps) return solver_linear(b, c);
  float delta = b * b - (4.0f * a) * c;
  if (delta < 0.0f) return r;
  delta = sqrtf(delta);
  r.x = (-b + delta) / (2.0f * a);
  r.y = (-b - delta) / (2.0f * a);
  r.w = 1.0f + ((delta < kEps) ? 0.0f : 1.0f);      /* 1 + step(Epsilon(), delta) */
  return r;
}
static float glm_sign(float x) { return (float)((0.0f < x) - (x < 0.0f)); }
/* inc_solver.glsl:51-90 */
static v4 solver_cubic_normalized(float a, float b, float c) {
  v4 roots = { 0.f, 0.f, 0.f, 0.f };
  if (fabsf(c) < kEps) {
    roots = solver_quadratic(1.0f, a, b);
    float* comp = &roots.x;
    comp[(int)roots.w] = 0.0f;
    roots.w += 1.0f;
  } else {
    const float Q = (3.0f * b - a * a) / 9.0f;
    const float R = (((9.0f * a) * b - 27.0f * c) - ((2.0f * a) * a) * a) / 54.0f;
    const float Q3 = (Q * Q) * Q;
    const float D = Q3 + R * R;
    const float third_a = a / 3.0f;
    if (D > 0.0f) {
      const float sqrtD = sqrtf(D);
      const float s = glm_sign(R + sqrtD) * powf(fabsf(R + sqrtD), 0.333f);
      const float t = glm_sign(R - sqrtD) * powf(fabsf(R - sqrtD), 0.333f);
      roots.x = (s + t) - third_a;
      roots.w = 1.0f;
    } else {
      const float theta = acosf(R * (1.0f / sqrtf(-Q3)));
      const float twoSqrtQ = 2.0f * sqrtf(-Q);
      roots.x = fmaf(twoSqrtQ, cosf(theta / 3.0f), -third_a);
      roots.y = fmaf(twoSqrtQ, cosf((theta + 2.0f * kPi) / 3.0f), -third_a);
      roots.z = fmaf(twoSqrtQ, cosf((theta + 4.0f * kPi) / 3.0f), -third_a);
      roots.w = 3.0f;
    }
  }
  return roots;
}
/* inc_solver.glsl:42-47 */
static v4 solver_cubic(float a, float b, float c, float d) {
  return (fabsf(a) < kEps) ? solver_quadratic(b, c, d) : solver_cubic_normalized(b / a, c / a, d / a);
}

/* inc_marschner_n.glsl:32-41 */
static v4 polynomial_coefficients(int p, float c, float phi) {
  const float kMinusEightOverPiCube = -0.25801227547f, kSixOverPi = 1.90985931710f;
  const v4 r = { ((float)p * c) * kMinusEightOverPiCube, 0.0f, ((float)p * c) * kSixOverPi - 2.0f, (float)p * kPi - phi };
  return r;
}
/* inc_marschner_n.glsl:47-53 */
static float inv_first_derivative_factor(v4 k, float gammaI, float cosGammaI) {
  const float d = ((3.0f * k.x) * powf(gammaI, 2.0f) + k.z) / cosGammaI;
  return 1.0f / fabsf(2.0f * d);
}
/* inc_marschner_n.glsl:66-95 */
static float calculate_absorption(int p, float absorption, float etaPerp, float etaPar, float cosGammaI, float sinGammaI) {
  if (p == 0) return fresnel(1.0f, etaPerp, etaPar, cosGammaI, sinGammaI);
  const float sinGammaT = sinGammaI / etaPerp;
  const float gammaT = asinf(sinGammaT);
  const float cosGammaT = cosf(gammaT);
  const float fi = fresnel(1.0f, etaPerp, etaPar, cosGammaI, sinGammaI);
  const float ft = fresnel(1.0f, 1.0f / etaPerp, 1.0f / etaPar, cosGammaT, sinGammaT);
  const float t = expf((-4.0f * absorption) * powf(cosGammaT, 2.0f));
  return (powf(1.0f - fi, 2.0f) * powf(ft, (float)(p - 1))) * powf(t, (float)p);
}
/* inc_marschner_n.glsl:100-131 */
static float Np(int p, float absorption, float etaPerp, float etaPar, float c, float phi) {
  const v4 k = polynomial_coefficients(p, c, phi);
  const v4 roots = solver_cubic(k.x, k.y, k.z, k.w);
  const int nRoots = (int)roots.w;
  const float* r = &roots.x;
  float L = 0.0f;
  for (int i = 0; i < nRoots; ++i) {
    const float gammaI = r[i];
    const float sinGammaI = sinf(gammaI), cosGammaI = cosf(gammaI);
    const float a = calculate_absorption(p, absorption, etaPerp, etaPar, cosGammaI, sinGammaI);
    L += a * inv_first_derivative_factor(k, gammaI, cosGammaI);
  }
  return glm_min(L, 1.0f);
}

/* params: eta, absorption, eccentricity, ar, br, glintScale, azimuthalWidth, deltaCaustic, deltaHm (marschner.h:38-52).
 * m_rgba / n_rgba: resolution * resolution * 4 floats, texel (x, y) at (y * resolution + x) * 4. */
void bho_marschner_luts(const float* params, int resolution, float* m_rgba, float* n_rgba) {
  const float eta = params[0], absorption = params[1], ar = params[3], br = params[4];
  const float inv = 1.0f / (float)resolution;                                   /* marschner.h:33 */
  const float shifts[3] = { 1.0f * ar, -0.5f * ar, -1.5f * ar };                /* cs_marschner_m.glsl:46 */
  const float widths[3] = { 1.0f * br, 0.5f * br, 2.0f * br };                  /* cs_marschner_m.glsl:47 */
  for (int y = 0; y < resolution; ++y)
    for (int x = 0; x < resolution; ++x) {
      float* m = m_rgba + 4 * ((size_t)y * resolution + x);
      float* n = n_rgba + 4 * ((size_t)y * resolution + x);
      {                                                                         /* cs_marschner_m.glsl:35-62 */
        const float sinThetaI = (2.0f * (float)x) * inv - 1.0f;
        const float sinThetaR = (2.0f * (float)y) * inv - 1.0f;
        const float thetaI = asinf(sinThetaI), thetaR = asinf(sinThetaR);
        const float thetaH = (thetaI + thetaR) / 2.0f;
        const float thetaD = (thetaI - thetaR) / 2.0f;
        const float degH = thetaH * 57.295779513082320876798154814105f;         /* glm::degrees */
        for (int k = 0; k < 3; ++k) m[k] = 1.0f * gaussian(widths[k], degH - shifts[k]);
        m[3] = 1.0f * cosf(thetaD);
      }
      {                                                                         /* cs_marschner_n.glsl:34-83 */
        const float cosPhiD = (2.0f * (float)x) * inv - 1.0f;
        const float cosThetaD = (2.0f * (float)y) * inv - 1.0f;
        const float sinThetaDSquared = 1.0f - powf(cosThetaD, 2.0f);
        const float refractionSquared = eta * eta;
        const float etaPerp = sqrtf(refractionSquared - sinThetaDSquared) / cosThetaD;
        const float etaPar = refractionSquared / etaPerp;
        const float phiD = acosf(cosPhiD);
        const float c = asinf(1.0f / etaPerp);
        n[0] = 1.0f * Np(0, absorption, etaPerp, etaPar, c, phiD);
        n[1] = 1.0f * Np(1, absorption, etaPerp, etaPar, c, phiD);
        n[2] = 1.0f * Np(2, absorption, etaPerp, etaPar, c, phiD);
        n[3] = 1.0f * 1.0f;
      }
    }
}

/* fp32 -> fp16, round to nearest even: what storing to the GL_RGBA16F image (marschner.h:30) does to a texel. */
void bho_float_to_half(const float* in, int64_t count, uint16_t* out) {
  for (int64_t i = 0; i < count; ++i) {
    uint32_t f; memcpy(&f, in + i, 4);
    const uint32_t sign = (f >> 16) & 0x8000u, a = f & 0x7fffffffu;
    uint16_t h;
    if (a >= 0x7f800000u) h = (uint16_t)(0x7c00u | (a > 0x7f800000u ? 0x200u : 0u));                 /* inf / NaN */
    else if (a >= 0x477ff000u) h = 0x7c00u;                                                          /* rounds to >= 65520: inf */
    else if (a < 0x33000001u) h = 0;                                                                 /* <= 2^-25: zero */
    else if (a < 0x38800000u) {                                                                      /* subnormal half */
      const uint32_t m = (a & 0x7fffffu) | 0x800000u; const int shift = 126 - (int)(a >> 23);        /* 14..24 */
      const uint32_t q = m >> shift, rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
      h = (uint16_t)(q + ((rem > half) || (rem == half && (q & 1u))));
    } else {
      const uint32_t q = (a - 0x38000000u) >> 13, rem = a & 0x1fffu;
      h = (uint16_t)(q + ((rem > 0x1000u) || (rem == 0x1000u && (q & 1u))));
    }
    out[i] = (uint16_t)(sign | h);
  }
}
