/*
 * barbu_hair_oracle.c — CPU ORACLE (test infrastructure only; see barbu_hair_oracle.h).
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math -fopenmp -fPIC -shared
 * `-ffp-contract=off` is load-bearing: every '*' and '+' below is one IEEE-754 binary32
 * rounding; the only fused operations are the explicit fmaf() calls.
 *
 * Each function cites the reference lines it restates (paths relative to the reference root).
 */
#include "barbu_hair_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- GLM-equivalent scalar helpers (third_party/glm/glm/detail/func_geometric.inl) ---------- */

/* compute_dot<vec3>: tmp = a*b; return tmp.x + tmp.y + tmp.z   (func_geometric.inl:48-55) */
static inline float dot3(const float a[3], const float b[3]) {
  const float tx = a[0] * b[0], ty = a[1] * b[1], tz = a[2] * b[2];
  return (tx + ty) + tz;
}
/* inversesqrt(x) = 1 / sqrt(x)   (func_exponential.inl:135-139) */
static inline float inversesqrt1(float x) { return 1.0f / sqrtf(x); }

void bho_default_params(bho_params* p) {
  memset(p, 0, sizeof(*p));
  p->dt = 1.0f / 90.0f;          /* core/global_clock.cc:160-162 */
  p->scale = 1.45f;              /* fx/hair.h:41 render.lengthScale -> uScaleFactor (hair.cc:108) */
  p->sphere[0] = 0.0f; p->sphere[1] = 0.0f; p->sphere[2] = 0.0f; p->sphere[3] = 1.0f; /* cs:43 */
  p->iterations = 8;             /* cs:197 */
  p->gravity[0] = 0.0f; p->gravity[1] = -9.81f; p->gravity[2] = 0.0f;               /* cs:71 */
  p->force_coeff = 20.0f;        /* cs:72 */
  p->damp = 0.80f;               /* cs:100-102 evaluated for lane 0 (di = 0) */
}

/* Closest-point capsule push-out; extension with no reference code. A capsule with a == b is
 * exactly CollideSphere (cs:129-139). Same operation order as the sphere branch. */
static inline void collide_capsule(const bho_capsule* c, float pos[3], float vel[3]) {
  float center[3] = { c->a[0], c->a[1], c->a[2] };
  const float ab[3] = { c->b[0] - c->a[0], c->b[1] - c->a[1], c->b[2] - c->a[2] };
  const float l2 = dot3(ab, ab);
  if (l2 > 0.0f) {
    const float ap[3] = { pos[0] - c->a[0], pos[1] - c->a[1], pos[2] - c->a[2] };
    float t = dot3(ap, ab) / l2;
    t = fminf(fmaxf(t, 0.0f), 1.0f);
    center[0] = c->a[0] + t * ab[0];
    center[1] = c->a[1] + t * ab[1];
    center[2] = c->a[2] + t * ab[2];
  }
  const float pt[3] = { pos[0] - center[0], pos[1] - center[1], pos[2] - center[2] };
  const float dp = dot3(pt, pt);
  if (dp < c->radius * c->radius) {
    const float inv = inversesqrt1(dp);
    const float n[3] = { pt[0] * inv, pt[1] * inv, pt[2] * inv };
    pos[0] = center[0] + c->radius * n[0];
    pos[1] = center[1] + c->radius * n[1];
    pos[2] = center[2] + c->radius * n[2];
    const float d = dot3(n, vel);
    vel[0] = vel[0] - (n[0] * d) * 2.0f;
    vel[1] = vel[1] - (n[1] * d) * 2.0f;
    vel[2] = vel[2] - (n[2] * d) * 2.0f;
  }
}

/* One workgroup (= one strand) of cs_simulation.glsl main(), lock-step semantics:
 * publish -> lane-0 serial pass -> everyone reloads.
 * X/W/rest are scratch arrays of nverts entries. */
static void step_strand(float* pos4, float* vel4, int N, const bho_params* p,
                        float (*X)[3], float (*W)[3], float* rest) {
  const float dt = p->dt;
  /* CalculateForces (cs:70-77): force = kForceCoeff * gravity (float * vec3), + wind extension. */
  float force[3];
  for (int c = 0; c < 3; ++c) force[c] = p->force_coeff * p->gravity[c];
  if (p->wind[0] != 0.0f || p->wind[1] != 0.0f || p->wind[2] != 0.0f)
    for (int c = 0; c < 3; ++c) force[c] = force[c] + p->wind[c];
  const float dt2 = dt * dt;                                  /* vec3 dt*dt, cs:182 */
  const float keep = 1.0f - p->drag;                          /* extension; drag == 0 -> unused */

  /* UnpackParticle (cs:53-61) + integrate (cs:179-193). */
  for (int i = 0; i < N; ++i) {
    const float* P = pos4 + 4 * (size_t)i;
    const float* V = vel4 + 4 * (size_t)i;
    rest[i] = P[3];
    if (i > 0) {
      for (int c = 0; c < 3; ++c) {
        float v = V[c];
        if (p->drag != 0.0f) v = v * keep;
        /* fma(dt*dt, force, fma(dt, velocity, position)) — fused, component-wise (cs:182) */
        X[i][c] = fmaf(dt2, force[c], fmaf(dt, v, P[c]));
        W[i][c] = V[c];
      }
    } else {
      /* vec3(mat4(1.0) * vec4(lastPosition, 1.0)) with GLM's mat4*vec4 order
       * (type_mat4x4.inl:563-575): (m0*x + m1*y) + (m2*z + m3*1).  cs:190-192 */
      static const float I4[4][4] = { {1,0,0,0}, {0,1,0,0}, {0,0,1,0}, {0,0,0,1} };
      for (int c = 0; c < 3; ++c) {
        const float add0 = I4[0][c] * P[0] + I4[1][c] * P[1];
        const float add1 = I4[2][c] * P[2] + I4[3][c] * 1.0f;
        X[0][c] = add0 + add1;
        W[0][c] = X[0][c] - P[c];
      }
    }
  }

  const float r = p->sphere[3];
  const float r2 = r * r;                                     /* radius * radius, cs:133 */
  /* SatisfyConstraints (cs:155-161) */
  for (int k = 0; k < p->iterations; ++k) {
    /* DistanceConstraint, lane 0 only (cs:109-122) */
    for (int i = 1; i < N; ++i) {
      const float* p0 = X[i - 1];
      const float p1[3] = { X[i][0], X[i][1], X[i][2] };
      const float vd[3] = { p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2] };
      const float inv = inversesqrt1(dot3(vd, vd));           /* normalize = v * inversesqrt(dot) */
      const float L = p->scale * rest[i];                     /* (uScaleFactor * restlength) */
      for (int c = 0; c < 3; ++c) {
        const float nc = vd[c] * inv;
        const float bis = p0[c] + L * nc;                     /* p0 + L * normalize(vdiff) */
        X[i][c] = bis;
        W[i][c] = bis - p1[c];                                /* velocity = p1_bis - p1 */
      }
    }
    for (int i = 1; i < N - 1; ++i)                           /* cs:119-121 (ascending) */
      for (int c = 0; c < 3; ++c) W[i][c] = W[i + 1][c] * p->damp;
    /* CollisionConstraint, index > 0 (cs:149-153) -> CollideSphere(+1, ...) (cs:129-139) */
    for (int i = 1; i < N; ++i) {
      const float pt[3] = { X[i][0] - p->sphere[0], X[i][1] - p->sphere[1], X[i][2] - p->sphere[2] };
      const float dp = dot3(pt, pt);
      if (dp < r2) {
        const float inv = inversesqrt1(dp);
        const float n[3] = { pt[0] * inv, pt[1] * inv, pt[2] * inv };
        for (int c = 0; c < 3; ++c) X[i][c] = p->sphere[c] + r * n[c];
        const float d = dot3(n, W[i]);                        /* reflect: I - N*dot(N,I)*2 */
        for (int c = 0; c < 3; ++c) W[i][c] = W[i][c] - (n[c] * d) * 2.0f;
      }
      for (int q = 0; q < p->ncapsules; ++q) collide_capsule(&p->capsules[q], X[i], W[i]);
    }
  }

  /* PackParticle (cs:63-66) */
  for (int i = 0; i < N; ++i) {
    float* P = pos4 + 4 * (size_t)i;
    float* V = vel4 + 4 * (size_t)i;
    P[0] = X[i][0]; P[1] = X[i][1]; P[2] = X[i][2]; P[3] = rest[i];
    V[0] = W[i][0]; V[1] = W[i][1]; V[2] = W[i][2]; V[3] = 0.0f;
  }
}

void bho_step_mt(float* pos4, float* vel4, int64_t S, int N, const bho_params* p, int nthreads) {
  if (S <= 0 || N <= 0) return;
#ifdef _OPENMP
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
#endif
  {
    float (*X)[3] = malloc(sizeof(float[3]) * (size_t)N);
    float (*W)[3] = malloc(sizeof(float[3]) * (size_t)N);
    float* rest = malloc(sizeof(float) * (size_t)N);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int64_t s = 0; s < S; ++s)
      step_strand(pos4 + 4 * (size_t)s * N, vel4 + 4 * (size_t)s * N, N, p, X, W, rest);
    free(X); free(W); free(rest);
  }
}

void bho_step(float* pos4, float* vel4, int64_t S, int N, const bho_params* p) {
  bho_step_mt(pos4, vel4, S, N, p, 1);
}

/* fx/hair.cc:273-275 (rand() seeded by the harness instead of app.cc:96-97 time(NULL)). */
void bho_random_values(unsigned seed, int64_t S, float* out) {
  srand(seed);
  for (int64_t j = 0; j < S; ++j)
    out[j] = (float)(1.0 + 0.1 * (1.0 - 2.0 * (double)rand() / (double)RAND_MAX));
}

/* fx/hair.cc:255-287 */
void bho_init_strands(const float* root_pos3, const float* root_nrm3, const float* random_value,
                      int64_t S, int N, float maxlength, float* pos4, float* vel4) {
  const float scaleOffset = maxlength / (float)N;             /* hair.cc:265 */
  for (int64_t j = 0; j < S; ++j) {
    const float* v = root_pos3 + 3 * j;
    const float* n = root_nrm3 + 3 * j;
    const float random_value_j = random_value[j];
    float lastOffset = 0.0f;
    for (int i = 0; i < N; ++i) {
      const size_t idx = (size_t)j * N + i;
      const float offset = (float)i * scaleOffset * random_value_j;   /* hair.cc:281-283 */
      for (int c = 0; c < 3; ++c) pos4[4 * idx + c] = v[c] + offset * n[c];
      pos4[4 * idx + 3] = offset - lastOffset;                /* rest length, hair.cc:284 */
      for (int c = 0; c < 4; ++c) vel4[4 * idx + c] = 0.0f;   /* hair.cc:285 */
      lastOffset = offset;
    }
  }
}

/* ---- glm::simplex(vec2) — third_party/glm/glm/gtc/noise.inl:591-645, detail/_noise.hpp:9-18 --- */
static inline float mod289f(float x) { return x - floorf(x * (1.0f / 289.0f)) * 289.0f; }
static inline float permutef(float x) { return mod289f(((x * 34.0f) + 1.0f) * x); }
static inline float fractf_(float x) { return x - floorf(x); }   /* glm::fract */

float bho_simplex2(float vx, float vy) {
  const float Cx = (float)0.211324865405187, Cy = (float)0.366025403784439;
  const float Cz = (float)-0.577350269189626, Cw = (float)0.024390243902439;
  /* i = floor(v + dot(v, vec2(C.y))) ; dot(vec2) = tmp.x + tmp.y */
  const float d0 = vx * Cy + vy * Cy;
  float ix = floorf(vx + d0), iy = floorf(vy + d0);
  /* x0 = v - i + dot(i, vec2(C.x)) */
  const float d1 = ix * Cx + iy * Cx;
  const float x0x = vx - ix + d1, x0y = vy - iy + d1;
  const float i1x = (x0x > x0y) ? 1.0f : 0.0f, i1y = (x0x > x0y) ? 0.0f : 1.0f;
  float x12x = x0x + Cx, x12y = x0y + Cx;
  const float x12z = x0x + Cz, x12w = x0y + Cz;
  x12x = x12x - i1x; x12y = x12y - i1y;
  /* i = mod(i, vec2(289)) : x - y*floor(x/y) */
  ix = ix - 289.0f * floorf(ix / 289.0f);
  iy = iy - 289.0f * floorf(iy / 289.0f);
  /* p = permute(permute(i.y + vec3(0, i1.y, 1)) + i.x + vec3(0, i1.x, 1)) */
  const float a[3] = { iy + 0.0f, iy + i1y, iy + 1.0f };
  const float b[3] = { 0.0f, i1x, 1.0f };
  float pz[3];
  for (int c = 0; c < 3; ++c) pz[c] = permutef(permutef(a[c]) + ix + b[c]);
  /* m = max(0.5 - vec3(dot(x0,x0), dot(x12.xy,x12.xy), dot(x12.zw,x12.zw)), 0) ; m = m^4 */
  float m[3] = { 0.5f - (x0x * x0x + x0y * x0y), 0.5f - (x12x * x12x + x12y * x12y),
                 0.5f - (x12z * x12z + x12w * x12w) };
  for (int c = 0; c < 3; ++c) { m[c] = (m[c] < 0.0f) ? 0.0f : m[c]; m[c] = m[c] * m[c]; m[c] = m[c] * m[c]; }
  float g[3], a0[3], h[3];
  for (int c = 0; c < 3; ++c) {
    const float x = 2.0f * fractf_(pz[c] * Cw) - 1.0f;
    h[c] = fabsf(x) - 0.5f;
    const float ox = floorf(x + 0.5f);
    a0[c] = x - ox;
    m[c] = m[c] * ((float)1.79284291400159 - (float)0.85373472095314 * (a0[c] * a0[c] + h[c] * h[c]));
  }
  g[0] = a0[0] * x0x + h[0] * x0y;
  g[1] = a0[1] * x12x + h[1] * x12y;
  g[2] = a0[2] * x12z + h[2] * x12w;
  return 130.0f * dot3(m, g);
}

/* fx/hair.cc:290-328 */
void bho_init_tangents(const float* root_nrm3, int64_t S, int N, float maxlength, float* tan4) {
  const float inv_nroots = 1.0f / (float)S;
  const float kPi = (float)3.14159265358979323846264338327950288;
  const float scaleMaxLength = 0.125f * sqrtf(maxlength);
  float curly[3] = { 0.0f, 0.0f, 0.0f };
  for (int64_t j = 0; j < S; ++j) {
    const size_t A = (size_t)j * N, B = A + (size_t)(N - 1);
    const float* nr = root_nrm3 + 3 * j;
    const float dj = (float)(j + 1) * inv_nroots;
    {
      const float n = 1.25f * bho_simplex2(sinf(3.0f * dj), cosf(5.0f));
      curly[0] = cosf(n * 4.0f * kPi); curly[1] = -0.71f * n; curly[2] = sinf(n * 2.7f * kPi);
    }
    for (int c = 0; c < 3; ++c) tan4[4 * A + c] = .15f * nr[c];
    tan4[4 * A + 3] = .15f * 0.0f;
    for (int c = 0; c < 3; ++c) tan4[4 * B + c] = .2f * (-nr[c] + curly[c]);
    tan4[4 * B + 3] = .2f * 0.0f;
    const float dist_AB = (float)(B - A);
    const float inv_dist = 1.0f / dist_AB;
    for (size_t i = A + 1; i < B; ++i) {
      const float di = 10.0f * (float)(B - i) / (dist_AB - 1.0f);
      const float n = di * bho_simplex2(sinf(43.0f * dj), cosf(5.0f * di));
      curly[0] = -5.8f * (10.7f * cosf(n * kPi));
      curly[1] = -5.8f * (-2.3f * n);
      curly[2] = -5.8f * (20.5f * sinf(n * kPi));
      const float s = 0.1f * (float)(i - A) * inv_dist * scaleMaxLength;
      for (int c = 0; c < 3; ++c) tan4[4 * i + c] = s * curly[c];
      tan4[4 * i + 3] = s * 0.0f;
    }
  }
}

/* fx/hair.cc:397-409 */
int bho_patch_indices(const int32_t* tri, int64_t F, int N, int32_t* out) {
  const int nseg = N - 1;
  size_t idx = 0;
  for (int64_t i = 0; i < F; ++i)
    for (int j = 0; j < nseg; ++j)
      for (int k = 0; k < 3; ++k) {
        const int64_t e = (int64_t)N * tri[3 * i + k] + j;
        if (e + 1 > INT32_MAX) return -1;
        out[idx++] = (int32_t)e;
        out[idx++] = (int32_t)(e + 1);
      }
  return 0;
}

/* Synthetic scalp, SURVEY.md §8(d). Row/column trigonometry is evaluated in double by libm and
 * rounded to fp32 once; the per-vertex products are single fp32 multiplications. */
void bho_sphere_scalp(int R, int C, float* pos3, float* nrm3, int32_t* tri) {
  const double kPiD = 3.14159265358979323846;
  for (int r = 0; r < R; ++r) {
    const double th = kPiD * ((double)r + 0.5) / (double)R - kPiD / 2.0;
    const float ct = (float)cos(th), st = (float)sin(th);
    for (int c = 0; c < C; ++c) {
      const double ph = 2.0 * kPiD * (double)c / (double)C;
      const float cp = (float)cos(ph), sp = (float)sin(ph);
      const size_t v = (size_t)r * C + c;
      const float n[3] = { ct * cp, st, ct * sp };
      for (int k = 0; k < 3; ++k) { nrm3[3 * v + k] = n[k]; pos3[3 * v + k] = n[k]; }
    }
  }
  if (!tri) return;
  size_t t = 0;
  for (int r = 0; r + 1 < R; ++r)
    for (int c = 0; c < C; ++c) {
      const int c1 = (c + 1) % C;
      const int32_t v00 = r * C + c, v10 = (r + 1) * C + c, v01 = r * C + c1, v11 = (r + 1) * C + c1;
      tri[t++] = v00; tri[t++] = v10; tri[t++] = v01;
      tri[t++] = v01; tri[t++] = v10; tri[t++] = v11;
    }
}

/* Extension, no reference caller for hair (cs_simulation.glsl:186-192 says roots "should be" skinned).
 * Formula of src/shaders/shared/inc_skinning.glsl: apply_skinning (l.22-31): early-out when
 * weights.x <= 1e-6, w.w = 1 - (x+y+z); skinning_DQBS (l.54-82): antipodality fix
 * weights.xyz *= sign(dot(q[3], q[k])) (sign(0) = 0), A = Ma*w, B = Mb*w in GLM mat4*vec4 order
 * (m0*x + m1*y) + (m2*z + m3*w), normalise by inversesqrt(dot(A,A)), rotate + translate. Pinned bit for bit against
 * the shader source compiled over the reference's GLM (oracle/_ref, tests/test_oracle_vs_reference_live.py).
 * dq palette: njoints * 8 floats (real xyzw, dual xyzw), as uSkinningDatas texels 2j, 2j+1. */
static inline void cross3(const float a[3], const float b[3], float o[3]) {
  o[0] = a[1] * b[2] - b[1] * a[2];   /* glm::cross: x.y*y.z - y.y*x.z, ... */
  o[1] = a[2] * b[0] - b[2] * a[0];
  o[2] = a[0] * b[1] - b[0] * a[1];
}
static inline void dq_rotate(const float A[4], float v[3]) {
  /* v += 2 * cross(A.xyz, cross(A.xyz, v) + A.w*v) */
  float c1[3], c2[3];
  cross3(A, v, c1);
  for (int c = 0; c < 3; ++c) c1[c] = c1[c] + A[3] * v[c];
  cross3(A, c1, c2);
  for (int c = 0; c < 3; ++c) v[c] = v[c] + 2.0f * c2[c];
}
void bho_skin_roots_dq(const float* rest_pos3, const float* rest_nrm3, const int32_t* joints4,
                       const float* weights3, const float* dq, int64_t S,
                       float* out_pos3, float* out_nrm3) {
  for (int64_t s = 0; s < S; ++s) {
    float v[3] = { rest_pos3[3 * s], rest_pos3[3 * s + 1], rest_pos3[3 * s + 2] };
    float n[3] = { rest_nrm3[3 * s], rest_nrm3[3 * s + 1], rest_nrm3[3 * s + 2] };
    float w[4] = { weights3[3 * s], weights3[3 * s + 1], weights3[3 * s + 2], 0.0f };
    if (!(w[0] <= 1e-6f)) {
      w[3] = 1.0f - ((w[0] + w[1]) + w[2]);
      const float* q[4];
      for (int k = 0; k < 4; ++k) q[k] = dq + 8 * (size_t)joints4[4 * s + k];
      for (int k = 0; k < 3; ++k) {
        /* vec4 * mat3x4 -> dot(vec4, column k), GLM: ((x*x + y*y) + z*z) + w*w */
        const float d = ((q[3][0] * q[k][0] + q[3][1] * q[k][1]) + q[3][2] * q[k][2]) + q[3][3] * q[k][3];
        const float sg = (d > 0.0f) ? 1.0f : ((d < 0.0f) ? -1.0f : 0.0f);
        w[k] = w[k] * sg;
      }
      float A[4], B[4];
      for (int c = 0; c < 4; ++c) {
        A[c] = (q[0][c] * w[0] + q[1][c] * w[1]) + (q[2][c] * w[2] + q[3][c] * w[3]);
        B[c] = (q[0][4 + c] * w[0] + q[1][4 + c] * w[1]) + (q[2][4 + c] * w[2] + q[3][4 + c] * w[3]);
      }
      /* glm::dot(vec4, vec4) = (x*x + y*y) + (z*z + w*w)  (glm/detail/func_geometric.inl:58-65) */
      const float inv = 1.0f / sqrtf((A[0] * A[0] + A[1] * A[1]) + (A[2] * A[2] + A[3] * A[3]));
      for (int c = 0; c < 4; ++c) { A[c] = A[c] * inv; B[c] = B[c] * inv; }
      dq_rotate(A, v);
      float cab[3];
      cross3(A, B, cab);
      for (int c = 0; c < 3; ++c) v[c] = v[c] + 2.0f * ((A[3] * B[c] - B[3] * A[c]) + cab[c]);
      dq_rotate(A, n);
    }
    for (int c = 0; c < 3; ++c) { out_pos3[3 * s + c] = v[c]; out_nrm3[3 * s + c] = n[c]; }
  }
}

uint64_t bho_fnv1a64(const void* data, uint64_t n) {
  const unsigned char* p = (const unsigned char*)data;
  uint64_t h = 1469598103934665603ull;
  for (uint64_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
  return h;
}

/* ---- tess-stream stage: the consumer right after the simulation -------------------------
 * Restates, per output vertex, what the reference's VS -> TCS -> TES -> GS transform-feedback pass computes
 * (src/shaders/hair/02_tess_stream/*.glsl; draw call src/fx/hair.cc:141-173):
 *   patch = 6 control points (3 master strands x the two ends of one segment, hair.cc:397-409);
 *   TCS  tangents *= uScaleFactor (tcs_stream_hair.glsl:41); isolines: outer[0] = nlines, outer[1] = nsubsegments;
 *   TES  p_k = hermite_mix(P[2k], P[2k+1], T[2k], T[2k+1], x)   (tes:33-35, shared/inc_maths.glsl:210-228)
 *        position = sample_triangle2(p0, p1, p2, rand.xy)        (tes:59-60, inc_maths.glsl:107-115)
 *        relPos   = smoothstep2(0, 1, relPos(first CP) + x / N)  (tes:63-64, inc_maths.glsl:239-241,263-266;
 *                                                                 vs_stream_hair.glsl:24: (gl_VertexID % N) / float(N))
 *   GS   each isoline segment leaves as two vec4 (position.xyz, relPos) (gs_stream_hair.glsl:21-29), GL_LINES.
 * Defined here because the reference leaves it to the implementation (no reference parity is claimed for them):
 *   - isoline tess coordinates: x = k / nsubsegments (k = 0..nsubsegments), y = line / nlines (equal_spacing ideal);
 *   - primitive order: instance-major, then patch, line, segment;
 *   - vector arithmetic order: as the reference's vendored GLM evaluates the same expressions — vec4 * mat4 = per column
 *     ((x*x' + y*y') + z*z') + w*w', mat3x4 * vec3 = (m0*v.x + m1*v.y) + m2*v.z — pinned bit for bit by
 *     tests/test_oracle_vs_reference_live.py against the shader stages compiled over GLM (oracle/_ref);
 *   - the random pair: the reference indexes a std430 `vec3[]` view of 4096 mt19937(random_device) floats with
 *     int(y*40 + instance) % 4096 (tes:52-54), reading out of bounds past element 1023 (SURVEY §8 a-ext). Kept: the index
 *     formula; replaced: the table, by the counter-based hash below (seeded, reproducible, no out-of-bounds). */
static uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
static float u01(uint32_t h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }
void bho_tess_random_pair(uint32_t seed, int index, float st[2]) {
  st[0] = u01(lowbias32(seed + 0x9E3779B9u * (uint32_t)(2 * index + 1)));
  st[1] = u01(lowbias32(seed + 0x9E3779B9u * (uint32_t)(2 * index + 2)));
}
/* vec4 * mat4, one column: GLM sums the four products left to right (glm/detail/type_mat4x4.inl:584-595) */
static inline float dot4(const float a[4], const float b[4]) { return ((a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]) + a[3] * b[3]; }
/* hermite_mix: vU * mHermit * B, xyz only (w = h0 + h1 is dropped by the TES) */
static void hermite3(const float p0[3], const float p1[3], const float t0[3], const float t1[3], float u, float out[3]) {
  static const float M[4][4] = { { 2.0f, -3.0f, 0.0f, 1.0f }, { -2.0f, 3.0f, 0.0f, 0.0f }, { 1.0f, -2.0f, 1.0f, 0.0f }, { 1.0f, -1.0f, 0.0f, 0.0f } };
  const float vU[4] = { u * u * u, u * u, u, 1.0f };
  const float h[4] = { dot4(vU, M[0]), dot4(vU, M[1]), dot4(vU, M[2]), dot4(vU, M[3]) };
  for (int c = 0; c < 3; ++c) {
    const float col[4] = { p0[c], p1[c], t0[c], t1[c] };
    out[c] = dot4(h, col);
  }
}
int64_t bho_tess_stream_count(int64_t npatches, int ninstances, int nlines, int nsubsegments) {
  return npatches * ninstances * nlines * nsubsegments * 2;
}
void bho_tess_stream(const float* pos4, const float* tan4, const int32_t* patch_indices, int64_t npatches, int nverts,
                     float scale, int ninstances, int nlines, int nsubsegments, uint32_t seed, float* out4) {
  const float invN = 1.0f;  (void)invN;
  for (int inst = 0; inst < ninstances; ++inst)
    for (int64_t pa = 0; pa < npatches; ++pa) {
      const int32_t* e = patch_indices + 6 * pa;
      float P[6][3], T[6][3];
      for (int k = 0; k < 6; ++k)
        for (int c = 0; c < 3; ++c) { P[k][c] = pos4[4 * (int64_t)e[k] + c]; T[k][c] = tan4[4 * (int64_t)e[k] + c] * scale; }
      const float rel0 = (float)(e[0] % nverts) / (float)nverts;
      for (int line = 0; line < nlines; ++line) {
        const float y = (float)line / (float)nlines;
        float st[2];
        bho_tess_random_pair(seed, (int)(y * 40.0f + (float)inst) % 4096, st);
        if (st[0] + st[1] > 1.0f) {                       /* sample_triangle2 */
          st[0] = fmaxf(st[0], st[1]);
          st[1] = fminf(st[0], st[1]);
          st[0] = 1.0f - st[0];
        }
        const float cz = 1.0f - (st[0] + st[1]);
        float pts[3];
        for (int k = 0; k <= nsubsegments; ++k) {
          const float x = (float)k / (float)nsubsegments;
          float q0[3], q1[3], q2[3];
          hermite3(P[0], P[1], T[0], T[1], x, q0);
          hermite3(P[2], P[3], T[2], T[3], x, q1);
          hermite3(P[4], P[5], T[4], T[5], x, q2);
          for (int c = 0; c < 3; ++c) pts[c] = (q0[c] * st[0] + q1[c] * st[1]) + q2[c] * cz;
          float t = (rel0 + x / (float)nverts - 0.0f) / (1.0f - 0.0f);
          t = fminf(fmaxf(t, 0.0f), 1.0f);
          const float rel = t * t * t * (10.0f + t * (-15.0f + 6.0f * t));
          const int64_t base = ((((int64_t)inst * npatches + pa) * nlines + line) * nsubsegments) * 2;
          if (k < nsubsegments) { float* o = out4 + 4 * (base + 2 * k); o[0] = pts[0]; o[1] = pts[1]; o[2] = pts[2]; o[3] = rel; }
          if (k > 0) { float* o = out4 + 4 * (base + 2 * (k - 1) + 1); o[0] = pts[0]; o[1] = pts[1]; o[2] = pts[2]; o[3] = rel; }
        }
      }
    }
}

/* RawMeshData::recalculateNormals (src/utils/raw_mesh_file.cc:11-50), the path MeshData::setup takes for a mesh that came
 * without normals (src/memory/resources/mesh_data.cc:366-372): per face n = normalize(cross(v2 - v1, v3 - v2)) added,
 * unweighted, to its three vertices (l.14-32); every vertex sum normalised (l.34-36); then EVERY corner gets a normal
 * entry of its own, a copy of its vertex normal (l.38-49) — so the re-indexing of mesh_data.cc:384-406 finds 3 * F unique
 * (v, vt, vn) triples: one hair root per face corner. GLM: cross (func_geometric.inl:74-77), dot = (x*x + y*y) + z*z
 * (l.52-53), normalize = v * inversesqrt(dot) (l.88), inversesqrt = 1 / sqrt (func_exponential.inl:135-139).
 * corner_v: 3 * nfaces zero-based position indices; nrm3_corner: 3 * nfaces normals. */
void bho_recalc_normals(const float* pos3, int64_t nverts, const int32_t* corner_v, int64_t nfaces, float* nrm3_corner) {
  float* acc = (float*)calloc((size_t)(nverts > 0 ? nverts : 1) * 3, sizeof(float));
  for (int64_t f = 0; f < nfaces; ++f) {
    const int32_t i1 = corner_v[3 * f], i2 = corner_v[3 * f + 1], i3 = corner_v[3 * f + 2];
    const float *v1 = pos3 + 3 * (size_t)i1, *v2 = pos3 + 3 * (size_t)i2, *v3 = pos3 + 3 * (size_t)i3;
    const float u[3] = { v2[0] - v1[0], v2[1] - v1[1], v2[2] - v1[2] };
    const float v[3] = { v3[0] - v2[0], v3[1] - v2[1], v3[2] - v2[2] };
    const float c[3] = { u[1] * v[2] - v[1] * u[2], u[2] * v[0] - v[2] * u[0], u[0] * v[1] - v[0] * u[1] };
    const float inv = inversesqrt1(dot3(c, c));
    const int32_t ids[3] = { i1, i2, i3 };
    for (int k = 0; k < 3; ++k)
      for (int a = 0; a < 3; ++a) acc[3 * (size_t)ids[k] + a] += c[a] * inv;
  }
  for (int64_t i = 0; i < nverts; ++i) {
    float* n = acc + 3 * (size_t)i;
    const float inv = inversesqrt1(dot3(n, n));
    n[0] *= inv; n[1] *= inv; n[2] *= inv;
  }
  for (int64_t q = 0; q < 3 * nfaces; ++q)
    for (int a = 0; a < 3; ++a) nrm3_corner[3 * q + a] = acc[3 * (size_t)corner_v[q] + a];
  free(acc);
}

/* SkeletonController::generate_skinning_datas (src/fx/animation/skeleton_controller.cc:248-265): per joint
 * skin = global_pose * inverse_bind (GLM mat4 product, type_mat4x4.inl:643-646: ((A0*b0 + A1*b1) + A2*b2) + A3*b3 per column),
 * the 3x4 skinning matrix is the first three columns of its transpose (l.255), and the dual quaternion is
 * glm::dualquat(mat3x4) = dualquat_cast (gtx/dual_quaternion.inl:303-351): rotation by the largest-diagonal branch, then
 * dual = 0.5 * t * real spelled out. Matrices are GLM's: 16 floats, column-major. dq8: (real xyzw, dual xyzw) per joint —
 * the two RGBA32F texels per joint inc_skinning.glsl:37-52 fetches. */
void bho_dq_palette_from_matrices(const float* global_pose16, const float* inverse_bind16, int njoints, float* dq8) {
  for (int j = 0; j < njoints; ++j) {
    const float* A = global_pose16 + 16 * (size_t)j; const float* B = inverse_bind16 + 16 * (size_t)j;
    float M[4][4];                                            /* M[col][row] */
    for (int c = 0; c < 4; ++c)
      for (int r = 0; r < 4; ++r)
        M[c][r] = ((A[0 * 4 + r] * B[c * 4 + 0] + A[1 * 4 + r] * B[c * 4 + 1]) + A[2 * 4 + r] * B[c * 4 + 2]) + A[3 * 4 + r] * B[c * 4 + 3];
    float x[3][4];                                            /* x[c] = column c of transpose(M) = row c of M */
    for (int c = 0; c < 3; ++c)
      for (int k = 0; k < 4; ++k) x[c][k] = M[k][c];
    float q[4];                                               /* real: x y z w */
    const float trace = x[0][0] + x[1][1] + x[2][2];
    if (trace > 0.0f) {
      const float r = sqrtf(1.0f + trace), invr = 0.5f / r;
      q[3] = 0.5f * r; q[0] = (x[2][1] - x[1][2]) * invr; q[1] = (x[0][2] - x[2][0]) * invr; q[2] = (x[1][0] - x[0][1]) * invr;
    } else if (x[0][0] > x[1][1] && x[0][0] > x[2][2]) {
      const float r = sqrtf(1.0f + x[0][0] - x[1][1] - x[2][2]), invr = 0.5f / r;
      q[0] = 0.5f * r; q[1] = (x[1][0] + x[0][1]) * invr; q[2] = (x[0][2] + x[2][0]) * invr; q[3] = (x[2][1] - x[1][2]) * invr;
    } else if (x[1][1] > x[2][2]) {
      const float r = sqrtf(1.0f + x[1][1] - x[0][0] - x[2][2]), invr = 0.5f / r;
      q[0] = (x[1][0] + x[0][1]) * invr; q[1] = 0.5f * r; q[2] = (x[2][1] + x[1][2]) * invr; q[3] = (x[0][2] - x[2][0]) * invr;
    } else {
      const float r = sqrtf(1.0f + x[2][2] - x[0][0] - x[1][1]), invr = 0.5f / r;
      q[0] = (x[0][2] + x[2][0]) * invr; q[1] = (x[2][1] + x[1][2]) * invr; q[2] = 0.5f * r; q[3] = (x[1][0] - x[0][1]) * invr;
    }
    float* o = dq8 + 8 * (size_t)j;
    o[0] = q[0]; o[1] = q[1]; o[2] = q[2]; o[3] = q[3];
    o[4] = 0.5f * (x[0][3] * q[3] + x[1][3] * q[2] - x[2][3] * q[1]);
    o[5] = 0.5f * (-x[0][3] * q[2] + x[1][3] * q[3] + x[2][3] * q[0]);
    o[6] = 0.5f * (x[0][3] * q[1] - x[1][3] * q[0] + x[2][3] * q[3]);
    o[7] = -0.5f * (x[0][3] * q[0] + x[1][3] * q[1] + x[2][3] * q[2]);
  }
}
