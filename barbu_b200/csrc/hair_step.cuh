// hair_step.cuh — device-side argument block and launcher declarations of the step kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bh {

constexpr int kMaxCapsules = 8;

struct Capsule { float ax, ay, az, bx, by, bz, r; };

// Everything a step launch needs, passed by value (kernel parameter space, < 400 B).
struct StepArgs {
  float4* pos;            // plane 0: xyz + rest length
  float4* vel;            // plane 1: xyz + 0
  long long nstrands;
  int nverts;
  int iterations;
  float dt, dt2;          // uTimeStep, dt*dt
  float fx, fy, fz;       // kForceCoeff * gravity (+ wind)
  float sf;               // uScaleFactor
  float damp;             // 0.80
  float cx, cy, cz, r, r2;// uBoundingSphere, r*r
  float keep;             // 1 - drag (extension)
  int use_drag;
  int ncaps;
  Capsule caps[kMaxCapsules];
};

// Launch one fused step (integrate -> K x (FTL, collide) -> velocity fix) in place.
// math: 0 exact, 1 fast. `tile_counter`: 4 bytes of device scratch owned by the caller and not shared with
// any launch that may run concurrently (the streaming kernel's tile scheduler). Returns the CUDA error.
cudaError_t launch_step(const StepArgs& a, int math, cudaStream_t stream, unsigned int* tile_counter);

// Which kernel launch_step picks: 0 = streaming (hair_stream.cu), 1 = per-strand pipelined, 2 = generic.
int step_kernel_kind(const StepArgs& a);

// hair_stream.cu
bool stream_kernel_eligible(const StepArgs& a);
cudaError_t selftest_inversesqrt(unsigned long long* mismatches);
cudaError_t launch_step_stream(const StepArgs& a, int math, cudaStream_t stream, unsigned int* tile_counter);

}  // namespace bh
