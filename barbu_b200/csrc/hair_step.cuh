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
  int reverse;            // streaming kernel: hand the tiles out from the last one down (see bh_step: alternates per launch)
  int tip_step;           // streaming kernel: step of a root chunk at which the tip of the previous strand leaves the pipeline
  // Streaming kernel, frame-level fusion: `passes` > 1 runs that many consecutive steps (the substeps of a frame, same dt) in
  // ONE launch. A warp takes `group_tiles` tiles at a time through all passes before it asks for the next group, so a pass
  // re-reads what the previous one stored while it is still in L2: HBM sees 64 B per vertex per LAUNCH instead of per step.
  // group_tiles * chunks-per-strand >= 4 (launcher): the store of a chunk is then always at least one bulk group older
  // than the load of the same chunk for the next pass (the tile ring prefetches two chunks ahead).
  int passes;
  int group_tiles;
  int prefer_latency;     // bh_set_step_policy: take the latency-oriented kernel (hair_wave.cu) when the shape allows it
  Capsule caps[kMaxCapsules];
  // Filled by the streaming launcher: a capsule-shaped bound in front of the exact capsule arithmetic: axis (b - a), 1 / |b - a|^2 (0 for a
  // degenerate capsule) and the squared radius with its margin, for a packed fast-arithmetic distance to the axis.
  float capt[kMaxCapsules][8];   // abx, aby, abz, inv_l2, r2_tight, r_tight (rounded up), 0, 0
  // ... and in front of both, the shell around the sphere's centre that holds all capsules (squared radii, with margins):
  // the step already has every vertex's squared distance to that centre.
  float cap_lo2, cap_hi2;
  // ... and the error terms of the temporal bound (stream_step, capsule variant): per step a lane's lower bound on its distance
  // to the capsules shrinks by its largest rest length plus cap_e0 + cap_e1 * max |p - c|^2 (rounding of the positions and of
  // the distance evaluation, generously)
  float cap_e0, cap_e1;
  // ... and the per-capsule constants of the exact chain as capsule_center() / collide_pos() compute them (single IEEE
  // operations in their order: the host's float arithmetic gives the same bits): b - a, |b - a|^2, r * r
  float capx[kMaxCapsules][8];   // abx, aby, abz, l2, r2, 0, 0, 0
  float r2_maybe;         // streaming kernel, exact profile: r^2 (1 + 2^-20) — above it a contracted |p - c|^2 rules a push-out out
};

// Launch one fused step (integrate -> K x (FTL, collide) -> velocity fix) in place.
// math: 0 exact, 1 fast. `tile_counter`: kSchedWords 32-bit words of device scratch owned by the caller, initialised
// with init_sched_words() and not shared with any launch that may run concurrently (words 0-1: the streaming kernel's
// tile scheduler, left zero again by every launch; words 2-3: the constant (-0.0f, -0.0f)). Returns the CUDA error.
constexpr int kSchedWords = 4;
inline void init_sched_words(unsigned int* host_words) { host_words[0] = host_words[1] = 0u; host_words[2] = host_words[3] = 0x80000000u; }
cudaError_t launch_step(const StepArgs& a, int math, cudaStream_t stream, unsigned int* tile_counter);

// Which kernel launch_step picks: 0 = streaming (hair_stream.cu), 1 = per-strand pipelined, 2 = generic, 3 = latency-oriented
// wavefront (hair_wave.cu; only with StepArgs::prefer_latency).
int step_kernel_kind(const StepArgs& a);

// hair_stream.cu
bool stream_kernel_eligible(const StepArgs& a);
// Whether `passes` consecutive steps of this shape can run as one fused launch of the streaming kernel (see StepArgs::passes).
// `always` = false adds the question whether it pays: a group of tiles runs all passes on ONE warp, so on a shard with fewer
// groups than the GPU has resident warps fusion trades parallelism for L2 reuse and launch overhead — measured
// (tools/small_latency.py) it wins only when a group is a single tile (>= 25 vertices per strand) or there are >= 1024 groups.
bool stream_fusion_eligible(const StepArgs& a, int passes, bool always = true);
cudaError_t selftest_inversesqrt(unsigned long long* mismatches);
cudaError_t launch_step_stream(const StepArgs& a, int math, cudaStream_t stream, unsigned int* tile_counter);

// hair_wave.cu: 8 lanes per strand, one constraint iteration each (8 iterations, any nverts, sphere + capsules); runs
// StepArgs::passes consecutive steps in one launch.
bool wave_kernel_eligible(const StepArgs& a);
cudaError_t launch_step_wave(const StepArgs& a, int math, cudaStream_t stream);

}  // namespace bh
