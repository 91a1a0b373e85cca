#!/usr/bin/env python3
"""Per-frame cost of Hair::update at the sizes the reference itself runs (its only scalp asset: 448 roots x 4 control points)
up to configs[0] (4,096 x 16): host time of the call (launch overhead) and device time per frame, 1 and 4 substeps, fused and not."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import barbu_b200 as bb
DT = float(np.float32(1.0) / np.float32(90.0))
for S, N in ((448, 4), (4096, 4), (4096, 16), (16384, 16), (65536, 16), (65536, 32), (262144, 32)):
    rows = 1
    while rows * rows < S: rows *= 2
    rows = min(rows, S); cols = S // rows
    if rows * cols != S: rows, cols = 1, S
    for math in ("exact", "fast"):
        for sub, fuse, policy in ((1, False, 0), (4, False, 0), (4, True, 0), (1, False, 1), (4, False, 1)):
            with bb.HairSim(S, N) as sim:
                sim.configure(scale=1.45, sphere=(0.0, 0.0, 0.0, 0.98), math=bb.BH_MATH_EXACT if math == "exact" else bb.BH_MATH_FAST)
                sim.set_substep_fusion(fuse); sim.set_step_policy(policy)
                st = torch.cuda.Stream(); sim.set_stream(st.cuda_stream)
                if rows > 1: sim.init_sphere_scalp(rows, cols, 0, bb.random_values(1234, 0, S))
                else:
                    rng = np.random.default_rng(0); root = rng.standard_normal((S, 3)).astype(np.float32); root /= np.linalg.norm(root, axis=1, keepdims=True)
                    sim.init_strands(root, root, bb.random_values(1234, 0, S))
                for _ in range(50): sim.step(DT, sub)
                torch.cuda.synchronize()
                n = 2000
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0 = time.perf_counter(); e0.record(st)
                for _ in range(n): sim.step(DT, sub)
                t_issue = time.perf_counter() - t0
                e1.record(st); torch.cuda.synchronize()
                t_wall = time.perf_counter() - t0
                print(json.dumps({"strands": S, "nverts": N, "math": math, "substeps": sub, "fused": fuse, "policy": "latency" if policy else "throughput", "kernel_kind": sim.kernel_kind,
                                  "host_us_per_frame": round(1e6 * t_issue / n, 2), "device_us_per_frame": round(1e3 * e0.elapsed_time(e1) / n, 2),
                                  "wall_us_per_frame": round(1e6 * t_wall / n, 2)}), flush=True)
