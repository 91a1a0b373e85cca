#!/usr/bin/env python3
"""north_star: "<= 1e-5 relative per vertex after one step, and a bounded, reported drift over 100 steps".
Runs config 1 (4,096 strands x 16) and 4,096 x 32 for 100 steps on the GPU in both arithmetic profiles and compares with
the CPU oracle (test infrastructure) after 1, 10, 50 and 100 steps: max / p99.9 / p99 / median relative position error
and the fraction of vertices within 1e-5 (error measure of SURVEY.md App. B). The exact profile must be bit-identical."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import barbu_b200 as bb
from oracle import pyoracle as po

DT = float(np.float32(1.0) / np.float32(90.0)); SPHERE = (0.0, 0.0, 0.0, 0.98)
def rel(a, b):
    a, b = a[:, :3].astype(np.float64), b[:, :3].astype(np.float64)
    return np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-30)
print(f"{'case':28s} {'steps':>5s} {'median':>9s} {'p99':>9s} {'p99.9':>9s} {'max':>9s} {'<=1e-5':>8s}  bit-identical")
for N, scale, warm in ((16, 1.0, 0), (16, 1.45, 60), (32, 1.0, 0), (32, 1.45, 60)):
    rows = cols = 64; S = rows * cols
    rp, rn, _ = po.sphere_scalp(rows, cols)
    pos, vel = po.init_strands(rp, rn, po.random_values(1234, S), N)
    par = po.default_params(dt=DT, scale=scale, sphere=SPHERE)
    for _ in range(warm): po.step(pos, vel, S, N, par, nthreads=16)
    for mname, mid in (("exact", bb.BH_MATH_EXACT), ("fast", bb.BH_MATH_FAST)):
        op, ov = pos.copy(), vel.copy()
        sim = bb.HairSim(S, N); sim.configure(scale=scale, sphere=SPHERE, math=mid); sim.upload(pos, vel)
        done = 0
        for upto in (1, 10, 50, 100):
            while done < upto:
                sim.step(DT, 1); po.step(op, ov, S, N, par, nthreads=16); done += 1
            gp, gv, _ = sim.download()
            e = rel(gp, op); e = e[np.isfinite(e)]
            same = np.array_equal(gp.view(np.uint32), op.view(np.uint32)) and np.array_equal(gv.view(np.uint32), ov.view(np.uint32))
            print(f"N={N:3d} sf={scale:4.2f} {'warm' if warm else 'cold'} {mname:5s}   {upto:5d} {np.median(e):9.2e} {np.quantile(e, .99):9.2e} "
                  f"{np.quantile(e, .999):9.2e} {e.max():9.2e} {100 * (e <= 1e-5).mean():7.3f}%  {same}")
        sim.close()
