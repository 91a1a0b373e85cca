#!/usr/bin/env python3
"""Debug aid for a capsule fuzz seed that differs: steps device and oracle side by side and reports the first step at which
they differ, which vertices, and their distances to the colliders. Usage: python tests/reports/capsule_fuzz_debug.py <seed> ..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import barbu_b200 as bb
from oracle import pyoracle as po
import test_gpu_parity as t

for seed in [int(x) for x in sys.argv[1:]]:
    S, N, pos, vel, caps, sphere, wind, drag, scale = t._random_capsule_scene(seed)
    par, gcfg = t._capsule_params(caps, dt=float(t.DT), scale=scale, sphere=sphere)
    par.drag = gcfg.drag = drag
    for i, x in enumerate(wind):
        par.wind[i] = x; gcfg.wind[i] = x
    print(f"seed {seed}: S {S} N {N} caps {caps} sphere {sphere} wind {wind} drag {drag} scale {scale}")
    rp, rv = pos.copy(), vel.copy()
    with bb.HairSim(S, N) as sim:
        sim.set_params(gcfg); sim.upload(pos, vel)
        for step in range(40):
            prev = rp.copy()
            po.step(rp, rv, S, N, par, nthreads=16)
            sim.step(float(t.DT), 1)
            gp, gv, _ = sim.download()
            # bit-equal as tests/util.py defines it: a NaN may meet a NaN of another payload
            dp = ((gp.view(np.uint32) != rp.view(np.uint32)) & ~(np.isnan(gp) & np.isnan(rp))).any(axis=1)
            dv = ((gv.view(np.uint32) != rv.view(np.uint32)) & ~(np.isnan(gv) & np.isnan(rv))).any(axis=1)
            if dp.any() or dv.any():
                idx = np.nonzero(dp | dv)[0]
                print(f"  first difference at step {step}: {len(idx)} vertices; strands {sorted(set((idx // N).tolist()))[:10]}")
                for v in idx[:12]:
                    x = rp[v, :3].astype(np.float64); g = gp[v, :3].astype(np.float64)
                    d = []
                    for a, b, r in caps:
                        a, b = np.array(a), np.array(b); ab = b - a
                        tt = np.clip((x - a) @ ab / (ab @ ab), 0, 1) if ab @ ab > 0 else 0.0
                        tg = np.clip((g - a) @ ab / (ab @ ab), 0, 1) if ab @ ab > 0 else 0.0
                        d.append((round(float(np.linalg.norm(x - (a + tt * ab)) - r), 7), round(float(np.linalg.norm(g - (a + tg * ab)) - r), 7)))
                    ds = float(np.linalg.norm(x - np.array(sphere[:3])) - sphere[3])
                    print(f"    strand {v // N} vertex {v % N} lane {(v // N) % 32}: oracle {rp[v, :3]} device {gp[v, :3]} rest {rp[v, 3]:.5f} | oracle-dist to sphere {ds:.6f}, (oracle, device) dist to capsules {d} | pos differs {bool(dp[v])} vel differs {bool(dv[v])}")
                break
        else:
            print("  no difference in 40 steps")
