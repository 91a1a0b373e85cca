#!/usr/bin/env python3
"""configs[4] of BASELINE.json: sweep of vertices-per-strand 8..128 at a fixed vertex count, both arithmetic profiles,
plus the reference's own N = 4 and config 1 (4,096 x 16). One launch = one substep. Prints a table (and JSON lines).
Usage: python tools/sweep.py [--log2v 28] [--launches 20]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import barbu_b200 as bb

ap = argparse.ArgumentParser()
ap.add_argument("--log2v", type=int, default=28)
ap.add_argument("--launches", type=int, default=20)
args = ap.parse_args()
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
V = 1 << args.log2v
DT = float(np.float32(1.0) / np.float32(90.0)) / 4
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
shapes = [(V // n, n) for n in (8, 16, 32, 64, 128)] + [(V // 4, 4), (4096, 16)]
print(f"{'strands':>11s} {'N':>4s} {'kernel':>7s} {'math':>6s} {'ms/launch':>10s} {'GB/s':>8s} {'frac':>6s} {'updates/s':>11s}")
for S, N in shapes:
    rows = 1 << (int(np.log2(S)) // 2); cols = S // rows
    for mname, mid in (("exact", bb.BH_MATH_EXACT), ("fast", bb.BH_MATH_FAST)):
        sim = bb.HairSim(S, N)
        sim.set_stream(st.cuda_stream)
        sim.configure(scale=1.45, sphere=(0, 0, 0, 0.98), math=mid)
        sim.init_sphere_scalp(rows, cols, 0, bb.random_values(1234, 0, S))
        for _ in range(12): sim.step(DT, 1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(args.launches): sim.step(DT, 1)
        e1.record(st); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.launches
        gbs = 64.0 * S * N / (ms * 1e-3) / 1e9
        print(f"{S:11d} {N:4d} {sim.kernel_kind:7d} {mname:>6s} {ms:10.4f} {gbs:8.1f} {gbs / peak:6.3f} {S * N / (ms * 1e-3):11.3e}")
        sim.close()
