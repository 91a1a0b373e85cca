#!/usr/bin/env python3
"""Launch time of the step kernel against shard size in the SAME settled state (N = 32): T(S) = a + b * S separates the
per-launch overhead (ramp-up + last-tile tail) from the streaming rate. Usage: python tools/size_scan.py [--settle 60]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import barbu_b200 as bb
ap = argparse.ArgumentParser(); ap.add_argument("--settle", type=int, default=60); ap.add_argument("--launches", type=int, default=40)
args = ap.parse_args()
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
DT = float(np.float32(1.0) / np.float32(90.0)); N = 32
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
for mname, mid in (("exact", bb.BH_MATH_EXACT), ("fast", bb.BH_MATH_FAST)):
    pts = []
    for log2s in (18, 19, 20, 21, 22, 23):
        S = 1 << log2s; rows = 1 << (log2s // 2); cols = S // rows
        sim = bb.HairSim(S, N); sim.set_stream(st.cuda_stream)
        sim.configure(scale=1.45, sphere=(0, 0, 0, 0.98), math=mid)
        sim.init_sphere_scalp(rows, cols, 0, bb.random_values(1234, 0, S))
        for _ in range(args.settle): sim.step(DT, 4)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(args.launches // 4): sim.step(DT, 4)
        e1.record(st); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (args.launches // 4 * 4)
        pts.append((S, ms))
        print(f"{mname} S=2^{log2s} ms/launch {ms:.4f} frac {64.0 * S * N / (ms * 1e-3) / 1e9 / peak:.3f}", flush=True)
        sim.close()
    x = np.array([p[0] for p in pts], float); y = np.array([p[1] for p in pts])
    b, a = np.polyfit(x, y, 1)
    print(f"{mname}: T(S) = {a * 1e3:.1f} us + {b * (1 << 20):.4f} ms per 2^20 strands  -> asymptotic frac {64.0 * (1 << 20) * N / (b * (1 << 20) * 1e-3) / 1e9 / peak:.3f}")
