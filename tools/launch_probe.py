#!/usr/bin/env python3
"""Per-launch timing probe of the step kernel at configs[1]: back-to-back vs spaced launches (clock / power effects)."""
import os, sys, time, json, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import barbu_b200 as bb
math = bb.BH_MATH_FAST if "fast" in sys.argv else bb.BH_MATH_EXACT
S, N = 1 << 20, 32
sim = bb.HairSim(S, N)
sim.configure(scale=1.45, sphere=(0, 0, 0, 0.98), math=math)
sim.init_sphere_scalp(1024, 1024, 0, bb.random_values(1234, 0, S))
st = torch.cuda.Stream(); torch.cuda.set_stream(st); sim.set_stream(st.cuda_stream)
dt = 1.0 / 90.0 / 4
for _ in range(200): sim.step(dt, 1)
torch.cuda.synchronize()
def smi():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.active", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
def run(n, gap):
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in evs:
        a.record(st); sim.step(dt, 1); b.record(st)
        if gap: torch.cuda.synchronize(); time.sleep(gap)
    torch.cuda.synchronize()
    return np.array([a.elapsed_time(b) for a, b in evs])
print("idle smi:", smi())
t = run(400, 0); print("back-to-back per-launch ms: median %.4f min %.4f max %.4f first10 %s" % (np.median(t), t.min(), t.max(), np.round(t[:10], 3)), "| smi:", smi())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(400): sim.step(dt, 1)
e1.record(st); torch.cuda.synchronize(); print("400 launches bracketed: %.4f ms/launch" % (e0.elapsed_time(e1) / 400), "| smi:", smi())
t = run(60, 0.02); print("spaced 20 ms per-launch ms: median %.4f min %.4f max %.4f" % (np.median(t), t.min(), t.max()), "| smi:", smi())
t = run(400, 0); print("back-to-back again: median %.4f" % np.median(t))
