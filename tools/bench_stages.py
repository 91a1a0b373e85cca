#!/usr/bin/env python3
"""Device time and roofline of the stages either side of the step kernel (SURVEY.md §8f rows + generation kernels):
tess-stream, strand generation, patch indices, DQ-skinned roots, state checksum, Marschner LUTs.
Each line: algorithmic bytes (what the stage must read + write once), CUDA-event time on the sim's stream, GB/s and
the fraction of the measured HBM peak (MEASURED_PEAKS.json). Usage: python tools/bench_stages.py [--reps 10]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import barbu_b200 as bb

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=10)
args = ap.parse_args()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
st = torch.cuda.Stream(); torch.cuda.set_stream(st)


def timed(fn, reps=args.reps, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps): fn()
    e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(name, what, nbytes, ms, note=""):
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({"stage": name, "workload": what, "algorithmic_bytes": int(nbytes), "ms": round(ms, 4), "GB/s": round(gbs, 1),
                      "frac_of_measured_hbm_peak": round(gbs / peak, 3), "note": note}), flush=True)


# ---- strand generation at configs[1]: 2^20 strands x 32 -------------------------------------------------------------
rows, cols, N = 1024, 1024, 32
S = rows * cols; V = S * N
rv = bb.random_values(1234, 0, S)
sim = bb.HairSim(S, N); sim.set_stream(st.cuda_stream)
sim.configure(scale=1.45, sphere=(0, 0, 0, 0.98))
ms = timed(lambda: sim.init_sphere_scalp(rows, cols, 0, rv), reps=5)
report("init_sphere_scalp", f"{S} strands x {N}: roots + expand_strands (pos, vel planes) incl. the 4 MiB jitter upload", 32 * V + 4 * S, ms,
       "host call: includes one H2D of random_value[S] (pageable) and the stream synchronisation")
ms = timed(lambda: sim.checksum(3), reps=5)
report("state_checksum", f"pos + vel planes of {S} x {N}", 32 * V, ms, "host call: includes the 16-byte readback")

# skinned roots
root = np.zeros((S, 3), np.float32); root[:, 0] = 1.0
J = 8
joints = np.random.default_rng(0).integers(0, J, (S, 4)).astype(np.int32)
w = np.full((S, 3), 0.25, np.float32)
dq = np.zeros((J, 8), np.float32); dq[:, 3] = 1.0
sim.set_skin(root, joints, w)
ms = timed(lambda: sim.skin_roots(dq), reps=5)
report("skin_roots_dq", f"{S} roots, {J} joints", (12 + 16 + 12) * S + 12 * S, ms,
       "host call: palette upload + stream synchronisation; writes one 12-byte root per 512-byte strand row")
sim.close()

# ---- patch indices: sphere scalp 1024 x 1024, N = 4 (the reference's N) ... and tess-stream on a scalp that fits ---------
tri = bb.sphere_scalp_triangles(256, 256)
F = tri.reshape(-1, 3).shape[0]
for n in (4, 32):
    t0 = time.perf_counter(); out = bb.build_patch_indices(tri, n); dt_host = time.perf_counter() - t0
    report("build_patch_indices", f"{F} faces, N = {n}: {out.size} int32 (host call, includes H2D of the faces and D2H of the result)",
           12 * F + 4 * out.size, dt_host * 1e3, "PCIe-bound by construction: the result is returned to the host like the reference's std::vector")

rows, cols, N = 128, 128, 16
S = rows * cols
sim = bb.HairSim(S, N); sim.set_stream(st.cuda_stream)
sim.configure(scale=1.45, sphere=(0, 0, 0, 0.98))
sim.init_sphere_scalp(rows, cols, 0, bb.random_values(1234, 0, S))
root_nrm = np.zeros((S, 3), np.float32); root_nrm[:, 1] = 1.0
sim.upload(tan4=bb.init_tangents_host(root_nrm, S, 0, N))
for _ in range(40): sim.step(1.0 / 90.0, 1)
tri = bb.sphere_scalp_triangles(rows, cols)
patches = bb.build_patch_indices(tri, N)
sim.tess_set_patches(patches)
npatch = patches.size // 6
for (ni, nl, ns) in ((3, 2, 16), (1, 1, 4), (4, 2, 16), (4, 3, 16), (4, 4, 16), (6, 4, 32)):
    count = sim.tess_stream(ni, nl, ns, 7, download=False)
    ms = timed(lambda: sim.tess_stream(ni, nl, ns, 7, download=False))
    report("tess_stream", f"{npatch} patches ({rows}x{cols} scalp, N = {N}), ninstances {ni}, nlines {nl}, nsubsegments {ns}: {count} float4 out",
           16 * count + 2 * 16 * S * N + 4 * patches.size, ms, "write-bound: output stream + one pass over the pos/tan planes and the patch list")
sim.close()

# ---- Marschner LUTs ---------------------------------------------------------------------------------------------------
from barbu_b200 import marschner
m = marschner.Marschner(); m.init(); m.generate(); t0 = time.perf_counter(); m.generate(); dt_host = time.perf_counter() - t0
print(json.dumps({"stage": "marschner_luts", "workload": "2 x 128 x 128 RGBA16F (host call incl. readback)", "ms": round(dt_host * 1e3, 3),
                  "note": "runs only when a shading parameter changes (marschner.cc:35-69); latency, not bandwidth"}))
