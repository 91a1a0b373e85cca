#!/usr/bin/env python3
"""configs[2] of BASELINE.json: "skinned creature mesh growing 4M strands x 32 vertices with capsule colliders".

The reference's glTF asset is a git-LFS pointer and its hair roots are never skinned (SURVEY.md §8 a-ext), so this is the
synthetic stand-in SURVEY §8d defines: sphere scalp 2048 x 2048, an 8-joint chain along y, a dual-quaternion palette
re-posed every frame, roots skinned on the device (bh_skin_roots), sphere + 2 capsule colliders, 4 substeps per frame.
EXTENSION CONFIG: no reference parity exists; the check below is against the CPU oracle on sampled strands.

Prints one JSON line per arithmetic profile. Usage: python tests/reports/config3.py [--log2s 22] [--frames 10] [--check 2048]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import barbu_b200 as bb

ap = argparse.ArgumentParser()
ap.add_argument("--log2s", type=int, default=22)
ap.add_argument("--frames", type=int, default=10)
ap.add_argument("--settle", type=int, default=30)
ap.add_argument("--check", type=int, default=2048, help="strands compared bit-for-bit with the CPU oracle (0: none)")
ap.add_argument("--caps", default="arms", choices=["arms", "far", "none"], help="far: capsules nothing can reach (cost of the bound test alone)")
ap.add_argument("--math", default="both", choices=["both", "exact", "fast"])
args = ap.parse_args()
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]

S, N, SUB, J = 1 << args.log2s, 32, 4, 8
rows = 1 << (args.log2s // 2); cols = S // rows
DT = float(np.float32(1.0) / np.float32(90.0))
SPHERE = (0.0, 0.0, 0.0, 0.98)
CAPS = [((0.55, 0.25, 0.0), (1.25, -0.35, 0.0), 0.30), ((-0.55, 0.25, 0.0), (-1.25, -0.35, 0.0), 0.30)]   # two "arms"
if args.caps == "none":
    CAPS = []
if args.caps == "far":
    CAPS = [((50.0, 0.0, 0.0), (51.0, 0.0, 0.0), 0.30), ((-50.0, 0.0, 0.0), (-51.0, 0.0, 0.0), 0.30)]


def scalp():
    """SURVEY §8d sphere scalp in fp32 (the same formulas the device generator uses)."""
    r = np.arange(rows, dtype=np.float32); c = np.arange(cols, dtype=np.float32)
    th = (np.float32(np.pi) * (r + np.float32(0.5)) / np.float32(rows) - np.float32(np.pi / 2)).astype(np.float32)
    ph = (np.float32(2 * np.pi) * c / np.float32(cols)).astype(np.float32)
    ct, st_ = np.cos(th).astype(np.float32), np.sin(th).astype(np.float32)
    n = np.stack([np.outer(ct, np.cos(ph).astype(np.float32)), np.repeat(st_[:, None], cols, 1), np.outer(ct, np.sin(ph).astype(np.float32))], -1)
    return n.reshape(-1, 3).astype(np.float32)


def skin_data(root):
    """4 nearest joints of a chain along y, inverse-distance weights (w.w = 1 - (x + y + z) on the device)."""
    jy = np.linspace(-1.0, 1.0, J, dtype=np.float32)
    d = np.abs(root[:, 1:2] - jy[None, :])
    idx = np.argsort(d, axis=1)[:, :4].astype(np.int32)
    w = 1.0 / (np.take_along_axis(d, idx, 1) + 0.05)
    w = (w / w.sum(1, keepdims=True)).astype(np.float32)
    return idx, np.ascontiguousarray(w[:, :3])


def palette(frame):
    """Unit dual quaternions [J][8] = (real xyzw, dual xyzw): joint j rotates about z and sways along x."""
    out = np.zeros((J, 8), np.float32)
    for j in range(J):
        a = 0.05 * np.sin(0.35 * frame + 0.7 * j)
        q = np.array([0.0, 0.0, np.sin(a / 2), np.cos(a / 2)])                       # xyzw
        t = np.array([0.02 * np.sin(0.2 * frame + j), 0.0, 0.0, 0.0])                 # translation as a pure quaternion
        # dual = 0.5 * t * q (quaternion product, xyzw layout)
        tx, ty, tz, _ = t; qx, qy, qz, qw = q
        dual = 0.5 * np.array([tx * qw + ty * qz - tz * qy, -tx * qz + ty * qw + tz * qx, tx * qy - ty * qx + tz * qw, -(tx * qx + ty * qy + tz * qz)])
        out[j, :4], out[j, 4:] = q, dual
    return out


def params(math):
    p = bb.default_params()
    p.scale, p.math, p.ncapsules = 1.45, math, len(CAPS)
    for i, x in enumerate(SPHERE): p.sphere[i] = x
    for q, (a, b, r) in enumerate(CAPS):
        for i in range(3): p.capsules[q].a[i], p.capsules[q].b[i] = a[i], b[i]
        p.capsules[q].radius = r
    return p


root = scalp()
joints, weights = skin_data(root)
rv = bb.random_values(1234, 0, S)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)

for mname, mid in (("exact", bb.BH_MATH_EXACT), ("fast", bb.BH_MATH_FAST)):
    if args.math not in ("both", mname):
        continue
    sim = bb.HairSim(S, N)
    sim.set_stream(st.cuda_stream)
    sim.set_params(params(mid))
    sim.init_sphere_scalp(rows, cols, 0, rv)
    sim.set_skin(root, joints, weights)
    frame = 0
    for _ in range(args.settle):
        sim.skin_roots(palette(frame)); sim.step(DT, SUB); frame += 1
    torch.cuda.synchronize()
    l0 = sim.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(args.frames):
        sim.skin_roots(palette(frame)); sim.step(DT, SUB); frame += 1
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.frames
    launches = sim.launch_count - l0
    # step-kernel time alone (skin_roots synchronises the stream for its pageable palette: time the launches separately)
    e0.record(st)
    for _ in range(args.frames): sim.step(DT, SUB)
    e1.record(st); torch.cuda.synchronize()
    ms_launch = e0.elapsed_time(e1) / (args.frames * SUB)
    gbs = 64.0 * S * N / (ms_launch * 1e-3) / 1e9
    line = {"workload": f"configs[2]: skinned sphere scalp {rows}x{cols}, {S} strands x {N}, {J}-joint DQ palette per frame, "
                        f"sphere + {len(CAPS)} capsules, {SUB} substeps/frame (extension config: oracle parity only)",
            "capsules": args.caps, "math": mname, "kernel_kind": sim.kernel_kind, "ms_per_frame": ms, "ms_per_launch": ms_launch,
            "updates_per_s_frame": S * N * SUB / (ms * 1e-3), "updates_per_s_step_kernel": S * N / (ms_launch * 1e-3),
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak},
            "gpu_launches_timed": launches}
    if args.check and mname == "exact":
        # strands are independent: replay `frames_chk` frames of a strided sample of strands on the CPU oracle
        from oracle import pyoracle as po
        sim2 = bb.HairSim(S, N); sim2.set_stream(st.cuda_stream); sim2.set_params(params(mid))
        sim2.init_sphere_scalp(rows, cols, 0, rv); sim2.set_skin(root, joints, weights)
        sel = np.linspace(0, S - 1, args.check).astype(np.int64)
        gp, gv, _ = sim2.download()
        gp, gv = gp.reshape(S, N, 4), gv.reshape(S, N, 4)
        rp, rvv = np.ascontiguousarray(gp[sel]).reshape(-1, 4), np.ascontiguousarray(gv[sel]).reshape(-1, 4)
        par = po.default_params(dt=float(np.float32(DT) / np.float32(SUB)), scale=1.45, sphere=SPHERE, ncapsules=len(CAPS))
        for q, (a, b, r) in enumerate(CAPS):
            for i in range(3): par.capsules[q].a[i], par.capsules[q].b[i] = a[i], b[i]
            par.capsules[q].radius = r
        frames_chk = 6
        for f in range(frames_chk):
            dq = palette(f)
            sim2.skin_roots(dq); sim2.step(DT, SUB)
            sp, _ = po.skin_roots_dq(root[sel], root[sel], joints[sel], weights[sel], dq)
            rp.reshape(-1, N, 4)[:, 0, :3] = sp
            for _ in range(SUB): po.step(rp, rvv, len(sel), N, par)
        gp, gv, _ = sim2.download()
        gp, gv = gp.reshape(S, N, 4)[sel].reshape(-1, 4), gv.reshape(S, N, 4)[sel].reshape(-1, 4)
        line["oracle_check"] = {"strands": int(len(sel)), "frames": frames_chk,
                                "positions_bit_equal": bool(np.array_equal(gp.view(np.uint32), rp.view(np.uint32))),
                                "velocities_bit_equal": bool(np.array_equal(gv.view(np.uint32), rvv.view(np.uint32))),
                                "moved_by_capsules": int((np.abs(gp[:, :3]).max() > 0))}
        sim2.close()
    print(json.dumps(line), flush=True)
    sim.close()
