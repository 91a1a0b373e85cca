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
ap.add_argument("--settle", type=int, default=12, help="untimed launches before the timed ones")
args = ap.parse_args()
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
# under torch.distributed.run the FIXED total of 2^log2v vertices is sharded over the ranks (strong scaling, contiguous
# strand ranges, no exchange); times are the max over ranks between two barriers
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    os.dup2(2, 1) if rank != 0 else None
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
V = 1 << args.log2v
DT = float(np.float32(1.0) / np.float32(90.0)) / 4
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
shapes = [(V // n, n) for n in (8, 16, 32, 64, 128)] + [(V // 4, 4)] + ([(4096, 16)] if world == 1 else [])
if rank == 0:
    print(f"# {world} GPU(s), {V} vertices in total, {args.settle} settle launches, {args.launches} timed launches; GB/s and frac are per GPU")
    print(f"{'strands':>11s} {'N':>4s} {'kernel':>7s} {'math':>6s} {'ms/launch':>10s} {'GB/s':>8s} {'frac':>6s} {'updates/s':>11s}", flush=True)
for Stot, N in shapes:
    rows = 1 << (int(np.log2(Stot)) // 2); cols = Stot // rows
    S = Stot // world; first = rank * S
    for mname, mid in (("exact", bb.BH_MATH_EXACT), ("fast", bb.BH_MATH_FAST)):
        sim = bb.HairSim(S, N, device=local)
        sim.set_stream(st.cuda_stream)
        sim.configure(scale=1.45, sphere=(0, 0, 0, 0.98), math=mid)
        sim.init_sphere_scalp(rows, cols, first, bb.random_values(1234, first, S))
        for _ in range(args.settle): sim.step(DT, 1)
        torch.cuda.synchronize()
        if dist is not None: dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(args.launches): sim.step(DT, 1)
        e1.record(st); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.launches], device="cuda", dtype=torch.float64)
        if dist is not None: dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        gbs = 64.0 * S * N / (ms * 1e-3) / 1e9
        if rank == 0:
            print(f"{Stot:11d} {N:4d} {sim.kernel_kind:7d} {mname:>6s} {ms:10.4f} {gbs:8.1f} {gbs / peak:6.3f} {Stot * N / (ms * 1e-3):11.3e}", flush=True)
        sim.close()
if dist is not None:
    dist.destroy_process_group()
