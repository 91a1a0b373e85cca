#!/usr/bin/env python3
"""Host<->device copy ceiling of this box, for 1 rank or N concurrent ranks (what bh_step_host / bh_step_readback can reach at
best): pinned H2D alone, D2H alone and both at once, every rank on its own GPU starting together.

  python tools/pcie_probe.py                                                     # one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_probe.py [--affinity]

--affinity: pin the rank to the CPUs NVML names for its GPU (nvmlDeviceGetCpuAffinity) BEFORE the pinned buffers are
allocated, so that they are first touched — and thus placed — on the GPU's NUMA node. Rank 0 prints one JSON line."""
import argparse, json, os, time
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--affinity", action="store_true")
ap.add_argument("--mib", type=int, default=1024)
ap.add_argument("--chunks", type=int, default=8)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
cpus_before = sorted(os.sched_getaffinity(0))
numa = None
if args.affinity:
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, x in enumerate(words) for b in range(64) if (x >> b) & 1]
        cpus = [c for c in cpus if c in cpus_before] or cpus_before
        os.sched_setaffinity(0, cpus)
        try: numa = pynvml.nvmlDeviceGetNumaNodeId(h)
        except Exception: numa = None
    except Exception as exc:                                                 # report, never fail the probe
        numa = f"affinity unavailable: {exc}"
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = args.mib << 20
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_in.fill_(1); h_out.fill_(2)                                                # first touch here, under the affinity chosen above
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def run(h2d, d2h, reps=4):
    best = 1e9
    c = n // args.chunks
    for _ in range(reps):
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        t = time.perf_counter()
        for i in range(args.chunks):
            if h2d:
                with torch.cuda.stream(s1): d_in[i * c:(i + 1) * c].copy_(h_in[i * c:(i + 1) * c], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h_out[i * c:(i + 1) * c].copy_(d_out[i * c:(i + 1) * c], non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    return n / best / 1e9

res = torch.tensor([run(True, False), run(False, True), run(True, True)], device="cuda", dtype=torch.float64)
if world > 1:
    allr = [torch.zeros_like(res) for _ in range(world)]
    dist.all_gather(allr, res)
else:
    allr = [res]
if rank == 0:
    per = [[round(float(x), 2) for x in r.tolist()] for r in allr]
    line = {"ranks": world, "affinity": bool(args.affinity), "numa_node_rank0": numa, "cpus_rank0": len(os.sched_getaffinity(0)), "cpus_box": os.cpu_count(),
            "mib_per_direction": args.mib, "chunks": args.chunks,
            "per_rank_gbs": {"h2d_alone": [p[0] for p in per], "d2h_alone": [p[1] for p in per], "both_each_way": [p[2] for p in per]},
            "aggregate_gbs": {"h2d_alone": round(sum(p[0] for p in per), 1), "d2h_alone": round(sum(p[1] for p in per), 1),
                              "both_each_way": round(sum(p[2] for p in per), 1)},
            "note": "all ranks start each pass together (barrier); per-rank figure = 1 GiB / own wall time of the pass, best of 4"}
    print(json.dumps(line), flush=True)
if world > 1:
    dist.destroy_process_group()
