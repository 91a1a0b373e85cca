#!/usr/bin/env python3
"""bench.py — strand-vertex updates/s of the fused hair step on B200, with roofline and CPU baseline.

Contract:
  python bench.py --gpus N --steps K --warmup W          (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                    (the reference path on the host CPU cores)

A "step" is one frame of BASELINE.json configs[1]: 2^20 strands x 32 vertices per GPU, 4 substeps
(4 launches of the fused kernel with dt/4), sphere collider. Strands are independent, so ranks hold
disjoint strand ranges of one (1024 x 1024*N) scalp and never communicate (weak scaling).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "strand_vertex_updates_per_sec"
UNIT = "updates/s"
ROWS, COLS_PER_GPU, NVERTS, SUBSTEPS = 1024, 1024, 32, 4       # configs[1]: 1M strands x 32, 4 substeps/frame
DT = float(np.float32(1.0) / np.float32(90.0))                  # core/global_clock.cc:160-162
SPHERE = (0.0, 0.0, 0.0, 0.98)
SCALE = 1.45                                                    # reference default uScaleFactor
BYTES_PER_VERTEX_PER_LAUNCH = 64                                # float4 pos+vel read, float4 pos+vel written
SEED = 1234
COLUMN_MAJOR = True


CPU_SAMPLE_STRANDS = 1 << 17                                     # the CPU arm's bounded sample: 1/8 of the per-GPU workload per step
CPU_SAMPLE_STEPS, CPU_SAMPLE_WARMUP = 3, 1                       # cpu_baseline inside the GPU arm: same strands, same rule


def workload_config(args):
    """The `config` object of BOTH arms (the driver compares them): what is simulated, not how either arm ran it."""
    return {"workload": workload_name(args.gpus), "iterations": 8, "strand_order": args.order,
            "state": f"settled: {args.settle} untimed frames from the cold state before the warm-up steps (GPU arm); "
                     "the CPU arm steps its sample from the cold state after its own warm-up",
            "l2": f"state {32 * ROWS * COLS_PER_GPU * NVERTS / 2**30:g} GiB per GPU > 126 MB L2: every launch streams from HBM (no flush needed)",
            "timing": "GPU arm: CUDA events on the launching stream, max over ranks; CPU arm: wall clock around the sampled steps",
            "cpu_sample": cpu_sample_text(host_threads())}


def cpu_sample_text(threads):
    return (f"first {CPU_SAMPLE_STRANDS} of {ROWS * COLS_PER_GPU} strands x {NVERTS} vertices x {SUBSTEPS} substeps per step, "
            f"{threads} OpenMP threads, CPU restatement of the reference GLSL (oracle/, bit-exact to the shader source; no GL/llvmpipe in the image)")


def workload_name(n_gpus):
    tag = "configs[1]" if ROWS * COLS_PER_GPU == 1 << 20 else \
          ("configs[3] (64M strands over 8 GPUs) shard size" if ROWS * COLS_PER_GPU == 1 << 23 else "configs[1] at another shard size")
    return (f"{tag}: synthetic sphere scalp, {ROWS * COLS_PER_GPU} strands x {NVERTS} vertices per GPU, "
            f"{SUBSTEPS} substeps/frame, sphere collider r=0.98, scale {SCALE}, dt 1/90")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock, throttle reasons and power sampled during the timed region (B200_PROFILING.md): NVML, else nvidia-smi."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.other_mask = 0                        # every NVML clocks-event bit seen while sampling (NVML path only)
        self.t_begin = self.t_end = None

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def _nvml_loop(self, pynvml, handle):
        names = [("hw_slowdown", pynvml.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", pynvml.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", pynvml.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", pynvml.nvmlClocksEventReasonSwPowerCap)]
        smax = pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM)
        while not self._stop.is_set():
            try:
                sm = pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)
                mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(handle)
                pw = pynvml.nvmlDeviceGetPowerUsage(handle) / 1000.0
                self.rows.append((time.time(), [str(sm), str(smax)] + ["Active" if mask & bit else "Not Active" for _, bit in names] + [str(pw)]))
                self.other_mask |= mask
            except pynvml.NVMLError:
                pass
            self._stop.wait(0.002)

    def start(self):
        # NVML in-process (a sample every ~2 ms: the timed region of a short run still gets tens of samples); nvidia-smi as fallback
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            except Exception:
                pass
            handle = None
            if uuid:
                for cand in (uuid, "GPU-" + uuid):
                    try:
                        handle = pynvml.nvmlDeviceGetHandleByUUID(cand.encode() if hasattr(cand, "encode") else cand); break
                    except Exception:
                        handle = None
            if handle is None:
                handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)
            self._stop = threading.Event()
            self.proc = "nvml"
            self._thread = threading.Thread(target=self._nvml_loop, args=(pynvml, handle), daemon=True)
            self._thread.start()
            return
        except Exception:
            self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE, text=True,
                                         bufsize=1)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if self.proc == "nvml":
            self._stop.set()
            self._thread.join(timeout=2)
        else:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        def num(x):
            try:
                return float(x)
            except ValueError:
                return None
        good = [(t, r) for t, r in self.rows if len(r) >= 7 and num(r[0]) is not None]
        # samples whose arrival time falls inside the timed region (nvidia-smi prints one line per period)
        inside = [r for t, r in good if self.t_begin is not None and self.t_begin <= t <= self.t_end + 0.03]
        window = "timed region"
        if len(inside) < 2:            # region shorter than a few sampling periods: fall back to pre-roll + timed region
            inside, window = [r for _, r in good], "pre-roll + timed region"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower() == "active" for r in inside)]
        sm = [num(r[0]) for r in inside]
        pw = [num(r[6]) for r in inside if num(r[6]) is not None]
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(num(r[1]) for r in inside) if inside else None,
               "reasons": reasons, "samples": len(sm), "window": window, "power_w_max": max(pw) if pw else None}
        if self.proc == "nvml":
            # for the reader who sees a median below the maximum with an empty `reasons`: every event bit NVML reported at any
            # sample of this sampler's life (pre-roll included), by name
            try:
                import pynvml
                bits = {"gpu_idle": pynvml.nvmlClocksEventReasonGpuIdle, "applications_clocks_setting": pynvml.nvmlClocksEventReasonApplicationsClocksSetting,
                        "sw_power_cap": pynvml.nvmlClocksEventReasonSwPowerCap, "hw_slowdown": pynvml.nvmlClocksEventReasonHwSlowdown,
                        "sync_boost": pynvml.nvmlClocksEventReasonSyncBoost, "sw_thermal_slowdown": pynvml.nvmlClocksEventReasonSwThermalSlowdown,
                        "hw_thermal_slowdown": pynvml.nvmlClocksEventReasonHwThermalSlowdown, "hw_power_brake_slowdown": pynvml.nvmlClocksEventReasonHwPowerBrakeSlowdown,
                        "display_clock_setting": pynvml.nvmlClocksEventReasonDisplayClockSetting}
                out["event_bits_seen"] = sorted(k for k, b in bits.items() if self.other_mask & b)
            except Exception:
                pass
        return out


_emit_fd = None


def emit(text):
    """The one line of the bench contract, on the process's original stdout."""
    data = (text + "\n").encode()
    if _emit_fd is None:
        sys.stdout.write(text + "\n"); sys.stdout.flush()
    else:
        os.write(_emit_fd, data)


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_time(nstrands_sample, steps, warmup, threads):
    """Time the reference path on the host cores: the CPU oracle (a C restatement of cs_simulation.glsl pinned
    bit-exact to the reference shader source, oracle/_ref) on the first `nstrands_sample` strands of the workload."""
    from oracle import pyoracle as po
    root_pos, root_nrm, _ = po.sphere_scalp(ROWS, COLS_PER_GPU, column_major=COLUMN_MAJOR)   # same scalp, same strand order: its first strands
    root_pos, root_nrm = root_pos[:nstrands_sample], root_nrm[:nstrands_sample]
    rv = po.random_values(SEED, nstrands_sample)
    pos, vel = po.init_strands(root_pos, root_nrm, rv, NVERTS)
    h = float(np.float32(DT) / np.float32(SUBSTEPS))
    par = po.default_params(dt=h, scale=SCALE, sphere=SPHERE)
    for _ in range(warmup):
        for _ in range(SUBSTEPS):
            po.step(pos, vel, nstrands_sample, NVERTS, par, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        for _ in range(SUBSTEPS):
            po.step(pos, vel, nstrands_sample, NVERTS, par, nthreads=threads)
    dt = time.perf_counter() - t0
    return nstrands_sample * NVERTS * SUBSTEPS * steps / dt, dt / steps


def reference_source_time(nstrands_sample):
    """The reference's own shader SOURCE (oracle/_ref: cs_simulation.glsl compiled as C++ over the reference's GLM, one
    fiber per invocation, OpenMP over workgroups) on a small sample: one frame of 4 substeps. Reported beside the
    restatement; it is NOT the arm's value — its time is dominated by the fiber switches that emulate the workgroup
    barriers (~100x slower than the restatement, which computes the same bits), so using it would inflate every ratio."""
    from oracle import pyoracle as po
    if not po.ref_available(NVERTS):
        return None
    root_pos, root_nrm, _ = po.sphere_scalp(ROWS, COLS_PER_GPU, column_major=COLUMN_MAJOR)
    root_pos, root_nrm = root_pos[:nstrands_sample], root_nrm[:nstrands_sample]
    pos, vel = po.init_strands(root_pos, root_nrm, po.random_values(SEED, nstrands_sample), NVERTS)
    h = float(np.float32(DT) / np.float32(SUBSTEPS))
    po.ref_update(pos, vel, nstrands_sample, NVERTS, h, SCALE, SPHERE)   # warm-up (fiber stacks, page faults)
    t0 = time.perf_counter()
    for _ in range(SUBSTEPS):
        po.ref_update(pos, vel, nstrands_sample, NVERTS, h, SCALE, SPHERE)
    return nstrands_sample * NVERTS * SUBSTEPS / (time.perf_counter() - t0)


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = host_threads()
    sample = CPU_SAMPLE_STRANDS
    value, s_per_step = cpu_reference_time(sample, args.steps, args.warmup, threads)
    src_sample = 1 << 12
    src_value = reference_source_time(src_sample)
    sample_txt = cpu_sample_text(threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample_txt},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "reference_source": {"value": src_value, "unit": UNIT, "cores": threads, "kind": "reference",
                             "sample": f"first {src_sample} strands x {NVERTS} vertices x {SUBSTEPS} substeps, 1 step",
                             "note": "oracle/_ref: the reference shader source itself on the CPU (fiber per invocation); "
                                     "bit-identical results to the restatement timed above, reported for completeness"},
    }
    emit(json.dumps(line))


def run_gpu(args, rank, world, local_rank):
    import torch
    import barbu_b200 as bb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hair simulation has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"           # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    S = ROWS * COLS_PER_GPU                       # strands of this rank
    # one scalp grid ROWS x cols_total shared by all ranks; rank g owns the contiguous strand range [g*S, (g+1)*S): with
    # column-major strand order a longitude wedge (every rank the same contact load), with row-major a latitude band.
    # --emulate-rank/--emulate-world: this single GPU plays rank R of a W-GPU job (straggler diagnosis on one GPU).
    eworld, erank = (args.emulate_world, args.emulate_rank) if args.emulate_world > 0 else (world, rank)
    cols_total = COLS_PER_GPU * eworld
    first = erank * S
    order = bb.BH_SCALP_COLUMN_MAJOR if COLUMN_MAJOR else bb.BH_SCALP_ROW_MAJOR
    V = S * NVERTS
    math = bb.BH_MATH_FAST if args.math == "fast" else bb.BH_MATH_EXACT

    sim = bb.HairSim(S, NVERTS, device=local_rank)
    sim.configure(scale=SCALE, sphere=SPHERE, math=math)
    stream = torch.cuda.Stream()                  # a real (non-default) stream owned by torch ...
    torch.cuda.set_stream(stream)
    sim.set_stream(stream.cuda_stream)            # ... that our kernels are launched on, so torch events bracket them

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_region(math_id, sampler, fused=False):
        """W untimed steps (+ pre-roll so clocks settle), then exactly K steps between two CUDA events on the launching
        stream, barrier + synchronize on both sides, max over ranks. Returns (ms_total, launches).
        fused: the 4 substeps of a frame as 4 passes of ONE launch (bh_set_substep_fusion) — reported as value_fused."""
        sim.configure(math=math_id)
        sim.set_substep_fusion(fused or args.fused_main)
        sim.init_sphere_scalp(ROWS, cols_total, first, rv, order=order)   # every profile starts from the same cold state
        for i in range(args.settle):
            sim.step(DT, SUBSTEPS)
            if i % 16 == 15:
                torch.cuda.synchronize()
        t_w = time.perf_counter()
        done = 0
        while done < args.warmup or time.perf_counter() - t_w < args.preroll:
            sim.step(DT, SUBSTEPS)
            done += 1
            if done % 16 == 0:
                torch.cuda.synchronize()
        barrier()
        if sampler is not None:
            sampler.mark_begin()
        l0 = sim.launch_count
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(args.steps):
            sim.step(DT, SUBSTEPS)
        ev1.record(stream)
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda", dtype=torch.float64)
        per_rank = [float(ms.item())]
        if dist is not None:
            allms = [torch.zeros_like(ms) for _ in range(world)]
            dist.all_gather(allms, ms)
            per_rank = [float(x.item()) for x in allms]
        per_rank_ms.append(per_rank)
        launches = sim.launch_count - l0
        # the same loop again for >= args.sustain seconds: K = 20 frames is 28 ms, shorter than a power-management period
        sustained = None
        if args.sustain > 0:
            n = max(args.steps, int(args.sustain / max(max(per_rank) * 1e-3 / args.steps, 1e-6)) + 1)
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            s0.record(stream)
            for _ in range(n):
                sim.step(DT, SUBSTEPS)
            s1.record(stream)
            barrier()
            sms = torch.tensor([s0.elapsed_time(s1)], device="cuda", dtype=torch.float64)
            if dist is not None:
                dist.all_reduce(sms, op=dist.ReduceOp.MAX)
            sustained = {"steps": n, "seconds": float(sms.item()) * 1e-3, "ms_per_step": float(sms.item()) / n,
                         "value": world * V * SUBSTEPS * n / (float(sms.item()) * 1e-3), "unit": UNIT}
        if sampler is not None:
            sampler.mark_end()
        sustained_by_math[(math_id, fused)] = sustained
        return max(per_rank), launches

    # ---- device-resident timing: value + roofline ---------------------------------------------------
    per_rank_ms = []
    sustained_by_math = {}
    rv = bb.random_values(SEED, first, S)
    kernel_kind = sim.kernel_kind
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total, launches = timed_region(math, sampler if rank == 0 else None)
    clocks = sampler.stop() if rank == 0 else None
    # the other arithmetic profile, same workload and timing rules (reported beside the headline, never instead of it)
    other = None
    other_math = bb.BH_MATH_EXACT if math == bb.BH_MATH_FAST else bb.BH_MATH_FAST
    if not args.no_other_profile:
        sampler2 = ClockSampler(local_rank)
        if rank == 0:
            sampler2.start()
        ms2, launches2 = timed_region(other_math, sampler2 if rank == 0 else None)
        clocks2 = sampler2.stop() if rank == 0 else None
        other = (ms2, launches2, clocks2)
        sim.configure(math=math)

    # ---- the same frames with the substeps fused into one launch per frame (tiles stay in L2 between substeps) -----------
    fused = {}
    if not args.no_fused:
        for m_id, name in ((math, args.math),) + (() if args.no_other_profile else ((other_math, "exact" if args.math == "fast" else "fast"),)):
            msf, lf = timed_region(m_id, None, fused=True)
            fused[name] = (msf, lf, sustained_by_math.get((m_id, True)))
        sim.set_substep_fusion(False)
        sim.configure(math=math)

    # ---- end to end through the C ABI with HOST buffers (bh_step_host): H2D + 4 substeps + D2H per step ---
    e2e_value, e2e_resident, e2e_steps, e2e_copy_gbs, host_ceiling_gbs = None, None, 0, None, None
    if not args.no_e2e:
        hp, hv = bb.PinnedBuffer(4 * V), bb.PinnedBuffer(4 * V)
        p0, v0, _ = sim.download()
        hp.array[:] = p0.reshape(-1)
        hv.array[:] = v0.reshape(-1)
        del p0, v0
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            sim.step_host(DT, SUBSTEPS, hp.array, hv.array)      # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            sim.step_host(DT, SUBSTEPS, hp.array, hv.array)      # returns after the D2H copies completed
        torch.cuda.synchronize()
        t_e2e = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        e2e_value = world * V * SUBSTEPS * e2e_steps / float(t_e2e.item())
        # Second flavour, reported beside it: the reference's own call pattern. Hair::update(dt) takes no buffers — the
        # state lives in the module's GL buffer — so per frame the host sends the uniforms (bounding sphere, dt) and a
        # host-side consumer reads the position plane back.
        sim.upload(hp.array, hv.array)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            sim.set_bounding_sphere(SPHERE)
            sim.step_readback(DT, SUBSTEPS, hp.array)           # returns after the position plane arrived
        torch.cuda.synchronize()
        t_res = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t_res, op=dist.ReduceOp.MAX)
        e2e_resident = world * V * SUBSTEPS * e2e_steps / float(t_res.item())
        # What the box lets through at best: every rank copies its planes host->device and device->host at once, no kernel —
        # the same pinned buffers, all ranks together. The e2e figure above can at most reach bytes / this rate
        # (tools/pcie_probe.py, profiles/r02_pcie_probe.txt: the host side of these boxes saturates near 50 GB/s each way
        # however many GPUs copy, so e2e stops scaling with N while the device-timed `value` does not).
        dplane = torch.empty(8 * V, dtype=torch.float32, device="cuda")                # pos + vel: 32 B per vertex
        hin = torch.from_numpy(hp.array); hout = torch.from_numpy(hv.array)
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
        best = None
        for _ in range(3):
            torch.cuda.synchronize(); barrier()
            t0 = time.perf_counter()
            with torch.cuda.stream(s_up):
                dplane[:4 * V].copy_(hin, non_blocking=True); dplane[:4 * V].copy_(hin, non_blocking=True)
            with torch.cuda.stream(s_dn):
                hout.copy_(dplane[4 * V:], non_blocking=True); hout.copy_(dplane[4 * V:], non_blocking=True)
            torch.cuda.synchronize()
            t_c = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
            if dist is not None:
                dist.all_reduce(t_c, op=dist.ReduceOp.MAX)
            best = float(t_c.item()) if best is None else min(best, float(t_c.item()))
        host_ceiling_gbs = world * 32.0 * V / best / 1e9                               # aggregate, each way, both ways busy
        e2e_copy_gbs = world * 32.0 * V * e2e_steps / float(t_e2e.item()) / 1e9
        del dplane
        hp.free(); hv.free()

    # ---- optional exchange step (N > 1): all-gather of the position plane to every rank, NCCL over NVLink ----------
    gather = None
    if dist is not None and args.allgather:
        from barbu_b200 import shard
        local = shard.plane_tensor(sim, 0)
        counts = [V] * world
        sim.synchronize()
        for _ in range(2):
            full = shard.allgather_plane(local, counts)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        reps = 5
        for _ in range(reps):
            full = shard.allgather_plane(local, counts)
        g1.record()
        barrier()
        gms = torch.tensor([g0.elapsed_time(g1) / reps], device="cuda", dtype=torch.float64)
        dist.all_reduce(gms, op=dist.ReduceOp.MAX)
        recv = 16 * V * (world - 1)
        gather = {"ms": float(gms.item()), "bytes_received_per_gpu": recv, "gbs_per_gpu": recv / (float(gms.item()) * 1e-3) / 1e9,
                  "what": "ncclAllGather of the float4 position plane (optional, off the per-step path)"}
        del full
    sim.close()

    if rank == 0:
        peak, peak_src = peaks()
        per_launch_s = ms_total * 1e-3 / launches
        achieved = BYTES_PER_VERTEX_PER_LAUNCH * V / per_launch_s / 1e9
        value = world * V * SUBSTEPS * args.steps / (ms_total * 1e-3)
        threads = host_threads()
        # the CPU arm's own sampling rule (bench.py --impl reference): same strands, same warm-up, same steps
        cpu_value, _ = cpu_reference_time(CPU_SAMPLE_STRANDS, args.steps, args.warmup, threads) if world == 1 and not args.no_cpu_baseline else (None, None)
        kernel_names = {0: "hair_step_stream_kernel", 1: "hair_step_pipelined_kernel", 2: "hair_step_generic_kernel"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(args), "math": args.math,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "kernel": kernel_names.get(kernel_kind, "?"), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": BYTES_PER_VERTEX_PER_LAUNCH * V, "ms_per_launch": per_launch_s * 1e3},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 32 * V, "d2h_bytes_per_step": 32 * V,
                    "steps": e2e_steps, "api": "bh_step_host (pinned host pos+vel planes in and out every step)",
                    "copy_gbs_each_way": e2e_copy_gbs, "host_ceiling_gbs": host_ceiling_gbs,
                    "host_ceiling_note": "aggregate over all ranks, each way with both ways busy, same pinned buffers, no kernel, measured in this run"},
            "e2e_resident_state": {"value": e2e_resident, "unit": UNIT, "h2d_bytes_per_step": 16, "d2h_bytes_per_step": 16 * V,
                                   "steps": e2e_steps, "api": "bh_set_bounding_sphere + bh_step_readback(position plane): the reference's Hair::update "
                                   "call pattern, state resident on the device, positions read back to pinned host memory slice by slice behind the steps"},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if world > 1:
            line["ms_per_step_by_rank"] = [t / args.steps for t in per_rank_ms[0]]
        if cpu_value is not None:
            line["cpu_baseline"] = {"value": cpu_value, "unit": UNIT, "cores": threads, "kind": "port", "sample": cpu_sample_text(threads),
                                    "steps": args.steps, "warmup": args.warmup}
        traffic = {}
        traffic_file = os.path.join(ROOT, "profiles", "traffic_bytes_per_launch.json")
        if os.path.exists(traffic_file):
            with open(traffic_file) as f:
                traffic = json.load(f)
        if S != 1 << 20:
            traffic = {}                                      # the ncu capture is of the configs[1] launch
        line["roofline"]["traffic"] = traffic.get(args.math)
        line["roofline"]["traffic_source"] = ("profiles/traffic_bytes_per_launch.json (dram__bytes_read.sum + dram__bytes_write.sum of one "
                                              "`ncu --set full` capture of this launch shape; static, not measured in this run)") if traffic.get(args.math) else None
        if sustained_by_math.get((math, False)) is not None:
            line["sustained"] = dict(sustained_by_math[(math, False)], note="the timed loop continued for >= --sustain seconds, same events and rules")
        if args.emulate_world > 0:
            line["emulated_shard"] = {"rank": erank, "world": eworld, "order": args.order}
        if other is not None:
            ms2, launches2, clocks2 = other
            name2 = "exact" if args.math == "fast" else "fast"
            per2 = ms2 * 1e-3 / launches2
            ach2 = BYTES_PER_VERTEX_PER_LAUNCH * V / per2 / 1e9
            line["other_profile"] = {"math": name2, "value": world * V * SUBSTEPS * args.steps / (ms2 * 1e-3), "unit": UNIT,
                                     "ms_per_step": ms2 / args.steps, "gpu_launches": launches2, "clocks": clocks2,
                                     "sustained": sustained_by_math.get((other_math, False)),
                                     "roofline": {"bound": "hbm", "achieved": ach2, "peak": peak, "unit": "GB/s", "frac": ach2 / peak,
                                                  "traffic": traffic.get(name2), "ms_per_launch": per2 * 1e3},
                                     "note": "same workload, steps and timing rules; exact = bit-identical to the CPU oracle, "
                                             "fast = FMA-contracted + MUFU.RSQ arithmetic (<= 1e-5 relative per vertex after one step)"}
        def fused_obj(name):
            msf, lf, sus = fused[name]
            return {"value": world * V * SUBSTEPS * args.steps / (msf * 1e-3), "unit": UNIT, "ms_per_step": msf / args.steps, "gpu_launches": lf,
                    "substeps_per_launch": SUBSTEPS, "hbm_algorithmic_bytes_per_launch": BYTES_PER_VERTEX_PER_LAUNCH * V,
                    "hbm_achieved_gbs_per_launch": BYTES_PER_VERTEX_PER_LAUNCH * V / (msf * 1e-3 / max(lf, 1)) / 1e9,
                    "sustained": sus,
                    "note": "bh_set_substep_fusion(1): the 4 substeps of a frame as 4 passes of ONE launch, a warp re-reading from L2 what its "
                            "previous pass stored; bit-identical to 4 launches (tests). HBM traffic is 64 B per vertex per LAUNCH, so this "
                            "figure is updates/s only — `roofline` above stays the unfused, HBM-bound launch (SURVEY.md 8d accounting)"}
        if args.math in fused:
            line["value_fused"] = fused_obj(args.math)
        if other is not None and line["other_profile"]["math"] in fused:
            line["other_profile"]["value_fused"] = fused_obj(line["other_profile"]["math"])
        if gather is not None:
            line["allgather"] = gather
        emit(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def run_group(args):
    """--single-process: the same weak-scaling job through the C ABI's group API (bh_group_*): ONE process, ONE host thread,
    args.gpus devices — the form the reference's single-threaded frame loop can call. Timed with per-device CUDA events
    (bh_group_step_timed), max over shards; the optional gather of the position plane to device 0 is timed separately."""
    import torch
    import barbu_b200 as bb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hair simulation has no CPU path")
    G = args.gpus
    devices = list(range(G)) if torch.cuda.device_count() >= G else [i % torch.cuda.device_count() for i in range(G)]
    S_total = ROWS * COLS_PER_GPU * G
    V_shard = ROWS * COLS_PER_GPU * NVERTS
    math = bb.BH_MATH_FAST if args.math == "fast" else bb.BH_MATH_EXACT
    grp = bb.HairGroup(devices, S_total, NVERTS)
    grp.configure(scale=SCALE, sphere=SPHERE, math=math)
    grp.init_sphere_scalp(ROWS, COLS_PER_GPU * G, order=bb.BH_SCALP_COLUMN_MAJOR if COLUMN_MAJOR else bb.BH_SCALP_ROW_MAJOR, seed=SEED)
    for i in range(args.settle + args.warmup):
        grp.step(DT, SUBSTEPS)
        if i % 16 == 15:
            grp.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    sampler.mark_begin()
    l0 = grp.launch_count
    ms_total, per = grp.step_timed(DT, SUBSTEPS, args.steps)
    launches = grp.launch_count - l0
    sampler.mark_end()
    clocks = sampler.stop()
    gather = None
    if G > 1:
        grp.gather_plane(0, devices[0])
        gms = min(grp.gather_plane(0, devices[0])[1] for _ in range(5))
        recv = 16 * V_shard * (G - 1) if len(set(devices)) > 1 else 0
        gather = {"ms": gms, "bytes_received_by_render_gpu": recv, "gbs": recv / (gms * 1e-3) / 1e9 if gms > 0 else None,
                  "nvlink_reference_gbs": 770.0, "what": "bh_group_gather_plane: position plane of every shard pushed to device 0 by "
                  "cudaMemcpyPeerAsync over NVLink (optional, off the per-step path); 770 GB/s = measured per-direction NVLink figure of SURVEY.md §5"}
    grp.close()
    peak, peak_src = peaks()
    per_launch_s = ms_total * 1e-3 / (launches / G)
    achieved = BYTES_PER_VERTEX_PER_LAUNCH * V_shard / per_launch_s / 1e9
    line = {"metric": METRIC, "value": G * V_shard * SUBSTEPS * args.steps / (ms_total * 1e-3), "unit": UNIT, "n_gpus": G, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args), "math": args.math,
            "mode": "single process, one host thread, bh_group_* (C ABI)", "devices": devices,
            "ms_per_step_by_shard": [t / args.steps for t in per],
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "kernel": "hair_step_stream_kernel", "peak_source": peak_src, "per": "GPU (slowest shard)",
                         "algorithmic_bytes_per_launch": BYTES_PER_VERTEX_PER_LAUNCH * V_shard, "ms_per_launch": per_launch_s * 1e3},
            "gpu_launches": launches, "clocks": clocks}
    if gather:
        line["gather"] = gather
    emit(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--math", default="fast", choices=["exact", "fast"],
                    help="arithmetic profile of the headline numbers (the other one is timed too and reported in other_profile): "
                         "fast = FMA contraction + MUFU.RSQ, what a GPU GLSL compiler emits for the reference shader, inside "
                         "north_star's tolerance (<= 1e-5 relative per vertex after one step); exact = bit-identical to the CPU oracle")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--no-other-profile", action="store_true", help="time only the --math profile")
    ap.add_argument("--fused-main", action="store_true", help="profiling runs: the main timed region itself uses fused substeps")
    ap.add_argument("--no-fused", action="store_true", help="skip the fused-substeps legs (value_fused)")
    ap.add_argument("--allgather", action="store_true", help="N > 1: also time the optional NCCL all-gather of the position plane")
    ap.add_argument("--preroll", type=float, default=0.0, help="minimum seconds of untimed warm-up (0 for profiler runs)")
    ap.add_argument("--settle", type=int, default=100,
                    help="untimed frames run before the W warm-up steps so that the timed region sees the SETTLED hair (strands "
                         "draped over the collider, push-outs in ~40%% of warp-steps), not the cold straight state, which is "
                         "cheaper for the exact profile (0 for profiler runs that want the cold state)")
    ap.add_argument("--single-process", action="store_true",
                    help="drive all --gpus devices from this one process and host thread through the C ABI's group API (bh_group_*) "
                         "instead of one rank per GPU under torch.distributed.run")
    ap.add_argument("--order", default="column", choices=["column", "row"],
                    help="strand order of the sphere scalp: column = meridian by meridian (contiguous shards are balanced longitude "
                         "wedges), row = latitude circle by circle (round 1: shard = latitude band, the polar band straggles)")
    ap.add_argument("--sustain", type=float, default=0.25, help="seconds the timed loop is continued for the `sustained` figure (0: off)")
    ap.add_argument("--emulate-world", type=int, default=0, help="with --emulate-rank: run that rank's shard of a W-GPU job on this one GPU")
    ap.add_argument("--emulate-rank", type=int, default=0)
    ap.add_argument("--log2-strands", type=int, default=20,
                    help="strands per GPU = 2^k (default 20 = configs[1]; 23 = the per-GPU shard of configs[3], 64M strands over 8 GPUs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    global ROWS, COLS_PER_GPU, COLUMN_MAJOR
    COLUMN_MAJOR = args.order == "column"
    ROWS = 1 << (args.log2_strands // 2)
    COLS_PER_GPU = (1 << args.log2_strands) // ROWS
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE line (the JSON): libraries that print there on their own (NCCL's version banner does)
    # are sent to stderr for the duration of the run; emit() writes to the real stdout.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    global _emit_fd
    _emit_fd = real_stdout
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.single_process:
        run_group(args)
    else:
        run_gpu(args, rank, world, local_rank)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
