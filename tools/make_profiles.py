#!/usr/bin/env python3
"""Assemble the tracked evidence under profiles/ from one evidence run in gpurun_out/ (tools/gpu_round.sh <tag>).
Usage: tools/make_profiles.py <tag> <round-name>      e.g.  tools/make_profiles.py r01b r01"""
import collections, csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rnd = sys.argv[1], sys.argv[2]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

def cp(src, dst):
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst)); return True
    return False

cp(f"bench_{tag}.json", f"{rnd}_bench.json")
cp(f"bench_reference_{tag}.json", f"{rnd}_bench_reference.json")
cp(f"launches_{tag}.csv", f"{rnd}_ncu_launches.csv")
cp("probe_box.txt", f"{rnd}_probe_box.txt")
cp(f"smi_{tag}.txt", f"{rnd}_smi.txt")
for f in ("ubench_f32x2.txt", "ubench_lat.txt", "isqrt_probe.txt", "occ_sweep.txt", "launch_probe_exact.txt", "launch_probe_fast.txt",
          "sweep_v28.txt"):   # drift_report.txt and bench_2gpu.json are copied by hand when a run produced them
    cp(f, f"{rnd}_{f}")

# ---- launch list: share of the step kernel in the bench command -------------------------------------------------
lf = os.path.join(G, f"launches_{tag}.csv")
if os.path.exists(lf):
    rows = [r for r in csv.reader(l for l in open(lf) if l.startswith('"'))]
    h = rows[0]; ik, iv, im = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
    tot = collections.defaultdict(float); cnt = collections.Counter()
    for r in rows[1:]:
        if r[im] == "gpu__time_duration.sum":
            k = r[ik].split("(")[0][-70:]
            ns = float(r[iv].replace(",", ""))
            if "hair_step" in k:      # the same kernel runs on the whole shard (timed region) and on 1/32 slices (bh_step_host leg)
                k += "  [whole shard, 2^20 strands]" if ns > 2e5 else "  [bh_step_host slice, 2^16 strands]"
            tot[k] += ns; cnt[k] += 1
    s = sum(tot.values())
    with open(os.path.join(P, f"{rnd}_ncu_launches_summary.txt"), "w") as f:
        f.write(f"ncu --metrics gpu__time_duration.sum --clock-control none -s 412 -c 1200: python bench.py --steps 2 --warmup 3 --preroll 0 --no-cpu-baseline\n")
        f.write("(the 412 skipped launches are the bench's 100 settle frames + 3 warm-up steps; listed: timed region of the headline profile,\n"
                " settle + timed region of the other profile, and the bh_step_host slices of the e2e leg)\n")
        f.write("(cold-cache, serialised launch times: compare SHARES, not absolutes)\n\n")
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            f.write(f"{100 * v / s:6.2f}%  {cnt[k]:4d} launches  {v / cnt[k] / 1e3:10.1f} us avg  {k}\n")

# ---- full captures ------------------------------------------------------------------------------------------------
traffic = {}
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg",
        "launch__shared_mem_per_block_dynamic"]
V = 1 << 25
for m in ("exact", "fast"):
    rep = os.path.join(G, f"prof_{m}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    h, u, d = raw[0], raw[1], raw[2]
    get = lambda k: d[h.index(k)] if k in h else "n/a"
    out = [f"ncu --set full --clock-control none --import-source on -k regex:hair_step_ -s 412 -c 1: python bench.py --steps 2 --warmup 3 --preroll 0 --math {m}  (settled state: launch 413)",
           f"kernel: {get('Kernel Name')}", ""]
    for k in KEYS:
        out.append(f"  {k:72s} {get(k):>18s} {u[h.index(k)] if k in h else ''}")
    def num(k):
        v = float(get(k).replace(",", "")); unit = u[h.index(k)]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(unit, 1)
    traffic[m] = int(num("dram__bytes_read.sum") + num("dram__bytes_write.sum"))
    out.append(f"  dram traffic per launch = read + write = {traffic[m]} B; algorithmic = 64 B x {V} vertices = {64 * V} B  (ratio {traffic[m] / (64 * V):.3f})")
    out.append(f"  warp-level instructions per vertex = smsp__inst_executed.sum * 32 / V = {float(get('smsp__inst_executed.sum').replace(',', '')) * 32 / V:.1f}")
    items = [(k, float(d[i])) for i, k in enumerate(h) if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k]
    tot = sum(v for _, v in items) or 1
    out += ["", "stall samples (smsp__pcsamp_warps_issue_stalled_*):"]
    out += [f"  {k.replace('smsp__pcsamp_warps_issue_stalled_', ''):24s} {100 * v / tot:5.1f}%" for k, v in sorted(items, key=lambda x: -x[1]) if v > 0]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout.splitlines()
    idx = [i for i, l in enumerate(src) if l.startswith('"Address"')]
    rows = list(csv.reader(src[idx[0]:(idx[1] - 1 if len(idx) > 1 else None)]))
    h2, rows = rows[0], rows[1:]
    ia, ie, isamp = h2.index("Source"), h2.index("Instructions Executed"), h2.index("# Samples")
    mix = collections.Counter()
    for r in rows:
        op = r[ia].strip().split()
        if op[0].startswith("@"): op = op[1:]
        mix[op[0].split(".")[0]] += int(r[ie])
    t = sum(mix.values())
    out += ["", f"executed SASS mix (warp-level, {t} instructions, {len(rows)} static):"]
    out += [f"  {k:10s} {100 * v / t:5.1f}%  {v * 32 / V:7.1f} per vertex" for k, v in mix.most_common(24)]
    blackwell = {k: v for k, v in mix.items() if k in ("FFMA2", "FADD2", "FMUL2", "UTMALDG", "UTMASTG", "SYNCS", "UBLKCP")}
    out += ["", "Blackwell-specific mnemonics executed: " + ", ".join(f"{k} {v}" for k, v in sorted(blackwell.items()))]
    open(os.path.join(P, f"{rnd}_ncu_{m}.txt"), "w").write("\n".join(out) + "\n")
    # hot-loop SASS with execution counts and stall samples
    with open(os.path.join(P, f"{rnd}_sass_{m}.txt"), "w") as f:
        f.write(f"# SASS of {get('Kernel Name')[:100]} (cuobjdump -sass equivalent from the ncu source page)\n# columns: index, warp-level executions, stall samples, instruction\n")
        for n, r in enumerate(rows):
            f.write(f"{n:5d} {int(r[ie]):10d} {int(r[isamp]):6d}  {r[ia].strip()}\n")
if traffic:
    json.dump(traffic, open(os.path.join(P, "traffic_bytes_per_launch.json"), "w"), indent=1)
print("profiles/:", sorted(os.listdir(P)))
