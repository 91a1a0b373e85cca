#!/usr/bin/env python3
"""Summarise an .ncu-rep (first captured kernel): key counters, stall reasons, instruction mix.
Usage: tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [--top N]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2]
get = lambda k: data[hdr.index(k)] if k in hdr else "n/a"
print("kernel:", get("Kernel Name")[:110])
V = 33554432
for k in ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "smsp__warps_eligible.avg.per_cycle_active", "sm__inst_executed.avg.per_cycle_elapsed",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
          "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg"]:
    print(f"  {k:75s} {get(k):>18s} {units[hdr.index(k)] if k in hdr else ''}")
try:
    print("  warp-instructions per vertex (config 2): %.1f" % (float(get("smsp__inst_executed.sum")) * 32 / V))
except ValueError:
    pass
items = [(h, float(data[i])) for i, h in enumerate(hdr) if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
tot = sum(v for _, v in items) or 1
print("stall samples:")
for h, v in sorted(items, key=lambda x: -x[1])[:9]:
    print(f"  {h.replace('smsp__pcsamp_warps_issue_stalled_', ''):28s} {100 * v / tot:5.1f}%")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout.splitlines()
idx = [i for i, l in enumerate(src) if l.startswith('"Address"')]
rows = list(csv.reader(src[idx[0]:(idx[1] - 1 if len(idx) > 1 else None)]))
h2, rows = rows[0], rows[1:]
ia, ie, isamp = h2.index("Source"), h2.index("Instructions Executed"), h2.index("# Samples")
mix, samp = collections.Counter(), collections.Counter()
for r in rows:
    op = r[ia].strip().split()
    if op[0].startswith("@"):
        op = op[1:]
    mn = op[0].split(".")[0]
    mix[mn] += int(r[ie]); samp[mn] += int(r[isamp])
tot = sum(mix.values())
print(f"instruction mix (executed {tot}, static {len(rows)}):")
for k, v in mix.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 18):
    print(f"  {k:10s} {100 * v / tot:5.1f}%  {v * 32 / V:7.1f}/vertex   stall samples {samp[k]}")
