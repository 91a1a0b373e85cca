#!/usr/bin/env python3
"""Executed instructions and stall samples per CUDA source line of the first kernel in an .ncu-rep (needs -lineinfo and
--import-source on): tools/ncu_lines.py rep.ncu-rep [top]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
agg = collections.defaultdict(lambda: [0, 0, ""])
fname = ""
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    if r[2] == "-" and r[0] != "":                      # a CUDA source line header row: counts already aggregated for the line
        try: agg[(fname, int(r[0]))] = [int(r[ie] or 0), int(r[isamp] or 0), r[1].strip()[:110]]
        except ValueError: pass
tot_e = sum(v[0] for v in agg.values()); tot_s = sum(v[1] for v in agg.values())
print(f"total executed {tot_e}, samples {tot_s}")
for (f, ln), (e, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * e / max(tot_e, 1):5.1f}% exec {100 * s / max(tot_s, 1):5.1f}% samp  {f}:{ln:<4d} {src}")
