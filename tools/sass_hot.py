#!/usr/bin/env python3
"""Print the SASS of the first kernel in an .ncu-rep with executed counts, stall samples and the dominant stall reasons.
Usage: tools/sass_hot.py rep.ncu-rep"""
import csv, subprocess, sys
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout.splitlines()
idx = [i for i, l in enumerate(src) if l.startswith('"Address"')]
rows = list(csv.reader(src[idx[0]:(idx[1] - 1 if len(idx) > 1 else None)]))
h, rows = rows[0], rows[1:]
ia, ie, isamp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
reasons = [(i, k[6:]) for i, k in enumerate(h) if k.startswith("stall_") and "Not Issued" not in k]
for n, r in enumerate(rows):
    rs = sorted(((int(r[i] or 0), k) for i, k in reasons), reverse=True)
    top = " ".join(f"{k}:{v}" for v, k in rs[:3] if v > 0)
    print(f"{n:5d} {int(r[ie]):10d} {int(r[isamp]):6d}  {r[ia].strip():70s} {top}")
