#!/usr/bin/env python3
"""Static instruction mix of the innermost loops of a kernel in a .so: tools/sass_mix.py <lib> <kernel-substring>
Prints each backward branch (loop) with its body size and opcode mix."""
import collections, re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs, cur = {}, None
for l in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", l)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, rows in funcs.items():
    if pat not in name:
        continue
    print("==", name[:140], len(rows), "instructions")
    for a, t in rows:
        m = re.search(r"BRA(?:\.U)?\s+(?:[!\w]+,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a and a - int(m.group(1), 16) > 0x400:
            lo = int(m.group(1), 16)
            c = collections.Counter()
            for b, u in rows:
                if lo <= b <= a:
                    op = u.split()
                    if op[0].startswith("@"): op = op[1:]
                    c[op[0].split(".")[0]] += 1
            print(f"  loop {lo:#x}..{a:#x}: {sum(c.values())} instr:", ", ".join(f"{k} {v}" for k, v in c.most_common(14)))
