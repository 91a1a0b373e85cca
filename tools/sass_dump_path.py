#!/usr/bin/env python3
"""Print the instructions along one path of the innermost step loop (see sass_paths.py) — the path whose opcode counts match
FFMA2=<n> and, optionally, MUFU=<m>, shortest first. Usage: tools/sass_dump_path.py <lib> <kernel-substring> <ffma2> [mufu]"""
import collections, re, subprocess, sys
lib, pat, want = sys.argv[1], sys.argv[2], int(sys.argv[3])
want_mufu = int(sys.argv[4]) if len(sys.argv) > 4 else None
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs, cur = {}, None
for l in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", l)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
def opcode(t):
    p = t.split()
    if p[0].startswith("@"): p = p[1:]
    return p[0].split(".")[0]
for name, rows in funcs.items():
    if pat not in name: continue
    addr = {a: i for i, (a, _) in enumerate(rows)}
    loops = []
    for i, (a, t) in enumerate(rows):
        m = re.search(r"BRA(?:\.U)?\s+(?:[!\w]+,\s*)?0x([0-9a-f]+)", t)
        if m and "DIV" not in t:
            tgt = int(m.group(1), 16)
            if tgt < a and (a - tgt) // 16 >= 300 and tgt in addr: loops.append((addr[tgt], i))
    inner = [lp for lp in loops if not any(o != lp and lp[0] <= o[0] and o[1] <= lp[1] for o in loops)]
    best = None
    for lo, hi in inner:
        paths = []
        def walk(i, seq):
            seq = list(seq)
            while True:
                a, t = rows[i]
                seq.append(i)
                if i == hi: paths.append(seq); return
                m = re.search(r"BRA(?:\.U)?\s+(?:([!\w]+),\s*)?0x([0-9a-f]+)", t)
                if m and "DIV" not in t and opcode(t) == "BRA":
                    tgt = addr.get(int(m.group(2), 16))
                    cond = t.startswith("@")
                    if tgt is None or tgt <= i or tgt > hi + 1:
                        if not cond: return
                        i += 1; continue
                    if not cond: i = tgt; continue
                    skipped = [opcode(rows[k][1]) for k in range(i + 1, tgt)]
                    if "CALL" in skipped and len(skipped) < 40: i = tgt; continue
                    walk(tgt, seq)
                i += 1
        walk(lo, [])
        for seq in paths:
            c = collections.Counter(opcode(rows[k][1]) for k in seq)
            if c["FFMA2"] == want and (want_mufu is None or c["MUFU"] == want_mufu):
                if best is None or len(seq) < len(best): best = seq
    if best:
        c = collections.Counter(opcode(rows[k][1]) for k in best)
        print("==", len(best), "instr:", ", ".join(f"{k} {v}" for k, v in c.most_common(40)))
        for k in best: print(f"{rows[k][0]:#07x}  {rows[k][1]}")
