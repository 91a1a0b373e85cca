#!/usr/bin/env python3
"""Static path lengths through the innermost step loop of a stream kernel: tools/sass_paths.py <lib> <kernel-substring> [min_body]
For every loop (backward branch) with a body of at least min_body instructions, walks the body from the loop head and
enumerates the paths through its forward conditional branches. A branch that only skips a block ending in CALL (the
out-of-line rare paths) is always taken; BRA.DIV falls through. Prints instruction count and opcode mix per path."""
import collections, re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
min_body = int(sys.argv[3]) if len(sys.argv) > 3 else 200
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
    print("==", name[:150], len(rows), "instructions")
    loops = []
    for i, (a, t) in enumerate(rows):
        m = re.search(r"BRA(?:\.U)?\s+(?:[!\w]+,\s*)?0x([0-9a-f]+)", t)
        if m and "DIV" not in t:
            tgt = int(m.group(1), 16)
            if tgt < a and (a - tgt) // 16 >= min_body and tgt in addr: loops.append((addr[tgt], i))
    # innermost: loops that contain no other loop
    inner = [lp for lp in loops if not any(o != lp and lp[0] <= o[0] and o[1] <= lp[1] for o in loops)]
    for lo, hi in inner:
        head = f"  loop {rows[lo][0]:#x}..{rows[hi][0]:#x}: {hi - lo + 1} static"
        paths = []
        def walk(i, cnt, tag):
            cnt = collections.Counter(cnt)
            while True:
                a, t = rows[i]
                cnt[opcode(t)] += 1
                if i == hi: paths.append((tag, cnt)); return
                m = re.search(r"BRA(?:\.U)?\s+(?:([!\w]+),\s*)?0x([0-9a-f]+)", t)
                if m and "DIV" not in t and opcode(t) == "BRA":
                    tgt = addr.get(int(m.group(2), 16))
                    cond = t.startswith("@")
                    if tgt is None or tgt <= i or tgt > hi + 1:
                        if not cond: paths.append((tag + " exit", cnt)); return
                        i += 1; continue
                    if not cond: i = tgt; continue
                    skipped = [opcode(rows[k][1]) for k in range(i + 1, tgt)]
                    if "CALL" in skipped and len(skipped) < 40: i = tgt; continue
                    walk(tgt, cnt, tag + f" T@{a:#x}")
                    tag += f" N@{a:#x}"
                i += 1
        walk(lo, collections.Counter(), "")
        paths.sort(key=lambda tc: sum(tc[1].values()))
        seen, shown = set(), []
        for tag, c in paths:                                   # one line per distinct (length, FFMA2 count): the step variants
            key = (sum(c.values()) // 4, c["FFMA2"], c["MUFU"])
            if key in seen: continue
            seen.add(key); shown.append((tag, c))
        if not shown or shown[0][1]["FFMA2"] + shown[0][1]["FFMA"] < 24: continue       # not a step loop (a cold block placed behind it)
        print(head + f"; {len(paths)} paths, {len(shown)} distinct")
        for tag, c in shown[:10]:
            n = sum(c.values())
            print(f"    {n:4d} instr [{tag.strip()[:60]}]: " + ", ".join(f"{k} {v}" for k, v in c.most_common(40)))
