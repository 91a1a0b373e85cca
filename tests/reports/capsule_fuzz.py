#!/usr/bin/env python3
"""Fuzz of the capsule variant against the CPU oracle: random scenes (tests/test_gpu_parity.py::_random_capsule_scene), bit-exact
comparison after 40 steps. Usage: python tests/reports/capsule_fuzz.py [first_seed] [count] [fused_substeps] [capsules=1]  -> one line per seed + summary."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_parity import run_capsule_scene

first = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
count = int(sys.argv[2]) if len(sys.argv) > 2 else 200
fused = int(sys.argv[3]) if len(sys.argv) > 3 else 0      # k > 0: frames of k substeps as one fused launch
capsules = (int(sys.argv[4]) if len(sys.argv) > 4 else 1) != 0   # 0: the same generator without capsules, 1-70 vertices per strand
bad = 0
for seed in range(first, first + count):
    diff, kind, shape = run_capsule_scene(seed, nsteps=40 if fused != 3 else 39, fused_substeps=fused, capsules=capsules)
    if diff or (capsules and kind != 0):
        bad += 1
        print(f"seed {seed} S,N,caps={shape} kernel_kind={kind}: {diff} words differ", flush=True)
print(f"{'capsule' if capsules else 'sphere-only'} fuzz: seeds {first}..{first + count - 1}, {count} scenes x 40 steps{f' (fused frames of {fused} substeps)' if fused else ''}, exact profile vs oracle: {bad} scenes differ")
