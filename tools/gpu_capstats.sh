#!/bin/bash
# Capsule variant: GPU tests, then per-level pass rates and warp-time shares of the "arms" scene from the -DBH_STATS build.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
cp barbu_b200/lib/libbarbu_hair.so /tmp/prod.so; cp barbu_b200/lib/libbarbu_hair_stats.so barbu_b200/lib/libbarbu_hair.so
for m in exact fast; do for c in arms; do echo "stats $c $m"; timeout 300 python tests/reports/config3.py --caps $c --frames 1 --settle 30 --math $m --check 0 --log2s 20 2>&1 | grep BH_STATS | tail -1; done; done | tee gpurun_out/capstats.txt
cp /tmp/prod.so barbu_b200/lib/libbarbu_hair.so
