#!/bin/bash
# Capsule variant, final check: full GPU suite, stats build (skipped tests that would have passed must be 0), configs[2] table, ncu captures of "arms"
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
cp barbu_b200/lib/libbarbu_hair.so /tmp/prod.so; cp barbu_b200/lib/libbarbu_hair_stats.so barbu_b200/lib/libbarbu_hair.so
for m in exact fast; do echo "stats arms $m"; timeout 300 python tests/reports/config3.py --caps arms --frames 1 --settle 30 --math $m --check 0 --log2s 20 2>&1 | grep BH_STATS | tail -1; done | tee gpurun_out/capstats_${TAG}.txt
cp /tmp/prod.so barbu_b200/lib/libbarbu_hair.so
for c in none far arms; do timeout 300 python tests/reports/config3.py --caps $c --frames 10 --settle 30 2>/dev/null; done > gpurun_out/config3_${TAG}.json
python - <<PY
import json
for l in open("gpurun_out/config3_${TAG}.json"):
    d = json.loads(l); print(d["capsules"], d["math"], "ms/launch %.3f frac %.3f" % (d["ms_per_launch"], d["roofline"]["frac"]), d.get("oracle_check", ""))
PY
for m in exact fast; do bash tools/gpu_ncu_caps.sh ${TAG} $m; done
