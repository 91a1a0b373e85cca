#!/bin/bash
# Capsule variant on the GPU box: parity tests, configs[2] timings (far / arms), then the contact statistics of a -DBH_STATS build.
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "caps or capsule" 2>&1 | tail -3
for c in none far arms; do timeout 300 python tests/reports/config3.py --caps $c --frames 10 --settle 30 2>/dev/null; done > gpurun_out/config3_${TAG}.json
python - <<PY
import json
for l in open("gpurun_out/config3_${TAG}.json"):
    d = json.loads(l); print(d["capsules"], d["math"], "ms/launch %.3f frac %.3f" % (d["ms_per_launch"], d["roofline"]["frac"]), d.get("oracle_check", ""))
PY
if [ -f barbu_b200/lib/libbarbu_hair_stats.so ]; then
  cp barbu_b200/lib/libbarbu_hair.so /tmp/prod.so; cp barbu_b200/lib/libbarbu_hair_stats.so barbu_b200/lib/libbarbu_hair.so
  for c in far arms; do echo "stats $c"; timeout 300 python tests/reports/config3.py --caps $c --frames 2 --settle 30 --math exact --check 0 --log2s 20 2>&1 | grep BH_STATS | tail -1; done
  cp /tmp/prod.so barbu_b200/lib/libbarbu_hair.so
fi
