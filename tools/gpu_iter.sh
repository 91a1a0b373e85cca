#!/bin/bash
# Iteration run: GPU tests, then benches for the variants named in $VARIANTS ("occ:math" pairs).
TAG=${1:-it}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_${TAG}.log
for v in ${VARIANTS:-4:exact 4:fast 3:exact 3:fast}; do
  occ=${v%%:*}; m=${v##*:}
  BH_SCHED_OCC=$occ timeout 600 python bench.py --steps 50 --warmup 3 --math $m --no-cpu-baseline --no-e2e --no-other-profile > gpurun_out/bench_${m}_occ${occ}_${TAG}.json 2>gpurun_out/bench_${m}_occ${occ}_${TAG}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${m}_occ${occ}_${TAG}.json"))
    print("$m occ$occ", "ms/launch %.4f frac %.3f value %.3e clocks %s" % (d["roofline"]["ms_per_launch"], d["roofline"]["frac"], d["value"], d["clocks"]))
except Exception as e:
    print("$m occ$occ FAILED", e); print(open("gpurun_out/bench_${m}_occ${occ}_${TAG}.err").read()[-800:])
PY
done
