#!/bin/bash
# Iteration run: GPU tests (optionally a -k filter), then benches of both math profiles, optional ncu capture.
# Usage: tools/gpu_iter2.sh <tag> [pytest -k expr] ; env: NCU=1 for full captures, STEPS
TAG=${1:-it}
mkdir -p gpurun_out
if [ -n "$2" ]; then K=(-k "$2"); else K=(); fi
timeout 1200 python -m pytest tests -m gpu -x -q "${K[@]}" > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_${TAG}.log
for m in exact fast; do
  timeout 600 python bench.py --steps ${STEPS:-50} --warmup 3 --math $m --no-cpu-baseline --no-e2e --no-other-profile > gpurun_out/bench_${m}_${TAG}.json 2>gpurun_out/bench_${m}_${TAG}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${m}_${TAG}.json"))
    print("$m", "ms/launch %.4f frac %.3f value %.3e clocks %s" % (d["roofline"]["ms_per_launch"], d["roofline"]["frac"], d["value"], d["clocks"]))
except Exception as e:
    print("$m FAILED", e); print(open("gpurun_out/bench_${m}_${TAG}.err").read()[-1500:])
PY
done
if [ -n "$NCU" ]; then
  PROF="python bench.py --steps 2 --warmup 3 --preroll 0 --settle 0 --no-cpu-baseline --no-e2e --no-other-profile"
  for m in ${MATHS:-exact fast}; do
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:hair_step_ -s 6 -c 1 \
        -f -o gpurun_out/prof_${m}_${TAG} $PROF --math $m > gpurun_out/ncu_full_${m}_${TAG}.log 2>&1
    tail -1 gpurun_out/ncu_full_${m}_${TAG}.log
  done
fi
