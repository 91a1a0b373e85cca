#!/bin/bash
# Full evidence run: probe, GPU tests, smoke, benches (exact+fast), ncu launch list + full captures.
TAG=${1:-r01}
mkdir -p gpurun_out
bash tools/probe_box.sh > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_${TAG}.log
bash tools/gpu_profile.sh ${TAG}
