#!/bin/bash
# One full ncu capture of the step kernel per math profile named in $MATHS. Usage: tools/gpu_ncu.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
PROF="python bench.py --steps 2 --warmup 3 --preroll 0 --settle 0 --no-cpu-baseline --no-e2e --no-other-profile"
for m in ${MATHS:-exact fast}; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:hair_step_ -s 6 -c 1 \
      -f -o gpurun_out/prof_${m}_${TAG} $PROF --math $m > gpurun_out/ncu_full_${m}_${TAG}.log 2>&1
  tail -1 gpurun_out/ncu_full_${m}_${TAG}.log
done
