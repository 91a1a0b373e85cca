#!/bin/bash
# Bench both math profiles, then the ncu launch list and one full capture of the step kernel.
# Usage: bash tools/gpu_profile.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
for m in exact fast; do
  timeout 600 python bench.py --steps 50 --warmup 3 --math $m $([ $m = fast ] && echo --no-cpu-baseline) \
      > gpurun_out/bench_${m}_${TAG}.json 2> gpurun_out/bench_${m}_${TAG}.err
  cat gpurun_out/bench_${m}_${TAG}.json; tail -2 gpurun_out/bench_${m}_${TAG}.err
done
PROF="python bench.py --steps 2 --warmup 3 --preroll 0 --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $PROF > gpurun_out/ncu_launches_${TAG}.log 2>&1
for m in exact fast; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:hair_step_pipelined -s 6 -c 2 \
      -f -o gpurun_out/prof_${m}_${TAG} $PROF --math $m > gpurun_out/ncu_full_${m}_${TAG}.log 2>&1
  tail -2 gpurun_out/ncu_full_${m}_${TAG}.log
done
ls -la gpurun_out
