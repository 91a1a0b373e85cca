#!/bin/bash
# Evidence run: default bench (both math profiles in one line, e2e, cpu baseline), reference arm, the ncu launch list
# of the bench command and one full capture of the step kernel per math profile, all in the SETTLED state the bench
# times (100 settle frames + 3 warm-up steps = 412 launches are skipped).  Usage: bash tools/gpu_profile.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
cat gpurun_out/bench_${TAG}.json; tail -2 gpurun_out/bench_${TAG}.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference_${TAG}.json 2> gpurun_out/bench_reference_${TAG}.err
cat gpurun_out/bench_reference_${TAG}.json
PROF="python bench.py --steps 2 --warmup 3 --preroll 0 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 412 -c 1200 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $PROF > gpurun_out/ncu_launches_${TAG}.log 2>&1
PROF="python bench.py --steps 2 --warmup 3 --preroll 0 --no-cpu-baseline --no-e2e --no-other-profile"
for m in exact fast; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:hair_step_ -s 412 -c 1 \
      -f -o gpurun_out/prof_${m}_${TAG} $PROF --math $m > gpurun_out/ncu_full_${m}_${TAG}.log 2>&1
  tail -1 gpurun_out/ncu_full_${m}_${TAG}.log
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,power.limit --format=csv > gpurun_out/smi_${TAG}.txt
ls -la gpurun_out | head -40
