#!/bin/bash
# Bench (unfused + fused legs) and one full ncu capture of a FUSED launch (4 substeps as 4 passes) in the settled state.
TAG=${1:-r02}
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_fused_${TAG}.json 2> gpurun_out/bench_fused_${TAG}.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_fused_${TAG}.json'))
o=d['other_profile']
print('fast  unfused %.4f ms/frame frac %.3f | fused %.4f ms/frame' % (d['ms_per_step'], d['roofline']['frac'], d['value_fused']['ms_per_step']))
print('exact unfused %.4f ms/frame frac %.3f | fused %.4f ms/frame' % (o['ms_per_step'], o['roofline']['frac'], o['value_fused']['ms_per_step']))
PY
PROF="python bench.py --steps 2 --warmup 3 --preroll 0 --sustain 0 --no-cpu-baseline --no-e2e --no-other-profile --no-fused --fused-main"
for m in ${MATHS:-fast}; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:hair_step_ -s 103 -c 1 \
      -f -o gpurun_out/prof_fused_${m}_${TAG} $PROF --math $m > gpurun_out/ncu_fused_${m}_${TAG}.log 2>&1
  tail -1 gpurun_out/ncu_fused_${m}_${TAG}.log
done
