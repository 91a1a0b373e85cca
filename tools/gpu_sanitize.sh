#!/bin/bash
# compute-sanitizer over tests/reports/sanitizer_case.py (all four tools) -> gpurun_out/sanitizer_<tag>.txt
TAG=${1:-r02}; OUT=gpurun_out/sanitizer_${TAG}.txt; mkdir -p gpurun_out
echo "# compute-sanitizer (CUDA 12.9) over tests/reports/sanitizer_case.py on 1x B200" > $OUT
for tool in memcheck racecheck initcheck synccheck; do
  echo "== $tool" >> $OUT
  timeout 900 compute-sanitizer --tool $tool python tests/reports/sanitizer_case.py 2>&1 | grep -v "^=========$" | grep "^ok\|^tess\|ERROR SUMMARY\|RACECHECK SUMMARY\|Error\|error\|hazard\|Traceback\|assert" | head -60 >> $OUT
done
echo "== memcheck + racecheck with BH_STREAM_BIG_BLOCKS=2 (one block of 12 warps per SM forced for every launch)" >> $OUT
for tool in memcheck racecheck; do
  BH_STREAM_BIG_BLOCKS=2 timeout 900 compute-sanitizer --tool $tool python tests/reports/sanitizer_case.py 2>&1 | grep -v "^=========$" | grep "ERROR SUMMARY\|RACECHECK SUMMARY\|Error\|error\|hazard\|Traceback\|assert" | head -20 >> $OUT
done
cat $OUT
