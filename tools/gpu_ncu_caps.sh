#!/bin/bash
# One full ncu capture (with source) of the capsule variant on the "arms" scene at 2^20 strands: tools/gpu_ncu_caps.sh <tag> [exact|fast]
TAG=${1:-x}; M=${2:-exact}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hair_step_stream -s 130 -c 1 -f -o gpurun_out/prof_caps_arms_${M}_${TAG} \
  python tests/reports/config3.py --caps arms --frames 1 --settle 32 --math $M --check 0 --log2s 20 > gpurun_out/ncu_caps_arms_${M}_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_caps_arms_${M}_${TAG}.log
