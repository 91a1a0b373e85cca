#!/bin/bash
# Debug build with contact statistics (-DBH_STATS) into barbu_b200/lib/libbarbu_hair_stats.so; the product build is untouched.
cd "$(dirname "$0")/.."
cd barbu_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DBH_STATS -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math -ccbin /usr/bin/g++ -shared \
  -o ../lib/libbarbu_hair_stats.so hair_step.cu hair_stream.cu hair_wave.cu hair_gen.cu hair_tess.cu hair_state.cu hair_marschner.cu hair_capi.cu hair_group.cu hair_host.cc
