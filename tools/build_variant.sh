#!/bin/bash
# Build a variant of libbarbu_hair.so with extra nvcc flags for A/B runs (tools/ab2.sh): tools/build_variant.sh <name> [flags...]
# -> barbu_b200/lib/libbarbu_hair_<name>.so  (git-ignored; travels to the GPU box)
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
cd barbu_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math -ccbin /usr/bin/g++ -shared "$@" \
  -o ../lib/libbarbu_hair_$NAME.so hair_step.cu hair_stream.cu hair_wave.cu hair_gen.cu hair_tess.cu hair_state.cu hair_marschner.cu hair_capi.cu hair_group.cu hair_host.cc
echo built barbu_b200/lib/libbarbu_hair_$NAME.so
