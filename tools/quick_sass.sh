#!/bin/bash
# Compile hair_stream.cu alone to a cubin and print register counts + static step-path lengths (no GPU needed).
# Usage: tools/quick_sass.sh [kernel-substring] [extra nvcc flags...]
cd "$(dirname "$0")/.."
PAT=${1:-PackedExactELb1ELi8ELb0ELb0E}; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -cubin -Xptxas -v "$@" -o /tmp/hair_stream.cubin barbu_b200/csrc/hair_stream.cu 2> /tmp/hair_stream.ptxas.txt || { tail -20 /tmp/hair_stream.ptxas.txt; exit 1; }
grep -A1 "$PAT" /tmp/hair_stream.ptxas.txt | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" | paste -sd' '
python tools/sass_paths.py /tmp/hair_stream.cubin "$PAT" 300 | cut -c1-260
