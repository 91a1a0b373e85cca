#!/bin/bash
# Capsule variant: parity tests, then the "arms" scene — pass rates and warp-time shares from the stats build (with its count of
# skipped tests that would have passed: must be zero), timings of none / far / arms, optional variant builds.
mkdir -p gpurun_out
cp barbu_b200/lib/libbarbu_hair.so /tmp/prod.so
run() { timeout 300 python tests/reports/config3.py --caps $1 --frames 5 --settle 30 --math $2 --check ${CHECK:-0} 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$3', d['capsules'], d['math'], 'ms/launch %.3f frac %.3f' % (d['ms_per_launch'], d['roofline']['frac']), d.get('oracle_check',''))"; }
{
timeout 900 python -m pytest tests -m gpu -x -q -k "caps or capsule or shell" 2>&1 | tail -3
cp barbu_b200/lib/libbarbu_hair_stats.so barbu_b200/lib/libbarbu_hair.so
for m in exact fast; do echo "stats arms $m"; timeout 300 python tests/reports/config3.py --caps arms --frames 1 --settle 30 --math $m --check 0 --log2s 20 2>&1 | grep BH_STATS | tail -1; done
cp /tmp/prod.so barbu_b200/lib/libbarbu_hair.so
for m in exact fast; do run none $m prod; run far $m prod; CHECK=2048 run arms $m prod; done
for v in "$@"; do cp barbu_b200/lib/libbarbu_hair_$v.so barbu_b200/lib/libbarbu_hair.so; for m in exact fast; do run far $m $v; run arms $m $v; done; done
cp /tmp/prod.so barbu_b200/lib/libbarbu_hair.so
} | tee gpurun_out/capexp.txt
