#!/bin/bash
# A/B of variant builds on configs[2] (none / far / arms, exact by default): tools/gpu_capab.sh <variant> ...   (MATHS="exact fast")
mkdir -p gpurun_out
cp barbu_b200/lib/libbarbu_hair.so /tmp/prod.so
run() { timeout 300 python tests/reports/config3.py --caps $1 --frames 5 --settle 30 --math $2 --check ${CHECK:-0} 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$3', d['capsules'], d['math'], 'ms/launch %.3f frac %.3f' % (d['ms_per_launch'], d['roofline']['frac']), d.get('oracle_check',''))"; }
{
for v in prod "$@"; do
  if [ $v = prod ]; then cp /tmp/prod.so barbu_b200/lib/libbarbu_hair.so; else cp barbu_b200/lib/libbarbu_hair_$v.so barbu_b200/lib/libbarbu_hair.so; fi
  for m in ${MATHS:-exact}; do run far $m $v; CHECK=2048 run arms $m $v; done
done
cp /tmp/prod.so barbu_b200/lib/libbarbu_hair.so
} | tee gpurun_out/capab.txt
