#!/bin/bash
# A/B alternative builds on the bench workload: tools/ab2.sh <alt1.so> [alt2.so ...]; "prod" is the in-tree build.
cp barbu_b200/lib/libbarbu_hair.so /tmp/prod.so
for rep in 1 2; do
for which in prod "$@"; do
  if [ $which = prod ]; then cp /tmp/prod.so barbu_b200/lib/libbarbu_hair.so; else cp $which barbu_b200/lib/libbarbu_hair.so; fi
  for m in ${MATHS:-exact}; do
  python bench.py --steps 50 --warmup 3 --math $m --no-cpu-baseline --no-e2e --no-other-profile 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$which $m ms/launch %.4f frac %.3f fused ms/step %.4f'%(d['roofline']['ms_per_launch'], d['roofline']['frac'], d.get('value_fused',{}).get('ms_per_step',0)))"
  done
done; done
cp /tmp/prod.so barbu_b200/lib/libbarbu_hair.so
