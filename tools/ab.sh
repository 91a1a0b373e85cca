#!/bin/bash
# A/B two builds of libbarbu_hair.so on the bench workload: tools/ab.sh <old.so> ; "new" is the in-tree build.
OLD=$1
cp barbu_b200/lib/libbarbu_hair.so /tmp/new.so
for rep in 1 2; do
for which in old new; do
  if [ $which = old ]; then cp $OLD barbu_b200/lib/libbarbu_hair.so; else cp /tmp/new.so barbu_b200/lib/libbarbu_hair.so; fi
  python bench.py --steps 50 --warmup 3 --math exact --no-cpu-baseline --no-e2e --no-other-profile 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$which exact ms/launch %.4f frac %.3f'%(d['roofline']['ms_per_launch'], d['roofline']['frac']))"
done; done
cp /tmp/new.so barbu_b200/lib/libbarbu_hair.so
