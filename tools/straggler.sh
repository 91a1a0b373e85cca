#!/bin/bash
# Straggler diagnosis on ONE GPU: time each rank's shard of the 8-GPU weak-scaling job alone, for both strand orders.
# Row-major order: rank = latitude band (rank 7 = the pole where all hair lies on the collider); column-major: rank = wedge.
OUT=${1:-gpurun_out/straggler.txt}
: > $OUT
for order in row column; do
  for r in 0 2 4 5 6 7; do
    python bench.py --steps 20 --warmup 3 --order $order --emulate-world 8 --emulate-rank $r --no-e2e --no-cpu-baseline --sustain 0 2>/dev/null | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); o=d['other_profile']; print('order=$order rank=$r/8  fast ms/frame %.4f frac %.3f | exact ms/frame %.4f frac %.3f' % (d['ms_per_step'], d['roofline']['frac'], o['ms_per_step'], o['roofline']['frac']))" >> $OUT
  done
done
cat $OUT
