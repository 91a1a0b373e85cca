#!/bin/bash
# A/B of programmatic dependent launch (BH_STREAM_PDL=0/1) on the bench workload, both profiles, twice; then the GPU suite.
for rep in 1 2; do for pdl in 0 1; do for m in exact fast; do
  BH_STREAM_PDL=$pdl python bench.py --steps 50 --warmup 3 --math $m --no-cpu-baseline --no-e2e --no-other-profile 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('pdl=$pdl $m ms/launch %.4f frac %.3f fused ms/step %.4f'%(d['roofline']['ms_per_launch'], d['roofline']['frac'], d.get('value_fused',{}).get('ms_per_step',0)))"
done; done; done
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 2
