for o in 1 2 3; do for m in exact fast; do
BH_STREAM_BLOCKS_PER_SM=$o python bench.py --steps 30 --warmup 3 --math $m --no-cpu-baseline --no-e2e --no-other-profile 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('occ $o $m ms/launch %.4f frac %.3f'%(d['roofline']['ms_per_launch'], d['roofline']['frac']))"
done; done
