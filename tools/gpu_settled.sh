mkdir -p gpurun_out
for m in exact fast; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hair_step_ -s 420 -c 1 -f -o gpurun_out/prof_${m}_settled python bench.py --steps 2 --warmup 110 --preroll 0 --settle 0 --no-cpu-baseline --no-e2e --no-other-profile --math $m > gpurun_out/ncu_settled_${m}.log 2>&1
tail -1 gpurun_out/ncu_settled_${m}.log
done
