#!/bin/bash
# First-contact run on the B200 box: GPU tests, smoke, a short bench. Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_exact.json 2> gpurun_out/bench_exact.err; echo "bench rc=$?"
cat gpurun_out/bench_exact.json; tail -3 gpurun_out/bench_exact.err
timeout 600 python bench.py --steps 10 --warmup 3 --math fast --no-cpu-baseline > gpurun_out/bench_fast.json 2> gpurun_out/bench_fast.err
cat gpurun_out/bench_fast.json; tail -3 gpurun_out/bench_fast.err
