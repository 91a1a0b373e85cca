#!/bin/bash
# Host<->device ceiling for 1/2/4/8 concurrent ranks, with and without NUMA affinity, plus the box topology. Needs gpurun --gpus 8.
mkdir -p gpurun_out
OUT=gpurun_out/pcie_probe_${1:-r02}.txt
{ nvidia-smi topo -m; echo; lscpu | grep -i "numa\|socket\|model name\|^CPU(s)"; echo; } > $OUT 2>&1
NG=$(nvidia-smi -L | wc -l)
for n in 1 2 4 8; do
  [ $n -le $NG ] || continue
  for aff in "" $([ $n = $NG ] && echo "--affinity"); do        # the affinity variant once, at full width
    if [ $n = 1 ]; then timeout 300 python tools/pcie_probe.py $aff 2>/dev/null | grep '^{' >> $OUT
    else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) tools/pcie_probe.py $aff 2>/dev/null | grep '^{' >> $OUT; fi
  done
done
cat $OUT
