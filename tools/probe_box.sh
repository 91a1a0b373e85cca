mkdir -p gpurun_out
{
echo "== nvidia-smi"; nvidia-smi -L; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv
echo "== cpu"; nproc; lscpu | grep -E 'Model name|Socket|Core|Thread|^CPU\(s\)'; free -g | head -2
echo "== GL probe"; ldconfig -p | grep -E 'libEGL|libGLX|libOSMesa|libGL\.|libnvidia-(egl|gl)core|libX11' || echo "no GL libs in ldconfig"
find /usr/share /etc -name "*_icd.json" -o -name "10_nvidia.json" 2>/dev/null | head; echo "NVIDIA_DRIVER_CAPABILITIES=$NVIDIA_DRIVER_CAPABILITIES"
ls /usr/lib/x86_64-linux-gnu | grep -iE 'nvidia|egl|gl' | head -30
echo "== topo"; nvidia-smi topo -m 2>&1 | head -20
} > gpurun_out/probe_box.txt 2>&1
cat gpurun_out/probe_box.txt
