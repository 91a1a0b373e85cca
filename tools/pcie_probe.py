#!/usr/bin/env python3
"""PCIe ceiling on this box: pinned H2D / D2H alone and both at once (what bh_step_host can reach at best)."""
import torch, time
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, chunks=1, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter()
        c = n // chunks
        for i in range(chunks):
            if h2d:
                with torch.cuda.stream(s1): d_in[i*c:(i+1)*c].copy_(h_in[i*c:(i+1)*c], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h_out[i*c:(i+1)*c].copy_(d_out[i*c:(i+1)*c], non_blocking=True)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
    return n / best / 1e9
print("H2D alone   %.1f GB/s" % run(True, False))
print("D2H alone   %.1f GB/s" % run(False, True))
for ch in (1, 8, 32, 128):
    print("both, %3d chunks: %.1f GB/s each way" % (ch, run(True, True, ch)))
