"""Host<->device bandwidth of pinned (cudaHostRegister'ed NumPy, torch pinned) and pageable memory."""
import time
import numpy as np
import torch

n = 256 * 1024 * 1024 // 4
dev = torch.empty(n, dtype=torch.float32, device="cuda")
for name, host in (("torch pinned", torch.empty(n, dtype=torch.float32, pin_memory=True)),
                   ("pageable", torch.empty(n, dtype=torch.float32))):
    for direction in ("h2d", "d2h"):
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            t0 = time.perf_counter()
            if direction == "h2d":
                dev.copy_(host, non_blocking=True)
            else:
                host.copy_(dev, non_blocking=True)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        print(f"{name:14s} {direction}: {n * 4 / best / 1e9:6.1f} GB/s")
# full duplex with pinned buffers on two streams
a = torch.empty(n, dtype=torch.float32, pin_memory=True)
b = torch.empty(n, dtype=torch.float32, pin_memory=True)
d2 = torch.empty(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
with torch.cuda.stream(s1):
    dev.copy_(a, non_blocking=True)
with torch.cuda.stream(s2):
    b.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"duplex h2d+d2h: {2 * n * 4 / dt / 1e9:6.1f} GB/s aggregate")
