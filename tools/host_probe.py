"""What the GPU box's host offers for the end-to-end path: cores, memory bandwidth of a
multi-threaded fill (numpy, one array slice per thread), pinned D2H bandwidth."""
import os
import sys
import threading
import time

import numpy as np

print("nproc", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
try:
    with open("/proc/cpuinfo") as f:
        names = [ln.split(":")[1].strip() for ln in f if ln.startswith("model name")]
    print("cpu", names[0] if names else "?", "x", len(names))
    with open("/proc/meminfo") as f:
        print(f.readline().strip())
except OSError:
    pass
n = 1 << 29   # 4 GiB of doubles
a = np.empty(n)
a[:] = 0.0    # fault in
for nt in (1, 2, 4, 8, 16, 32):
    if nt > 2 * (os.cpu_count() or 1):
        break
    parts = np.array_split(np.arange(n), 1)  # placeholder (avoid big arange)
    bounds = [(i * n // nt, (i + 1) * n // nt) for i in range(nt)]

    def work(lo, hi):
        a[lo:hi] = np.nan

    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=b) for b in bounds]
    [t.start() for t in th]
    [t.join() for t in th]
    dt = time.perf_counter() - t0
    print(f"numpy fill {nt:2d} threads: {n * 8 / dt / 1e9:6.1f} GB/s")
if "--cuda" in sys.argv:
    import torch
    d = torch.empty(n // 2, dtype=torch.float64, device="cuda")
    h = torch.empty(n // 2, dtype=torch.float64, pin_memory=True)
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h.copy_(d, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"pinned D2H: {d.numel() * 8 / dt / 1e9:.1f} GB/s")
