"""Host<->device copy bandwidth of the box (pinned memory), for reading the
end-to-end numbers of bench.py: python tools/pcie_bw.py [GiB]"""
import sys
import time

import torch

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
n = int(gib * (1 << 30) / 8)
dev = torch.empty(n, dtype=torch.float64, device="cuda")
t0 = time.perf_counter()
host = torch.empty(n, dtype=torch.float64, pin_memory=True)
print(f"pinned allocation of {gib} GiB: {time.perf_counter() - t0:.2f} s")
for name, dst, src in (("D2H", host, dev), ("H2D", dev, host)):
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"{name}: {n * 8 / dt / 1e9:.1f} GB/s")
