#!/usr/bin/env python
"""PCIe copy rates of this box with pinned buffers (what bounds the end-to-end SMEM figure): D2H alone, H2D alone, both at once."""
import time, torch
n = 1 << 30
h_a = torch.empty(n, dtype=torch.uint8).pin_memory(); h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(d2h, h2d, reps=5):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps):
        if d2h:
            with torch.cuda.stream(s1): h_a.copy_(d_a, non_blocking=True)
        if h2d:
            with torch.cuda.stream(s2): d_b.copy_(h_b, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    return reps * n / dt / 1e9
run(True, True, 1)
print("D2H alone %.1f GB/s, H2D alone %.1f GB/s, both: %.1f GB/s each direction" % (run(True, False), run(False, True), run(True, True)))
