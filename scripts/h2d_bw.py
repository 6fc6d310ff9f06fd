"""Pinned host->device copy bandwidth on this box (the e2e ceiling: 6912 B per window)."""
import torch
x = torch.empty(65536 * 1728, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
for n in (16384, 65536):
    v = x[:n * 1728]
    for _ in range(3):
        d[:n * 1728].copy_(v, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        d[:n * 1728].copy_(v, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("H2D %6d windows: %.3f ms  %.1f GB/s  -> %.2f M windows/s" % (n, ms, n * 6912 / ms / 1e6, n / ms / 1e3))
