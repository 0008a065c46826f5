"""Kernel-tuning helper: device->host copy-engine bandwidth into page-locked memory for the e2e output size (9.4 MB) and
for 1/16 of it, to judge shc_step_host's zero-copy store path (45 GB/s after the first wave of tiles) against the DMA path."""
import torch

dev = torch.device("cuda", 0)
for nbytes, label in ((131072 * 18 * 4, "9.4 MB"), (131072 * 18 * 4 // 16, "590 KB"), (256 << 20, "256 MB")):
    a = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    for _ in range(5):
        h.copy_(a, non_blocking=True)
    torch.cuda.synchronize()
    reps = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        h.copy_(a, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"D2H {label}: {ms * 1e3:.1f} us per copy, {nbytes / ms / 1e6:.1f} GB/s")
