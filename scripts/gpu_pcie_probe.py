"""Raw pinned host->device copy bandwidth at the e2e leg's sizes (is 48 GB/s the link or the pipeline?)."""
import torch
dev = torch.device("cuda:0")
for mib in (8, 64, 256):
    h = torch.empty(mib << 18, dtype=torch.float32).pin_memory()
    d = torch.empty_like(h, device=dev)
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        d.copy_(h, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"H2D {mib} MiB pinned: {ms*1e3:.1f} us  {mib * 1.048576 / ms:.1f} GB/s", flush=True)
