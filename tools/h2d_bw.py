"""Raw pinned host->device copy bandwidth at the e2e call's transfer sizes (what bounds bench.py's e2e value)."""
import torch, time
dev = torch.device("cuda", 0)
for mb in (5.1, 25.7, 102.7, 410.0):
    n = int(mb * 1e6 / 4)
    h = torch.empty(n, dtype=torch.float32).pin_memory()
    h.normal_()
    d = torch.empty(n, dtype=torch.float32, device=dev)
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"H2D {mb:7.1f} MB: {ms*1e3:8.1f} us  {n*4/ms/1e6:6.1f} GB/s")
