"""Host<->device copy bandwidth of this box with pinned memory (context for the e2e figures: predict returns 240 MB per 1e7 BAUs)."""
import torch, time
for mb in (80, 240):
    n = mb * (1 << 20) // 8
    d = torch.empty(n, dtype=torch.float64, device="cuda")
    h = torch.empty(n, dtype=torch.float64, pin_memory=True)
    for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
        for _ in range(2): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5): fn()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
        print(f"{name} {mb} MB pinned: {mb / 1024 / dt:.1f} GiB/s ({1e3 * dt:.2f} ms)")
