import torch, time
for mb in (2, 8, 64):
    n = mb * 1024 * 1024
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(5): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 200 if mb < 64 else 20
    e0.record()
    for _ in range(reps): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("H2D %3d MB: %.1f us, %.1f GB/s" % (mb, ms * 1000, n / ms / 1e6))
    e0.record()
    for _ in range(reps): h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("D2H %3d MB: %.1f us, %.1f GB/s" % (mb, ms * 1000, n / ms / 1e6))
