"""Latency of ONE synchronous frame through the library (what GOFPID.detect() pays per call), split by where the
frame lives: pageable host memory (staged through pinned memory by the copier threads), pinned host memory, HBM."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import WORKLOADS, make_clip, new_hot_path
name = sys.argv[1] if len(sys.argv) > 1 else "1080p_mog2_open_close_blobs"
wl = WORKLOADS[name]
clip = make_clip(wl, 48)
H, W, C = wl["H"], wl["W"], wl["C"]
row = W * C
def us(f, n=96):
    for i in range(24): f(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n): f(i)
    return (time.perf_counter() - t0) / n * 1e6
hp = new_hot_path(wl, 1, 0)
pinned = torch.from_numpy(np.stack(clip)).pin_memory()
dev = pinned.cuda()
fb = pinned[0].numel()
print(name)
print("  pageable frame -> process():           %7.1f us" % us(lambda i: hp.process(clip[i % 48])))
print("  pinned frame   -> process_pinned():    %7.1f us" % us(lambda i: hp.process_pinned(pinned[i % 48].data_ptr(), row, row * H)))
def devstep(i):
    hp.process_device(dev[i % 48].data_ptr(), row, row * H); hp.frame_info()
print("  frame in HBM   -> process_device()+info: %5.1f us" % us(devstep))
hp.set_profiling(True)
for i in range(8):
    hp.process_pinned(pinned[i].data_ptr(), row, row * H)
    print("  stage_ms of a synchronous frame:", {k: round(v * 1e3, 1) for k, v in hp.stage_ms().items()})
hp.set_profiling(False)
t0 = time.perf_counter()
for i in range(200): torch.cuda.synchronize()
print("  empty synchronize: %.1f us" % ((time.perf_counter() - t0) / 200 * 1e6))
x = torch.empty(fb, dtype=torch.uint8, device="cuda")
def h2d(i):
    x.copy_(pinned[i % 48].view(-1), non_blocking=True); torch.cuda.synchronize()
print("  2 MB pinned H2D + synchronize: %.1f us" % us(h2d))
