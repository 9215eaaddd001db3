"""Debug aid: end-to-end frame rate (pinned host frames in) of partial pipelines -- is the host link
or the device the limit?  Frames are submitted with gofpid_pipeline_submit_host; only the full
pipeline also collects blob lists."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pygofpid_b200 import HotPath
from bench import WORKLOADS, morph_list, make_clip
wl = WORKLOADS["1080p_mog2_open_close_blobs"]
NCLIP = 72
pinned = torch.from_numpy(make_clip(wl, NCLIP)).pin_memory()
row, fb = 1920, 1920 * 1080
for name, kw, collect in (("H2D + mog2", dict(morph=(), find_blobs=False), False),
                          ("H2D + mog2 + morph", dict(morph=morph_list(wl["morph"]), find_blobs=False), False),
                          ("full, not collected", dict(morph=morph_list(wl["morph"]), find_blobs=True), False),
                          ("full, collected", dict(morph=morph_list(wl["morph"]), find_blobs=True), True)):
    hp = HotPath((1080, 1920), detector="MOG2", **kw)
    LAG = hp.depth - 1
    def run(n):
        for i in range(n + (LAG if collect else 0)):
            if i < n:
                hp.submit_pinned(pinned[i % NCLIP].data_ptr(), row, fb)
            if collect and i >= LAG:
                hp.select(min(i, n - 1) - (i - LAG))
                hp.raw_blobs(0, copy=False)
        torch.cuda.synchronize()
    run(100)
    N = 1500
    w0 = time.perf_counter()
    run(N)
    dt = time.perf_counter() - w0
    print("%-22s %7.1f us/frame end to end (%.0f frames/s)" % (name, dt * 1e6 / N, N / dt), flush=True)
    hp.close()
