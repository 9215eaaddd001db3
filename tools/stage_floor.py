"""Debug aid: device-resident frame rate of partial pipelines (which stage bounds the step)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pygofpid_b200 import HotPath
from bench import WORKLOADS, morph_list, make_clip
wl = WORKLOADS["1080p_mog2_open_close_blobs"]
frames = make_clip(wl, 24)
dev = torch.from_numpy(frames).cuda()
row, fb = 1920, 1920 * 1080
sh = torch.cuda.current_stream().cuda_stream
for name, kw in (("mog2 only", dict(morph=(), find_blobs=False)),
                 ("mog2 + morph", dict(morph=morph_list(wl["morph"]), find_blobs=False)),
                 ("full", dict(morph=morph_list(wl["morph"]), find_blobs=True))):
    hp = HotPath((1080, 1920), detector="MOG2", **kw)
    for t in range(100):
        hp.process_device(dev[t % 24].data_ptr(), row, fb, sh)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    for t in range(400):
        hp.process_device(dev[t % 24].data_ptr(), row, fb, sh)
    w1 = time.perf_counter()
    e1.record()
    torch.cuda.synchronize()
    print("%-14s %7.1f us/frame on the device, host issue %6.1f us/frame" % (name, e0.elapsed_time(e1) * 1000 / 400, (w1 - w0) * 1e6 / 400))
    hp.close()
