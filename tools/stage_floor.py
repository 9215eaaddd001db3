"""Debug aid: device-resident frame rate of partial pipelines (which stage bounds the step), frame by
frame (process_device) and through the clip entry point with G frames per MOG2 launch.
GOFPID_LIB selects a variant build (csrc/Makefile: variant)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pygofpid_b200 import HotPath
from bench import WORKLOADS, morph_list, make_clip
wl = WORKLOADS["1080p_mog2_open_close_blobs"]
NCLIP = int(os.environ.get("FLOOR_NCLIP", "24"))  # 24 frames = 50 MB: the clip stays in L2; 72: it streams from HBM as in bench.py
frames = make_clip(wl, NCLIP)
dev = torch.from_numpy(frames).cuda()
row, fb = 1920, 1920 * 1080
sh = torch.cuda.current_stream().cuda_stream
groups = [int(g) for g in (sys.argv[1].split(",") if len(sys.argv) > 1 else "0,1,2,4,8".split(","))]
parts = (("mog2 only", dict(morph=(), find_blobs=False)),
         ("mog2 + morph", dict(morph=morph_list(wl["morph"]), find_blobs=False)),
         ("full", dict(morph=morph_list(wl["morph"]), find_blobs=True)))
for G in groups:  # 0: frame-by-frame calls
    if G:
        os.environ["GOFPID_MOG2_GROUP"] = str(G)
    for name, kw in parts:
        hp = HotPath((1080, 1920), detector="MOG2", **kw)
        def run(n):
            if G == 0:
                for t in range(n):
                    hp.process_device(dev[t % NCLIP].data_ptr(), row, fb, sh)
            else:
                for t in range(0, n, NCLIP):
                    hp.enqueue_device_sequence(dev.data_ptr(), NCLIP, fb, row, fb, sh)
        run(96)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        N = 480
        w0 = time.perf_counter()
        e0.record()
        run(N)
        w1 = time.perf_counter()
        e1.record()
        torch.cuda.synchronize()
        print("G=%d depth=%2d %-14s %7.1f us/frame on the device, host issue %6.1f us/frame" %
              (G, hp.depth, name, e0.elapsed_time(e1) * 1000 / N, (w1 - w0) * 1e6 / N), flush=True)
        hp.close()
