"""Debug aid: host-side time of each call of the pipelined end-to-end loop."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pygofpid_b200 import HotPath
from bench import WORKLOADS, morph_list, make_clip
wl = WORKLOADS["1080p_mog2_open_close_blobs"]
hp = HotPath((1080, 1920), detector="MOG2", morph=morph_list(wl["morph"]))
frames = make_clip(wl, 24)
pinned = torch.from_numpy(frames).pin_memory()
row, fb = 1920, 1920 * 1080
for t in range(24):
    hp.process_pinned(pinned[t].data_ptr(), row, fb)
torch.cuda.synchronize()
T = {"submit": [], "select": [], "blobs": [], "sync_process": []}
for i in range(49):
    if i < 48:
        a = time.perf_counter(); hp.submit_pinned(pinned[i % 24].data_ptr(), row, fb); T["submit"].append(time.perf_counter() - a)
    if i > 0:
        a = time.perf_counter(); hp.select(1 if i < 48 else 0); T["select"].append(time.perf_counter() - a)
        a = time.perf_counter(); recs, pts = hp.raw_blobs(0); T["blobs"].append(time.perf_counter() - a)
torch.cuda.synchronize()
for i in range(24):
    a = time.perf_counter(); hp.process_pinned(pinned[i].data_ptr(), row, fb); hp.raw_blobs(0); T["sync_process"].append(time.perf_counter() - a)
for k, v in T.items():
    v = np.array(v) * 1e6
    print("%-14s mean %8.1f us  median %8.1f  max %8.1f" % (k, v.mean(), np.median(v), v.max()))
