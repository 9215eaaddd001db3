"""End-to-end frame rate of the CLIP entry point (pinned host frames in) for partial pipelines, next to the
plain copy of the same frames: where does the distance to the host-link floor come from?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pygofpid_b200 import HotPath
from bench import WORKLOADS, morph_list, make_clip
wl = WORKLOADS["1080p_mog2_open_close_blobs"]
NCLIP = 72
REP = int(os.environ.get("E2E_REP", "8"))
clip = torch.from_numpy(make_clip(wl, NCLIP)).pin_memory()
long_clip = clip.repeat(REP, 1, 1).pin_memory()
n = long_clip.shape[0]
dev = torch.empty((12,) + tuple(clip.shape[1:]), dtype=torch.uint8, device="cuda")
def copies(g):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(0, n - g + 1, g):
            dev[(i % 12):(i % 12) + g].copy_(long_clip[i:i + g], non_blocking=True)
    s.synchronize()
for g in (3, 6):
    copies(g)
    t0 = time.perf_counter(); copies(g); dt = time.perf_counter() - t0
    print("%-28s %7.1f us/frame (%.0f frames/s)" % ("copies alone, %d frames each" % g, dt * 1e6 / n, n / dt), flush=True)
for name, kw in (("mog2 + labelling", dict(morph=(), area_min=1e9)),
                 ("mog2 + morph + labelling", dict(morph=morph_list(wl["morph"]), area_min=1e9)),
                 ("full", dict(morph=morph_list(wl["morph"]), area_min=100.0))):
    hp = HotPath((1080, 1920), detector="MOG2", **kw)
    hp.process_sequence(clip)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); hp.process_sequence(long_clip); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    print("%-28s %7.1f us/frame (%.0f frames/s)  depth %d group %d" % (name, best * 1e6 / n, n / best, hp.depth, hp.group), flush=True)
    hp.close()
