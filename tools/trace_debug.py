"""Debug aid: per-candidate phase cycle counts of trace_kernel (needs libgofpid_b200_dbg.so,
built with -DGOFPID_TRACE_DEBUG; see the commands in DESIGN.md)."""
import ctypes as ct, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygofpid_b200 import _native as nat
nat.LIB_PATH = os.path.join(os.path.dirname(nat.LIB_PATH), "libgofpid_b200_dbg.so")
from pygofpid_b200 import HotPath
from bench import WORKLOADS, morph_list, make_clip
wl = WORKLOADS["1080p_mog2_open_close_blobs"]
hp = HotPath((1080, 1920), detector="MOG2", morph=morph_list(wl["morph"]))
frames = make_clip(wl, 40)
L = nat.lib()
buf = (ct.c_longlong * (64 * 16))()
for t in range(40):
    hp.process(frames[t])
    n = L.gofpid_debug_trace_dump(buf)
    if t in (1, 10, 25, 39):
        a = np.array(buf[:]).reshape(64, 16)[:min(n, 64)]
        print("frame", t, "candidates", n, hp.frame_info())
        print("  ci   bw   bh  ncp rmask keep |   enum  pass1 rank+sum decide pass2 (cycles) cta rank round1")
        for r in a[np.argsort(-a[:, 6:11].sum(1))][:8]:
            print(" %3d %4d %4d %4d %4d %4d | %6d %6d %6d %6d %6d  %d %d %d" % tuple(r[:14]))
