"""Debug aid: when do stage A (detector) and stage B (morphology .. blobs) of every frame slot end,
relative to a mark on the launch stream?  GOFPID_MOG2_GROUP / GOFPID_DEPTH as usual.  Needs the debug
build, which alone exports gofpid_debug_*: `make -C pygofpid_b200/csrc debug`, then GOFPID_LIB=libgofpid_b200_dbg.so."""
import os, sys
import ctypes as ct
os.environ["GOFPID_EVENT_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pygofpid_b200 import HotPath
from pygofpid_b200 import _native as nat
from bench import WORKLOADS, morph_list, make_clip
wl = WORKLOADS["1080p_mog2_open_close_blobs"]
NCLIP = 24
dev = torch.from_numpy(make_clip(wl, NCLIP)).cuda()
row, fb = 1920, 1920 * 1080
sh = torch.cuda.current_stream().cuda_stream
hp = HotPath((1080, 1920), detector="MOG2", morph=morph_list(wl["morph"]), find_blobs=True)
L = hp._L
L.gofpid_debug_mark.argtypes = [ct.c_void_p, ct.c_void_p]
L.gofpid_debug_slot_times.argtypes = [ct.c_void_p, ct.POINTER(ct.c_float), ct.POINTER(ct.c_int)]
for _ in range(8):
    hp.enqueue_device_sequence(dev.data_ptr(), NCLIP, fb, row, fb, sh)
L.gofpid_debug_mark(hp._h, sh)
hp.enqueue_device_sequence(dev.data_ptr(), NCLIP, fb, row, fb, sh)
D = hp.depth
ms = (ct.c_float * (2 * D))()
fi = (ct.c_int * D)()
L.gofpid_debug_slot_times(hp._h, ms, fi)
rows = sorted((fi[i], ms[2 * i] * 1e3, ms[2 * i + 1] * 1e3) for i in range(D))
print("group", hp.group, "depth", D, "(us after the mark; the clip call enqueued %d frames)" % NCLIP)
for f, a, b in rows:
    print("frame %4d  stage A end %8.1f   stage B end %8.1f   (B - A %7.1f)" % (f, a, b, b - a))
