"""Debug aid: phase cycle counts of runs_ccl_kernel (needs libgofpid_b200_dbg.so: make -C pygofpid_b200/csrc debug)."""
import ctypes as ct, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["GOFPID_NO_GRAPH"] = "1"
from pygofpid_b200 import _native as nat
nat.LIB_PATH = os.path.join(os.path.dirname(nat.LIB_PATH), "libgofpid_b200_dbg.so")
from pygofpid_b200 import HotPath
from bench import WORKLOADS, morph_list, make_clip
wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "1080p_mog2_open_close_blobs"]
hp = HotPath((wl["H"], wl["W"]) if wl["C"] == 1 else (wl["H"], wl["W"], wl["C"]), detector=wl["detector"],
             morph=morph_list(wl["morph"]))
frames = make_clip(wl, 72)
L = nat.lib()
buf = (ct.c_longlong * 16)()
for t in range(160):
    hp.process(frames[t % 72])
    if t in (1, 10, 40, 100, 159):
        L.gofpid_debug_rc_dump(buf)
        a = np.array(buf[:])
        print("frame", t, "runs", a[8], "comps", a[9], "phase cycles load C D E F:", [int(a[3]-a[0]), int(a[4]-a[3]), int(a[5]-a[4]), int(a[6]-a[5]), int(a[7]-a[6])], hp.frame_info())
