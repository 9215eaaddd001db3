"""Debug aid: phase cycle counts of sampled mog2_kernel tiles (needs libgofpid_b200_dbg.so: make -C pygofpid_b200/csrc debug)."""
import ctypes as ct, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygofpid_b200 import _native as nat
nat.LIB_PATH = os.path.join(os.path.dirname(nat.LIB_PATH), "libgofpid_b200_dbg.so")
import torch
from pygofpid_b200 import HotPath
from bench import WORKLOADS, make_clip
wl = WORKLOADS["1080p_mog2_open_close_blobs"]
hp = HotPath((1080, 1920), detector="MOG2", morph=(), find_blobs=False)
frames = make_clip(wl, 24)
dev = torch.from_numpy(frames).cuda()
sh = torch.cuda.current_stream().cuda_stream
for t in range(120):
    hp.process_device(dev[t % 24].data_ptr(), 1920, 1920 * 1080, sh)
torch.cuda.synchronize()
buf = (ct.c_longlong * 512)()
nat.lib().gofpid_debug_mog2_dump(buf)
a = np.array(buf[:]).reshape(64, 8)
a = a[a[:, 0] > 0]
d = np.diff(a[:, :7], axis=1)
print("sampled tiles:", len(a))
print("phase cycles (median / mean): setup+blur %d/%d | wait model %d/%d | fast path %d/%d | general path %d/%d | barrier+issue store %d/%d | store drain %d/%d | total %d" % (
    np.median(d[:, 0]), d[:, 0].mean(), np.median(d[:, 1]), d[:, 1].mean(), np.median(d[:, 2]), d[:, 2].mean(),
    np.median(d[:, 3]), d[:, 3].mean(), np.median(d[:, 4]), d[:, 4].mean(), np.median(d[:, 5]), d[:, 5].mean(),
    np.median(a[:, 6] - a[:, 0])))
