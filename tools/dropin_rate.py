"""Frame rate of the drop-in class on the headline scene: GOFPID.detect() frame by frame (one
synchronous GPU round trip per frame) against GOFPID.detect_clip() (clip entry point + host tail)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cv2 as cv
from pygofpid_b200 import GOFPID
from bench import WORKLOADS, make_clip
wl = WORKLOADS["1080p_mog2_open_close_blobs"]
clip = make_clip(wl, 72)
se = cv.getStructuringElement(cv.MORPH_RECT, (5, 5))
mm = [dict(fun=cv.erode, kernel=se), dict(fun=cv.dilate, kernel=se), dict(fun=cv.dilate, kernel=se), dict(fun=cv.erode, kernel=se)]
PF = {'perimeter': None, 'anchor': 'bottom', 'perspective': None, 'perspective_coeff': 0.5,
      'presence_min': 3, 'distance_min': 0.25}
def new():  # (own parameter dicts: the class, like the reference, calibrates them in place)
    return GOFPID(frg_detect='MOG2', mat_morph=[dict(d) for d in mm], post_filter=dict(PF)).initialize()
a, b = new(), new()
for f in clip[:24]:
    a.detect(f)
b.detect_clip(clip[:24])
t0 = time.perf_counter(); ya = [a.detect(f) for f in clip]; ta = time.perf_counter() - t0
t0 = time.perf_counter(); yb = b.detect_clip(clip); tb = time.perf_counter() - t0
assert [bool(v) for v in ya] == [bool(v) for v in yb], (ya, yb)
assert len(a.tracked_blobs_) == len(b.tracked_blobs_), (len(a.tracked_blobs_), len(b.tracked_blobs_))
print("detect() frame by frame: %.0f frames/s; detect_clip(): %.0f frames/s (%d blobs, %d tracks in the last frame)"
      % (72 / ta, 72 / tb, len(b.blobs_), len(b.tracked_blobs_)))
