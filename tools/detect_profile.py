"""Where the time of one synchronous GOFPID.detect() goes: cProfile over a few hundred frames of the headline scene."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2 as cv
from pygofpid_b200 import GOFPID
from bench import WORKLOADS, make_clip
wl = WORKLOADS["1080p_mog2_open_close_blobs"]
clip = make_clip(wl, 72)
se = cv.getStructuringElement(cv.MORPH_RECT, (5, 5))
mm = [dict(fun=cv.erode, kernel=se), dict(fun=cv.dilate, kernel=se), dict(fun=cv.dilate, kernel=se), dict(fun=cv.erode, kernel=se)]
PF = {'perimeter': None, 'anchor': 'bottom', 'perspective': None, 'perspective_coeff': 0.5, 'presence_min': 3, 'distance_min': 0.25}
g = GOFPID(frg_detect='MOG2', mat_morph=mm, post_filter=PF).initialize()
for f in clip[:36]:
    g.detect(f)
t0 = time.perf_counter()
for f in clip:
    g.detect(f)
print("plain: %.1f us per detect()" % ((time.perf_counter() - t0) / 72 * 1e6))
hp = g._hot if hasattr(g, "_hot") else None
pr = cProfile.Profile()
pr.enable()
for _ in range(4):
    for f in clip:
        g.detect(f)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(18)
