"""Generates tests/golden/tail_streams.npz: what the REAL reference's host tail (steps 5-7:
_track_blob, _post_filter, _detect_blob, _maintain_alarm; pipeline.py:590-799) answers on the
synthetic blob streams of tests/golden_cases.py (tail_stream).  Run in the build container only
(needs /root/reference):  python tests/golden/make_tail_golden.py"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
from golden_cases import (DISPLAY_FRAMES, DISPLAY_TRIALS, TAIL_FRAMES, TAIL_TRIALS, tail_post_filter,  # noqa: E402
                          tail_stream)
from pygofpid.pipeline import GOFPID  # noqa: E402  (the real reference)

out = {}
for trial in range(TAIL_TRIALS):
    H, W, kind, frames = tail_stream(trial)
    g = GOFPID(post_filter=tail_post_filter(kind)).initialize()
    g.input_shape_ = (H, W)
    g._calib_first_frame()
    alarm, ntracked, presence, nfiltered = [], [], [], []
    for t, blobs in enumerate(frames):
        g.blobs_ = blobs
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            g._track_blob()
            g._post_filter()
            y = g._maintain_alarm(g._detect_blob())
        alarm.append(bool(y))
        ntracked.append(len(g.tracked_blobs_))
        presence.append(sum(tr['presence'] for tr in g.tracked_blobs_))
        nfiltered.append(sum(len(tr['filtered']) for tr in g.tracked_blobs_))
        if trial in DISPLAY_TRIALS and t in DISPLAY_FRAMES:  # GOFPID.display draws into the frame in place
            canvas = np.zeros((H, W, 3), np.uint8)
            g.display(canvas, {'tracking': True, 'perspective': True})
            out["display_%d_%d" % (trial, t)] = canvas
    out["alarm_%d" % trial] = np.array(alarm)
    out["ntracked_%d" % trial] = np.array(ntracked, np.int32)
    out["presence_%d" % trial] = np.array(presence, np.int32)
    out["nfiltered_%d" % trial] = np.array(nfiltered, np.int32)
    assert len(alarm) == TAIL_FRAMES
np.savez_compressed(os.path.join(HERE, "tail_streams.npz"), **out)
print("wrote tail_streams.npz:", {k: int(v.sum()) for k, v in out.items() if k.startswith("alarm")})
