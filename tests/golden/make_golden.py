#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the REAL reference (qbarthelemy/PyGOFPID, imported
from /root/reference) on small seeded synthetic clips.  Run in the dev container only:

    python tests/golden/make_golden.py

The GPU box has no /root/reference; the tests read the committed .npz files.  Each file holds,
per frame: the reference's foreground mask (bit-packed by value), its blob list (contours,
centres, bottoms) and the alarm output of GOFPID.detect, plus the config needed to replay it.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import cv2 as cv  # noqa: E402
from pygofpid.pipeline import GOFPID  # noqa: E402  (the reference)
from pygofpid.segmentation import FrameDifferencing  # noqa: E402

from pygofpid_b200.synth import SyntheticClip  # noqa: E402

POST = dict(perimeter=None, anchor='bottom', perspective=None, perspective_coeff=0.5, presence_min=3,
            distance_min=0.25)

CASES = {
    # name: (H, W, C, n_frames, clip kwargs, GOFPID kwargs builder)
    "fd_gray_open3": dict(H=60, W=80, C=1, n=12, clip=dict(seed=1, n_obj=3),
                          kw=lambda: dict(frg_detect='FD',
                                          mat_morph=[dict(fun=cv.erode, kernel=np.ones((3, 3), np.uint8)),
                                                     dict(fun=cv.dilate, kernel=np.ones((3, 3), np.uint8))])),
    "mog2_gray_default": dict(H=72, W=96, C=1, n=40, clip=dict(seed=2, n_obj=3, obj_size=(12, 30, 12, 30)),
                              kw=lambda: dict(frg_detect='MOG2')),
    "mog2_bgr_default": dict(H=64, W=80, C=3, n=30, clip=dict(seed=3, n_obj=3, obj_size=(12, 30, 12, 30)),
                             kw=lambda: dict(frg_detect='MOG2')),
    "knn_gray_default": dict(H=64, W=80, C=1, n=40, clip=dict(seed=4, n_obj=3, obj_size=(12, 30, 12, 30)),
                             kw=lambda: dict(frg_detect='KNN'), rng_seed=4242),
    "mog2_gray_box_openclose": dict(H=70, W=90, C=1, n=30, clip=dict(seed=5, n_obj=4, obj_size=(12, 30, 12, 30)),
                                    kw=lambda: dict(frg_detect='MOG2', blur=dict(fun=cv.blur, ksize=(3, 3)),
                                                    mat_morph=[dict(fun=f, kernel=cv.getStructuringElement(cv.MORPH_ELLIPSE, (5, 5)))
                                                               for f in (cv.erode, cv.dilate, cv.dilate, cv.erode)])),
    "mog2_gray_convert_bgr": dict(H=64, W=80, C=3, n=25, clip=dict(seed=6, n_obj=3, obj_size=(12, 30, 12, 30)),
                                  kw=lambda: dict(frg_detect='MOG2', convert=cv.COLOR_BGR2GRAY)),
}


def pack_contours(blobs):
    if not blobs:
        return np.zeros((0, 2), np.int32), np.zeros(0, np.int32)
    pts = np.concatenate([b['contour'].reshape(-1, 2) for b in blobs]).astype(np.int32)
    lens = np.array([len(b['contour']) for b in blobs], np.int32)
    return pts, lens


def main():
    for name, case in CASES.items():
        if case.get("rng_seed") is not None:
            cv.setRNGSeed(case["rng_seed"])
        post = dict(POST)
        g = GOFPID(post_filter=post, **case["kw"]()).initialize()
        clip = SyntheticClip(case["H"], case["W"], case["C"], **case["clip"])
        out = dict(shape=np.array([case["H"], case["W"], case["C"]]), n=case["n"],
                   rng_seed=np.array(case.get("rng_seed") or 0))
        alarms, ntracked = [], []
        for t in range(case["n"]):
            f = clip.next()
            y = g.detect(f)
            alarms.append(bool(y))
            ntracked.append(len(g.tracked_blobs_))
            out["mask_%03d" % t] = g.foreground_mask_.copy()
            pts, lens = pack_contours(g.blobs_)
            out["pts_%03d" % t] = pts
            out["lens_%03d" % t] = lens
            out["center_%03d" % t] = np.array([b['center'] for b in g.blobs_], np.int16).reshape(-1, 2)
            out["bottom_%03d" % t] = np.array([b['bottom'] for b in g.blobs_], np.int16).reshape(-1, 2)
        out["alarm"] = np.array(alarms)
        out["ntracked"] = np.array(ntracked)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "frames", case["n"], "blobs/frame", np.mean([len(out["lens_%03d" % t]) for t in range(case["n"])]),
              "alarms", int(np.sum(alarms)))
    # stand-alone FrameDifferencing of the reference (segmentation.py:21-70), 3-channel quirk included
    rs = np.random.RandomState(7)
    fd_out = {}
    for tag, shape in (("g", (20, 30)), ("g1", (20, 30, 1)), ("c", (20, 30, 3))):
        fd = FrameDifferencing(threshold=7.5)
        frames = rs.randint(0, 256, (4,) + shape).astype(np.uint8)
        fd_out["frames_" + tag] = frames
        fd_out["masks_" + tag] = np.stack([fd.apply(f) for f in frames])
    np.savez_compressed(os.path.join(HERE, "framediff.npz"), **fd_out)
    print("framediff ok; cv2", cv.__version__)


if __name__ == "__main__":
    main()
