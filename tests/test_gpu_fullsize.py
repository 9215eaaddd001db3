"""Full-size differential tests: the BASELINE.json configurations at their real frame sizes, CUDA
path (through the C ABI) against the reference's own sequence of cv2 calls (oracle/cv_path.py:
GOFPID.detect lines pipeline.py:461-479 + _find_blob / _create_blob) on the same seeded frames.
Bit-exact masks; identical blob lists (contours, order, centre, bottom).
"""
import numpy as np
import pytest

from oracle import cv_path
from pygofpid_b200.synth import SyntheticClip

pytestmark = pytest.mark.gpu
cv = cv_path.cv


def _run(H, W, C, detector, morph_cv, morph_hp, n_frames, seed, n_obj, obj_size=None, check_blobs_every=1,
         knn_seed=0):
    from pygofpid_b200 import HotPath
    hp = HotPath((H, W) if C == 1 else (H, W, C), blur=("gaussian", (3, 3), 4.0), detector=detector,
                 morph=morph_hp, area_min=100.0, knn_seed=knn_seed)
    if detector == "KNN":
        cv.setRNGSeed(knn_seed)
    ref = cv_path.CvHotPath(frg_detect=detector, mat_morph=morph_cv)
    clip = SyntheticClip(H, W, C, seed=seed, n_obj=n_obj, obj_size=obj_size)
    for t in range(n_frames):
        f = clip.next()
        want_mask, want_blobs = ref.step(f)
        hp.process(f)
        got_mask = hp.mask()
        assert np.array_equal(got_mask, want_mask), (t, int((got_mask != want_mask).sum()))
        if t % check_blobs_every == 0:
            got = hp.blobs()
            assert len(got) == len(want_blobs), (t, len(got), len(want_blobs))
            for g, w in zip(got, want_blobs):
                assert np.array_equal(g["contour"], w["contour"]), t
                assert np.array_equal(g["center"], w["center"]) and np.array_equal(g["bottom"], w["bottom"]), t
    hp.close()


def test_1080p_mog2_open_close_blobs_matches_cv2():
    """BASELINE configs[1], the headline: 1080p, Gaussian 3x3, MOG2, open + close 5x5, blobs."""
    se = cv.getStructuringElement(cv.MORPH_RECT, (5, 5))
    chain = [cv.erode, cv.dilate, cv.dilate, cv.erode]
    _run(1080, 1920, 1, "MOG2", [dict(fun=f, kernel=se) for f in chain],
         [("erode" if f is cv.erode else "dilate", se) for f in chain], n_frames=40, seed=2, n_obj=12)


def test_1080p_bgr_mog2_matches_cv2():
    se = cv.getStructuringElement(cv.MORPH_RECT, (5, 5))
    _run(1080, 1920, 3, "MOG2", [dict(fun=cv.erode, kernel=se), dict(fun=cv.dilate, kernel=se)],
         [("erode", se), ("dilate", se)], n_frames=12, seed=2, n_obj=12)


def test_4k_mog2_ellipse15_open_matches_cv2():
    """BASELINE configs[2]: 3840x2160, 15x15 ellipse opening."""
    se = cv.getStructuringElement(cv.MORPH_ELLIPSE, (15, 15))
    _run(2160, 3840, 1, "MOG2", [dict(fun=cv.erode, kernel=se), dict(fun=cv.dilate, kernel=se)],
         [("erode", se), ("dilate", se)], n_frames=8, seed=3, n_obj=16)


def test_1080p_knn_many_blobs_matches_cv2():
    """BASELINE configs[4]: KNN at 1080p over ~400 small objects (labelling / tracing stress).  cv2's
    KNN draws from the calling thread's cv::RNG: both sides start from the same seed."""
    se = cv.getStructuringElement(cv.MORPH_RECT, (5, 5))
    _run(1080, 1920, 1, "KNN", [dict(fun=cv.erode, kernel=se), dict(fun=cv.dilate, kernel=se)],
         [("erode", se), ("dilate", se)], n_frames=16, seed=5, n_obj=400, obj_size=(12, 90, 8, 80),
         knn_seed=12345)


def test_ilids_720x576_bgr_defaults_matches_cv2():
    """BASELINE configs[3] (one of its streams): 720x576 BGR, all defaults."""
    se = cv.getStructuringElement(cv.MORPH_RECT, (5, 5))
    _run(576, 720, 3, "MOG2", [dict(fun=cv.erode, kernel=se), dict(fun=cv.dilate, kernel=se)],
         [("erode", se), ("dilate", se)], n_frames=30, seed=1000, n_obj=6)
