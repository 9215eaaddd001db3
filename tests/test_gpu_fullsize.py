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


# ----------------------------------------------------------------------------------------------
# The BENCHMARKED entry points at the benchmarked shapes: clips through
# gofpid_pipeline_process_host_sequence / _device_sequence with G frames per MOG2 launch -- whole
# 64x8 tiles at 1920x1080 and 3840x2160, so this is mog2_kernel<1, FULL=true> carried through frame
# groups with the next frame's words prefetched in registers -- against live cv2: EVERY frame's mask
# and blob list, and the final model planes against the C restatement of MOG2.
def _run_clip(H, W, morph_cv, morph_hp, G, via, n_chunks, seed, n_obj, C=1):
    import torch
    from oracle import cport
    from pygofpid_b200 import HotPath
    from pygofpid_b200.hotpath import blobs_from_records
    depth = 16 if G == 8 else 12
    hp = HotPath((H, W) if C == 1 else (H, W, C), blur=("gaussian", (3, 3), 4.0), detector="MOG2", morph=morph_hp,
                 area_min=100.0, depth=depth, mog2_group=G)
    assert hp.group == G and hp.depth == depth
    ref = cv_path.CvHotPath(frg_detect="MOG2", mat_morph=morph_cv)
    ora = cport.MOG2(H, W, C)
    clip = SyntheticClip(H, W, C, seed=seed, n_obj=n_obj)
    L = depth - depth % G  # frames per clip call: whole groups, all still selectable afterwards
    t0 = 0
    for chunk in range(n_chunks):
        frames = np.stack(clip.frames(L))
        want = [ref.step(f) for f in frames]
        for f in frames:
            ora.apply(ref.blurred(f))
        src = torch.from_numpy(frames).cuda() if via == "device" else (
            torch.from_numpy(frames).pin_memory() if via == "pinned" else frames)
        counts, recs, pts = hp.process_sequence(src)
        ob = op = 0
        for t in range(L):
            nb, npt = int(counts[t, 0, 0]), int(counts[t, 0, 1])
            got = blobs_from_records(recs[ob:ob + nb], pts[op:op + npt])
            ob += nb; op += npt
            wm, wb = want[t]
            assert len(got) == len(wb), (t0 + t, len(got), len(wb))
            for g, w in zip(got, wb):
                assert np.array_equal(g["contour"], w["contour"]), t0 + t
                assert np.array_equal(g["center"], w["center"]) and np.array_equal(g["bottom"], w["bottom"]), t0 + t
            hp.select(L - 1 - t)
            gm = hp.mask()
            assert np.array_equal(gm, wm), (t0 + t, int((gm != wm).sum()))
        hp.select(0)
        t0 += L
    w, v, m, k = hp.mog2_model()
    ow, ov, om, ok = ora.planes()
    assert np.array_equal(k, ok)
    live = np.arange(5)[None, None, :] < ok[..., None]
    assert np.array_equal(w[live], ow[live]) and np.array_equal(v[live], ov[live]) and np.array_equal(m[live], om[live])
    hp.close()
    return t0


@pytest.mark.parametrize("G,via", [(3, "device"), (3, "pinned"), (8, "device"), (8, "host")])
def test_1080p_headline_clip_entry_points_match_cv2(G, via):
    se = cv.getStructuringElement(cv.MORPH_RECT, (5, 5))
    chain = [cv.erode, cv.dilate, cv.dilate, cv.erode]
    n = _run_clip(1080, 1920, [dict(fun=f, kernel=se) for f in chain],
                  [("erode" if f is cv.erode else "dilate", se) for f in chain], G, via, n_chunks=4, seed=2, n_obj=12)
    assert n >= 40


@pytest.mark.parametrize("G,via", [(3, "device"), (8, "pinned")])
def test_4k_clip_entry_points_match_cv2(G, via):
    se = cv.getStructuringElement(cv.MORPH_ELLIPSE, (15, 15))
    n = _run_clip(2160, 3840, [dict(fun=cv.erode, kernel=se), dict(fun=cv.dilate, kernel=se)],
                  [("erode", se), ("dilate", se)], G, via, n_chunks=3, seed=3, n_obj=16)
    assert n >= 36


def test_1080p_bgr_clip_entry_point_matches_cv2():
    se = cv.getStructuringElement(cv.MORPH_RECT, (5, 5))
    _run_clip(1080, 1920, [dict(fun=cv.erode, kernel=se), dict(fun=cv.dilate, kernel=se)],
              [("erode", se), ("dilate", se)], 3, "device", n_chunks=2, seed=2, n_obj=12, C=3)


def test_1080p_noise_without_morphology_through_the_dropin():
    """a noisy sensor with mat_morph=None at 1080p -- tens of thousands of components per
    frame, far beyond the per-slot capacities -- must give the reference's blob list, not an exception."""
    from pygofpid_b200 import GOFPID
    rs = np.random.RandomState(11)
    H, W = 1080, 1920
    g = GOFPID(frg_detect='FD', mat_morph=None, blur=None).initialize()
    ref = cv_path.CvHotPath(blur=None, frg_detect="FD", mat_morph=None)
    base = rs.randint(0, 256, (H, W)).astype(np.uint8)
    for t in range(4):
        f = base.copy()
        flip = rs.rand(H, W) < (0.30 if t != 2 else 0.45)  # t == 2: beyond the 8-connected percolation threshold
        f[flip] = f[flip] + 100  # wraps: |f - prev| is mostly > 5 where flipped
        wm, wb = ref.step(f)
        g.detect(f)
        assert np.array_equal(g.foreground_mask_, wm)
        assert len(g.blobs_) == len(wb), (t, len(g.blobs_), len(wb))
        for a, b in zip(g.blobs_, wb):
            assert np.array_equal(a["contour"], b["contour"]), t
            assert np.array_equal(a["center"], b["center"]) and np.array_equal(a["bottom"], b["bottom"]), t
        base = f


# ----------------------------------------------------------------------------------------------
# ... and the other end of the size range: frames smaller than one tile, one row, one pixel, through the
# whole chain (blur, detector, open 3x3, blobs) against live cv2 -- frames alternate between a flat image and
# noise, so nearly every pixel changes class every frame.
@pytest.mark.parametrize("det", ["MOG2", "KNN", "FD"])
@pytest.mark.parametrize("shape", [(1, 1), (1, 9), (3, 4), (7, 8, 3), (2, 8), (5, 3, 3), (9, 16), (4, 4, 3)])
def test_tiny_frames_match_cv2(shape, det):
    from pygofpid_b200 import HotPath
    se = cv.getStructuringElement(cv.MORPH_RECT, (3, 3))
    cv.setRNGSeed(7)
    ref = cv_path.CvHotPath(frg_detect=det, mat_morph=[dict(fun=cv.erode, kernel=se), dict(fun=cv.dilate, kernel=se)],
                            area_min=1.0)
    hp = HotPath(shape, blur=("gaussian", (3, 3), 4.0), detector=det, morph=[("erode", se), ("dilate", se)],
                 area_min=1.0, knn_seed=7, fd_threshold=5)
    rng = np.random.default_rng(1)
    for t in range(8):
        f = rng.integers(0, 256, size=shape, dtype=np.uint8) if t % 2 else np.full(shape, 40, np.uint8)
        want_mask, want_blobs = ref.step(f)
        hp.process(f)
        got_mask = hp.mask()
        assert np.array_equal(got_mask, want_mask), (t, got_mask, want_mask)
        got = hp.blobs()
        assert len(got) == len(want_blobs), (t, len(got), len(want_blobs))
        for g, w in zip(got, want_blobs):
            assert np.array_equal(g["contour"], w["contour"]), t
            assert np.array_equal(g["center"], w["center"]) and np.array_equal(g["bottom"], w["bottom"]), t
    hp.close()
