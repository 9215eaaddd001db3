"""GPU parity of the pipeline handle (through the C ABI) against the CPU oracle.

FD / MOG2 / KNN masks and model state are compared bit-for-bit with oracle/gofpid_oracle.c
(itself pinned to cv2 4.13 and to the reference-generated vectors in tests/golden/), far
inside the <= 0.1 % mask-disagreement bar BASELINE.json allows for the mixture models.
Blob lists (contours, order, centre / bottom) are compared with the oracle's restatement of
cv.findContours(RETR_EXTERNAL, CHAIN_APPROX_NONE) + contourArea + moments.
"""
import numpy as np
import pytest

from oracle import blobs as oblobs
from oracle import cport
from pygofpid_b200.synth import SyntheticClip

pytestmark = pytest.mark.gpu


def _hot(*a, **k):
    from pygofpid_b200 import HotPath
    return HotPath(*a, **k)


def _shape(H, W, C):
    return (H, W) if C == 1 else (H, W, C)


# ----------------------------------------------------------------------------- FD
@pytest.mark.parametrize("H,W,C", [(64, 50, 1), (20, 32, 3), (240, 320, 1), (33, 131, 3)])
@pytest.mark.parametrize("blur", [None, ("gaussian", (3, 3), 4.0), ("box", (3, 3))])
def test_fd_masks_and_state(H, W, C, blur):
    hp = _hot(_shape(H, W, C), blur=blur, detector="FD", morph=(), find_blobs=False, fd_threshold=7.5)
    fd = cport.FrameDifferencing(7.5)
    clip = SyntheticClip(H, W, C, seed=3, n_obj=3)
    for t in range(6):
        f = clip.next()
        b = f
        if blur and blur[0] == "gaussian":
            b = cport.gaussian_blur(f, blur[1], blur[2]).reshape(f.shape)
        elif blur:
            b = cport.box_blur(f, blur[1]).reshape(f.shape)
        want = fd.apply(b)
        hp.process(f)
        got = hp.mask()
        assert got.shape == (H, W)
        assert np.array_equal(got, want), (t, int((got != want).sum()))
        assert np.array_equal(hp.fd_model(), b)
    if C == 3:
        assert set(np.unique(got)) <= {0, 84, 85}


# ----------------------------------------------------------------------------- MOG2
@pytest.mark.parametrize("generic", [False, True])
@pytest.mark.parametrize("H,W,C,sigma", [(48, 64, 1, 6), (48, 64, 3, 6), (37, 53, 1, 25), (37, 53, 3, 25),
                                          (120, 200, 1, 4)])
def test_mog2_masks_and_model_bitexact(H, W, C, sigma, generic, monkeypatch):
    """Both MOG2 kernels against the oracle: the lean one (ordinary parameters: the default) and the generic
    one (GOFPID_MOG2_GENERIC=1 keeps it for ordinary parameters too)."""
    if generic:
        monkeypatch.setenv("GOFPID_MOG2_GENERIC", "1")
    hp = _hot(_shape(H, W, C), blur=None, detector="MOG2", morph=(), find_blobs=False)
    ora = cport.MOG2(H, W, C)
    clip = SyntheticClip(H, W, C, seed=10 * C + sigma, n_obj=4, noise_sigma=sigma)
    nframes = 300 if H * W < 5000 else 120
    for t in range(nframes):
        f = clip.next()
        want = ora.apply(f)
        hp.process(f)
        got = hp.mask()
        assert np.array_equal(got, want), (t, int((got != want).sum()))
        if t in (0, 1, 7, nframes - 1):
            w, v, m, k = hp.mog2_model()
            ow, ov, om, ok = ora.planes()
            assert np.array_equal(k, ok), t
            live = np.arange(5)[None, None, :] < ok[..., None]
            assert np.array_equal(w[live], ow[live]), t
            assert np.array_equal(v[live], ov[live]), t
            assert np.array_equal(m[live], om[live]), t
    assert set(np.unique(got)) <= {0, 127, 255}
    assert abs(hp.mog2_mean_modes() - ok.mean()) < 1e-9


@pytest.mark.parametrize("det,C", [("MOG2", 1), ("MOG2", 3), ("KNN", 1)])
@pytest.mark.parametrize("H,W", [(2, 8), (9, 68), (16, 64), (24, 132), (17, 260), (8, 12)])
def test_register_blur_on_border_tiles(det, C, H, W):
    """3x3 Gaussian evaluated in registers on EVERY tile of frames that are a whole number of words wide:
    reflected rows (H = 2: both neighbours of a row are the other row), reflected pixels at both borders, a
    last tile with a single valid lane (W = 68: that lane loads its left neighbour and reflects on the right),
    lanes and rows of a tile outside the frame -- masks of a few frames against the oracle on the oracle's blur."""
    blur = ("gaussian", (3, 3), 4.0)
    seed = 4242
    hp = _hot(_shape(H, W, C), blur=blur, detector=det, morph=(), find_blobs=False, knn_seed=seed)
    ora = cport.MOG2(H, W, C) if det == "MOG2" else cport.KNN(H, W, C, seed)
    rng = np.random.default_rng(H * 1000 + W + C)
    base = rng.integers(0, 256, size=_shape(H, W, C), dtype=np.uint8)
    for t in range(12):
        f = base.copy()
        if t % 3 == 2:  # a frame that differs everywhere: every pixel leaves the matched-background path
            f = rng.integers(0, 256, size=_shape(H, W, C), dtype=np.uint8)
        else:
            f = (f.astype(np.int16) + rng.integers(-3, 4, size=f.shape)).clip(0, 255).astype(np.uint8)
        b = cport.gaussian_blur(f, blur[1], blur[2]).reshape(f.shape)
        want = ora.apply(b)
        hp.process(f)
        got = hp.mask()
        assert np.array_equal(got, want), (det, C, H, W, t, int((got != want).sum()))
    hp.close()


@pytest.mark.parametrize("blur", [("gaussian", (3, 3), 4.0), ("gaussian", (3, 3), 0.0), ("box", (3, 3)),
                                  ("gaussian", (3, 3), 0.05)])
@pytest.mark.parametrize("H,W", [(40, 260), (41, 331)])
def test_mog2_fused_blur_interior_and_border_tiles(blur, H, W):
    """Frames wide and tall enough to have INTERIOR 64x8 tiles (3x3 blur evaluated in registers with
    dp4a) next to border tiles (general shared-memory path): blurred state and masks bit-exact.
    sigma 0.05 gives the taps [0, 256, 0], which do not fit a byte and must take the general path."""
    hp = _hot((H, W), blur=blur, detector="MOG2", morph=(), find_blobs=False)
    ora = cport.MOG2(H, W, 1)
    clip = SyntheticClip(H, W, 1, seed=77, n_obj=4, noise_sigma=5)
    for t in range(40):
        f = clip.next()
        b = cport.gaussian_blur(f, blur[1], blur[2]) if blur[0] == "gaussian" else cport.box_blur(f, blur[1])
        want = ora.apply(b)
        hp.process(f)
        got = hp.mask()
        assert np.array_equal(got, want), (blur, t, int((got != want).sum()))
    w, v, m, k = hp.mog2_model()
    ow, ov, om, ok = ora.planes()
    live = np.arange(5)[None, None, :] < ok[..., None]
    assert np.array_equal(k, ok) and np.array_equal(m[live], om[live]) and np.array_equal(v[live], ov[live])
    hp.close()


def test_mog2_first_frame_is_shadow_everywhere():
    hp = _hot((32, 40), blur=None, detector="MOG2", find_blobs=False)
    f = np.full((32, 40), 90, np.uint8)
    f[3, 4] = 0
    hp.process(f)
    m = hp.mask()
    assert m[3, 4] == 255 and (np.delete(m.ravel(), 3 * 40 + 4) == 127).all()
    hp.process(f)
    assert (hp.mask() == 0).all()
    hp.reset()
    hp.process(f)
    assert np.array_equal(hp.mask(), m)


# ----------------------------------------------------------------------------- KNN
@pytest.mark.parametrize("kernel", ["default", "scalar"])
@pytest.mark.parametrize("H,W,C", [(37, 53, 1), (37, 53, 3), (50, 66, 1), (64, 128, 1), (40, 258, 1)])
def test_knn_masks_and_state_bitexact(H, W, C, kernel, monkeypatch):
    """Gray frames run the byte-SIMD kernel (knn_gray_kernel) by default; kernel="scalar" forces the
    general per-pixel kernel, which colour frames always use."""
    if kernel == "scalar":
        if C == 3:
            pytest.skip("colour frames always run the general kernel")
        monkeypatch.setenv("GOFPID_KNN_SCALAR", "1")
    seed = 12345
    hp = _hot(_shape(H, W, C), blur=None, detector="KNN", morph=(), find_blobs=False, knn_seed=seed)
    ora = cport.KNN(H, W, C, seed)
    clip = SyntheticClip(H, W, C, seed=C, n_obj=4)
    for t in range(200):
        f = clip.next()
        want = ora.apply(f)
        hp.process(f)
        got = hp.mask()
        assert np.array_equal(got, want), (t, int((got != want).sum()))
        if t in (0, 5, 60, 199):
            s, idx, nxt, cnt, rng = hp.knn_model()
            os_, oidx, onxt, ocnt, orng = ora.state()
            assert np.array_equal(s, os_), t
            assert np.array_equal(idx, oidx), t
            assert np.array_equal(nxt, onxt), t
            assert list(cnt) == list(ocnt), t
            assert rng == orng, t


# ----------------------------------------------------------------------------- fused chain
@pytest.mark.parametrize("det,C", [("MOG2", 1), ("MOG2", 3), ("KNN", 1), ("FD", 1)])
def test_fused_blur_detect_morph(det, C):
    H, W = 90, 150
    r5 = np.ones((5, 5), np.uint8)
    morph = [("erode", r5), ("dilate", r5)]
    hp = _hot(_shape(H, W, C), blur=("gaussian", (3, 3), 4.0), detector=det, morph=morph,
              find_blobs=False, knn_seed=99)
    ora = {"MOG2": lambda: cport.MOG2(H, W, C), "KNN": lambda: cport.KNN(H, W, C, 99),
           "FD": lambda: cport.FrameDifferencing(5)}[det]()
    clip = SyntheticClip(H, W, C, seed=21, n_obj=5)
    for t in range(40):
        f = clip.next()
        b = cport.gaussian_blur(f, (3, 3), 4.0).reshape(f.shape)
        want = cport.dilate(cport.erode(ora.apply(b), r5), r5)
        hp.process(f)
        got = hp.mask()
        assert np.array_equal(got, want), (det, t, int((got != want).sum()))


# ----------------------------------------------------------------------------- colour -> gray
@pytest.mark.parametrize("code", ["bgr2gray", "rgb2gray"])
@pytest.mark.parametrize("H,W", [(37, 53), (64, 130), (9, 257)])
@pytest.mark.parametrize("blur", [None, ("gaussian", (3, 3), 4.0), ("gaussian", (7, 7), 0.0), ("box", (5, 5))])
def test_fused_gray_conversion(code, H, W, blur):
    """cv.cvtColor(COLOR_BGR2GRAY / COLOR_RGB2GRAY) (pipeline.py:458-459) fused into the frame load:
    masks and state equal those of the oracle fed with cv2's gray image."""
    import cv2
    cvcode = cv2.COLOR_BGR2GRAY if code == "bgr2gray" else cv2.COLOR_RGB2GRAY
    rs = np.random.RandomState(17)
    hp = _hot((H, W, 3), blur=blur, detector="FD", find_blobs=False, fd_threshold=5, convert=code)
    fd = cport.FrameDifferencing(5)
    base = rs.randint(0, 256, (H, W, 3)).astype(np.uint8)
    for t in range(5):
        f = np.clip(base.astype(np.int16) + rs.randint(-9, 10, base.shape), 0, 255).astype(np.uint8)
        g = cv2.cvtColor(f, cvcode)
        if blur is None:
            b = g
        elif blur[0] == "gaussian":
            b = cport.gaussian_blur(g, blur[1], blur[2])
        else:
            b = cport.box_blur(g, blur[1])
        want = fd.apply(b)
        hp.process(f)
        assert np.array_equal(hp.mask(), want), (code, blur, t)
        assert np.array_equal(hp.fd_model(), b), (code, blur, t)
    hp.close()


@pytest.mark.parametrize("det", ["MOG2", "KNN"])
def test_fused_gray_conversion_mixture_models(det):
    import cv2
    H, W = 48, 70
    hp = _hot((H, W, 3), blur=("gaussian", (3, 3), 4.0), detector=det, find_blobs=False, knn_seed=5,
              convert="bgr2gray")
    ora = cport.MOG2(H, W, 1) if det == "MOG2" else cport.KNN(H, W, 1, 5)
    clip = SyntheticClip(H, W, 3, seed=4, n_obj=3)
    for t in range(30):
        f = clip.next()
        want = ora.apply(cport.gaussian_blur(cv2.cvtColor(f, cv2.COLOR_BGR2GRAY), (3, 3), 4.0))
        hp.process(f)
        assert np.array_equal(hp.mask(), want), (det, t)
    hp.close()


# ----------------------------------------------------------------------------- blobs
def _blob_masks():
    import cv2
    rs = np.random.RandomState(5)
    out = []
    for it in range(24):
        H, W = rs.randint(20, 140), rs.randint(20, 300)
        if it % 4 == 0:
            M = ((rs.rand(H, W) < 0.5) * 255).astype(np.uint8)
        else:
            M = np.zeros((H, W), np.uint8)
            for _ in range(10):
                x, y = rs.randint(0, W), rs.randint(0, H)
                w, h = rs.randint(2, 60), rs.randint(2, 60)
                cv2.rectangle(M, (x, y), (x + w, y + h), int(rs.choice([127, 255])),
                              int(rs.choice([-1, 1, 2])))
            M[rs.rand(H, W) < 0.03] = 0
        out.append(M)
    # nested rings: only the outermost component is external
    M = np.zeros((60, 60), np.uint8)
    cv2.rectangle(M, (2, 2), (57, 57), 255, 2)
    cv2.rectangle(M, (10, 10), (49, 49), 255, 2)
    cv2.rectangle(M, (20, 20), (39, 39), 255, -1)
    out.append(M)
    out.append(np.full((40, 70), 255, np.uint8))
    out.append(np.zeros((40, 70), np.uint8))
    return out


@pytest.mark.parametrize("runs_cap", [None, 40])
def test_blobs_match_reference_semantics(runs_cap, monkeypatch):
    """Feed masks through FD (|mask - 0| > 0 keeps them as they are) and compare the blob list.
    runs_cap=40 sends every frame with more than 40 runs through the tiled labelling fallback."""
    if runs_cap is not None:
        monkeypatch.setenv("GOFPID_RUNS_CAP", str(runs_cap))
    for M in _blob_masks():
        H, W = M.shape
        for area_min in (100.0, 0.5):
            hp = _hot((H, W), blur=None, detector="FD", morph=(), find_blobs=True, fd_threshold=0,
                      area_min=area_min)
            hp.process(np.zeros((H, W), np.uint8))
            hp.process(M)
            got_mask = hp.mask()
            want_mask = np.where(M > 0, 255, 0).astype(np.uint8)
            assert np.array_equal(got_mask, want_mask)
            # labels and per-component statistics
            want_labels = cport.ccl8(want_mask)
            assert np.array_equal(hp.labels(), want_labels)
            assert hp.components() == cport.component_stats(want_labels)
            want = oblobs.blobs_from_mask(want_mask, area_min=area_min)
            got = hp.blobs()
            assert len(got) == len(want), (M.shape, area_min, len(got), len(want))
            for g, w in zip(got, want):
                assert g["contour"].dtype == np.int32 and g["contour"].shape == w["contour"].shape
                assert np.array_equal(g["contour"], w["contour"])
                assert np.array_equal(g["bottom"], w["bottom"]) and g["bottom"].dtype == np.int16
                assert np.array_equal(g["center"], w["center"])
                assert np.array_equal(g["anchor"], w["anchor"])
            hp.close()


@pytest.mark.parametrize("runs_cap", [None, 64])
def test_device_clip_blob_stage_matches_reference_semantics(runs_cap, monkeypatch):
    """The blob stage as device-resident clips run it (stage-B graphs replayed per frame slot, the tiled
    fallback behind a conditional graph node): the masks of test_blobs_match_reference_semantics -- noise,
    rings, nesting, full and empty frames -- plus large blobs with long ragged borders (more visits than the
    tracer has checkpoints, so borders are walked in thinned segments), one clip per shape through
    process_sequence on a CUDA tensor; every frame's blob list, the last frame's labels and statistics
    against the oracle.  runs_cap=64: busy frames raise the conditional node in between calm ones."""
    import cv2
    import torch
    from pygofpid_b200.hotpath import blobs_from_records
    if runs_cap is not None:
        monkeypatch.setenv("GOFPID_RUNS_CAP", str(runs_cap))
    rs = np.random.RandomState(21)
    clips = {}
    for M in _blob_masks():
        clips.setdefault(M.shape, []).append(M)
    big = []
    for it in range(6):  # blobs with borders far longer than 512 visits, holes, thin bridges
        M = np.zeros((300, 420), np.uint8)
        for _ in range(5):
            c = (int(rs.randint(40, 380)), int(rs.randint(40, 260)))
            cv2.ellipse(M, c, (int(rs.randint(20, 120)), int(rs.randint(15, 90))), float(rs.randint(0, 180)), 0, 360, 255, -1)
        M[rs.rand(*M.shape) < 0.08] = 0
        if it % 2:
            M = cv2.dilate(M, np.ones((3, 3), np.uint8))
        big.append(M)
    clips[(300, 420)] = big
    n_checked = 0
    for (H, W), masks in clips.items():
        for area_min in (100.0, 0.5):
            hp = _hot((H, W), blur=None, detector="FD", morph=(), find_blobs=True, fd_threshold=0, area_min=area_min, depth=2)
            # FD yields |frame - previous| > 0: frames whose differences are the wanted masks; two empty
            # frames first (the first `depth` steps of a handle run eagerly, graphs replay from then on)
            seq_masks = [np.zeros((H, W), np.uint8)] * 3 + list(masks)
            frames, prev = [], np.zeros((H, W), np.uint8)
            for M in seq_masks:
                prev = np.where(M > 0, 255 - prev, prev).astype(np.uint8)
                frames.append(prev)
            counts, recs, pts = hp.process_sequence(torch.from_numpy(np.stack(frames)).cuda())
            ob = op = 0
            for t, M in enumerate(seq_masks):
                nb, npt = int(counts[t, 0, 0]), int(counts[t, 0, 1])
                got = blobs_from_records(recs[ob:ob + nb], pts[op:op + npt])
                ob += nb; op += npt
                want = oblobs.blobs_from_mask(np.where(M > 0, 255, 0).astype(np.uint8), area_min=area_min)
                assert len(got) == len(want), ((H, W), t, area_min, len(got), len(want))
                for g, w in zip(got, want):
                    assert np.array_equal(g["contour"], w["contour"]), ((H, W), t, area_min)
                    assert np.array_equal(g["center"], w["center"]) and np.array_equal(g["bottom"], w["bottom"])
                n_checked += 1
            want_mask = np.where(seq_masks[-1] > 0, 255, 0).astype(np.uint8)
            assert np.array_equal(hp.mask(), want_mask)
            assert np.array_equal(hp.labels(), cport.ccl8(want_mask))
            assert hp.components() == cport.component_stats(cport.ccl8(want_mask))
            hp.close()
    assert n_checked > 50


@pytest.mark.parametrize("no_cond", ["0", "1"])
def test_graph_replay_switches_between_run_and_tiled_labelling(no_cond, monkeypatch):
    """Frames alternate between a few blobs (run-based labelling) and dense noise (more runs than the
    capacity -> tiled fallback) while the step is replayed as a CUDA graph: the fallback kernels sit
    in a conditional node (no_cond=1: plain early-exit launches) that must fire exactly when needed."""
    import cv2
    monkeypatch.setenv("GOFPID_RUNS_CAP", "200")
    monkeypatch.setenv("GOFPID_NO_COND", no_cond)
    rs = np.random.RandomState(8)
    H, W = 120, 200
    hp = _hot((H, W), blur=None, detector="FD", morph=(), find_blobs=True, fd_threshold=0, area_min=4.0)
    prev = np.zeros((H, W), np.uint8)
    hp.process(prev)
    for t in range(12):
        if t % 3 == 2:
            M = ((rs.rand(H, W) < 0.4) * 255).astype(np.uint8)      # thousands of runs
        else:
            M = np.zeros((H, W), np.uint8)
            for _ in range(5):
                x, y = rs.randint(0, W - 30), rs.randint(0, H - 30)
                cv2.rectangle(M, (x, y), (x + rs.randint(5, 30), y + rs.randint(5, 30)), 255, -1)
        # FD gives |frame - previous| > 0: feed frames whose difference is the wanted mask
        frame = np.where(M > 0, 255 - prev, prev).astype(np.uint8)
        hp.process(frame)
        want_mask = np.where(frame != prev, 255, 0).astype(np.uint8)
        prev = frame
        assert np.array_equal(hp.mask(), want_mask)
        want = oblobs.blobs_from_mask(want_mask, area_min=4.0)
        got = hp.blobs()
        assert len(got) == len(want), (t, len(got), len(want))
        for g, w in zip(got, want):
            assert np.array_equal(g["contour"], w["contour"])
        want_labels = cport.ccl8(want_mask)
        assert np.array_equal(hp.labels(), want_labels), t
        assert hp.components() == cport.component_stats(want_labels), t
    hp.close()


def test_busy_frame_beyond_the_prefetched_result_prefix():
    """More kept blobs (> 256) and contour points (> 4096) than travel to the host with the step:
    the remainder is fetched on demand and the blob list is still complete and in order."""
    H, W = 200, 300
    M = np.zeros((H, W), np.uint8)
    for y in range(2, H - 6, 8):
        for x in range(2, W - 6, 8):
            M[y:y + 6, x:x + 6] = 255          # 24 x 37 = 888 squares of 20 contour points
    hp = _hot((H, W), blur=None, detector="FD", morph=(), find_blobs=True, fd_threshold=0, area_min=4.0)
    hp.process(np.zeros((H, W), np.uint8))
    for _ in range(6):                          # eager steps and graph replays alike
        hp.process(M)
        hp.process(np.zeros((H, W), np.uint8))
    hp.process(M)
    want = oblobs.blobs_from_mask(M, area_min=4.0)
    got = hp.blobs()
    assert len(want) > 256 and sum(len(w["contour"]) for w in want) > 4096
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert np.array_equal(g["contour"], w["contour"])
        assert np.array_equal(g["center"], w["center"])
    hp.close()


def test_sequence_call_equals_frame_by_frame_calls():
    """gofpid_pipeline_process_host_sequence (a clip in one call, frames in flight) gives the blob
    lists of the frame-by-frame calls, frame for frame."""
    H, W = 96, 160
    r3 = np.ones((3, 3), np.uint8)
    kw = dict(blur=("gaussian", (3, 3), 4.0), detector="MOG2", morph=[("erode", r3), ("dilate", r3)], area_min=20.0)
    clip = np.stack(SyntheticClip(H, W, 1, seed=3, n_obj=5).frames(25))
    a, b = _hot((H, W), **kw), _hot((H, W), **kw)
    counts, recs, pts = a.process_sequence(clip)
    assert counts.shape == (25, 1, 2)
    ob = op = 0
    for t in range(25):
        b.process(clip[t])
        want_recs, want_pts = b.raw_blobs()
        nb, npt = counts[t, 0]
        assert nb == len(want_recs) and npt == len(want_pts), t
        assert np.array_equal(pts[op:op + npt], want_pts), t
        for k in range(nb):
            for f in ("first", "x0", "y0", "x1", "y1", "area_px", "n_points", "point_offset", "a00", "a10", "a01"):
                assert recs[ob + k][f] == getattr(want_recs[k], f), (t, k, f)
        ob += nb; op += npt
    assert ob == len(recs) and op == len(pts)
    a.close(); b.close()


@pytest.mark.parametrize("group", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("H,W,C,sigma,blur", [(40, 260, 1, 6, ("gaussian", (3, 3), 4.0)), (37, 53, 3, 25, None),
                                              (41, 331, 1, 25, ("box", (3, 3))), (24, 200, 3, 6, ("gaussian", (5, 5), 4.0)),
                                              (48, 320, 1, 25, ("gaussian", (3, 3), 4.0)), (32, 192, 3, 6, ("gaussian", (3, 3), 4.0)),
                                              (48, 320, 1, 25, "generic")])
def test_mog2_frame_groups_bitexact(group, H, W, C, sigma, blur, monkeypatch):
    """Temporal blocking: the clip entry point carries every tile through `group` frames per MOG2
    launch (modes resident in shared memory).  Masks of EVERY frame and the model state stay
    bit-exact against the oracle, through the early learning-rate schedule (n < 250) and beyond."""
    import torch
    monkeypatch.setenv("GOFPID_MOG2_GROUP", str(group))
    monkeypatch.setenv("GOFPID_DEPTH", "16")
    if blur == "generic":  # whole tiles (48 x 320: the FULL instantiation, interior tiles) on the generic kernel
        monkeypatch.setenv("GOFPID_MOG2_GENERIC", "1")
        blur = ("gaussian", (3, 3), 4.0)
    hp = _hot(_shape(H, W, C), blur=blur, detector="MOG2", morph=(), find_blobs=False)
    assert hp.group == group and hp.depth == 16
    ora = cport.MOG2(H, W, C)
    clip = SyntheticClip(H, W, C, seed=7 * C + sigma, n_obj=4, noise_sigma=sigma)
    sh = torch.cuda.current_stream().cuda_stream
    row = W * C
    done = 0
    for chunk in (16, 11, 8, 1, 13, 16, 16, 16, 16, 7, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 5, 16):
        frames = np.stack(clip.frames(chunk))
        want = []
        for f in frames:
            g = f
            if blur is not None:
                g = cport.gaussian_blur(f, blur[1], blur[2]) if blur[0] == "gaussian" else cport.box_blur(f, blur[1])
            want.append(ora.apply(g))
        dev = torch.from_numpy(frames).cuda()
        hp.enqueue_device_sequence(dev.data_ptr(), chunk, row * H, row, row * H, sh)
        for age in range(chunk):
            hp.select(age)
            got = hp.mask()
            assert np.array_equal(got, want[chunk - 1 - age]), (done, chunk, age, int((got != want[chunk - 1 - age]).sum()))
        hp.select(0)
        done += chunk
        w, v, m, k = hp.mog2_model()
        ow, ov, om, ok = ora.planes()
        assert np.array_equal(k, ok), done
        live = np.arange(5)[None, None, :] < ok[..., None]
        assert np.array_equal(w[live], ow[live]) and np.array_equal(v[live], ov[live]) and np.array_equal(m[live], om[live]), done
    assert done > 300
    hp.close()


def test_device_clip_equals_host_clip_and_streams_stay_independent():
    """process_sequence on a CUDA tensor (gofpid_pipeline_process_device_sequence) returns what the
    host clip call returns, with several streams per handle and frame groups."""
    import torch
    H, W, S, T = 72, 136, 3, 41
    r3 = np.ones((3, 3), np.uint8)
    kw = dict(blur=("gaussian", (3, 3), 4.0), detector="MOG2", morph=[("erode", r3), ("dilate", r3)], area_min=20.0,
              n_streams=S)
    clips = [SyntheticClip(H, W, 1, seed=50 + i, n_obj=4).frames(T) for i in range(S)]
    clip = np.stack([np.stack(c) for c in clips], axis=1)  # (T, S, H, W)
    a, b = _hot((H, W), **kw), _hot((H, W), **kw)
    ca, ra, pa = a.process_sequence(clip)
    cb, rb, pb = b.process_sequence(torch.from_numpy(clip).cuda())
    assert np.array_equal(ca, cb) and np.array_equal(pa, pb) and ra.tobytes() == rb.tobytes()
    # (frames from the host take the stage-B graph without the conditional node: the four fallback
    # kernels are launched and exit early)
    assert a.launch_total() >= b.launch_total() > 0
    # and frame by frame on single-stream handles
    for i in range(S):
        kw1 = dict(kw, n_streams=1)
        c = _hot((H, W), **kw1)
        for t in range(T):
            c.process(clip[t, i])
            assert ca[t, i, 0] == len(c.raw_blobs()[0]), (i, t)
        assert np.array_equal(c.mask(), a.mask(i))
        c.close()
    a.close(); b.close()


def _assert_blobs_equal(got, want, ctx=None):
    assert len(got) == len(want), (ctx, len(got), len(want))
    for g, w in zip(got, want):
        assert np.array_equal(g["contour"], w["contour"]), ctx
        assert np.array_equal(g["center"], w["center"]) and np.array_equal(g["bottom"], w["bottom"]), ctx


def test_frames_beyond_the_blob_capacities_are_rerun_not_refused():
    """the reference returns a blob list for every frame, however busy.  A frame with more
    components / kept blobs / contour points than the handle's per-slot capacities is flagged
    (frame_info.overflow) and blob creation runs again with larger buffers when its blobs are fetched:
    same list as the oracle, temporal state untouched, later frames unaffected."""
    import cv2
    rs = np.random.RandomState(3)
    H, W = 200, 320
    busy = np.zeros((H, W), np.uint8)
    for _ in range(160):  # far more than 8 blobs of area >= 20, with holes and nesting
        x, y = rs.randint(0, W - 14), rs.randint(0, H - 14)
        cv2.rectangle(busy, (x, y), (x + rs.randint(6, 13), y + rs.randint(6, 13)), 255, int(rs.choice([-1, 1])))
    speckle = np.zeros((H, W), np.uint8)
    speckle[::2, ::2] = 255  # 16000 isolated pixels: components overflow, nothing kept
    calm = np.zeros((H, W), np.uint8)
    cv2.rectangle(calm, (30, 40), (90, 120), 255, -1)
    for caps in (dict(max_blobs=64, max_kept=8, max_points=256), dict(max_blobs=1 << 16, max_kept=8, max_points=0),
                 dict(max_blobs=1 << 16, max_kept=4096, max_points=64)):
        hp = _hot((H, W), blur=None, detector="FD", find_blobs=True, fd_threshold=0, area_min=20.0, **caps)
        prev = np.zeros((H, W), np.uint8)
        hp.process(prev)
        for k, M in enumerate((busy, calm, speckle | calm, busy, calm)):
            frame = np.where(M > 0, 255 - prev, prev).astype(np.uint8)  # FD: |frame - prev| > 0 == M
            hp.process(frame)
            prev = frame
            want_mask = np.where(M > 0, 255, 0).astype(np.uint8)
            assert np.array_equal(hp.mask(), want_mask)
            _assert_blobs_equal(hp.blobs(), oblobs.blobs_from_mask(want_mask, area_min=20.0), (caps, k))
            if M is busy:
                assert hp.frame_info()["overflow"] != 0
        hp.close()


def test_clip_call_with_small_output_buffers_still_returns_everything():
    """process_sequence with output capacities far too small must not lose the clip: the
    library runs it to the end, keeps the lists, and the host fetches them (sequence_results)."""
    H, W, T = 96, 160, 30
    r3 = np.ones((3, 3), np.uint8)
    kw = dict(blur=("gaussian", (3, 3), 4.0), detector="MOG2", morph=[("erode", r3), ("dilate", r3)], area_min=20.0)
    clip = np.stack(SyntheticClip(H, W, 1, seed=77, n_obj=5).frames(T))
    a, b = _hot((H, W), **kw), _hot((H, W), **kw)
    ca, ra, pa = a.process_sequence(clip)
    assert ca[..., 0].sum() > 8
    cb, rb, pb = b.process_sequence(clip, blob_capacity=3, point_capacity=50)
    assert np.array_equal(ca, cb) and np.array_equal(pa, pb) and ra.tobytes() == rb.tobytes()
    # and the handle goes on as if nothing had happened
    more = np.stack(SyntheticClip(H, W, 1, seed=78, n_obj=5).frames(12))
    ca, ra, pa = a.process_sequence(more)
    cb, rb, pb = b.process_sequence(more)
    assert np.array_equal(ca, cb) and np.array_equal(pa, pb) and ra.tobytes() == rb.tobytes()
    assert np.array_equal(a.mask(), b.mask())
    a.close(); b.close()


# ----------------------------------------------------------------------------- streams
def test_streams_are_independent_and_match_single_stream_runs():
    H, W, S = 48, 80, 3
    r3 = np.ones((3, 3), np.uint8)
    morph = [("erode", r3), ("dilate", r3)]
    multi = _hot((H, W), detector="MOG2", morph=morph, n_streams=S, area_min=20)
    singles = [_hot((H, W), detector="MOG2", morph=morph, n_streams=1, area_min=20) for _ in range(S)]
    clips = [SyntheticClip(H, W, 1, seed=100 + i, n_obj=3) for i in range(S)]
    for t in range(25):
        frames = np.stack([c.next() for c in clips])
        multi.process(frames)
        for i in range(S):
            singles[i].process(frames[i])
            assert np.array_equal(multi.mask(i), singles[i].mask())
            a, b = multi.blobs(i), singles[i].blobs()
            assert len(a) == len(b)
            for x, y in zip(a, b):
                assert np.array_equal(x["contour"], y["contour"])


def test_device_resident_frames_match_host_frames():
    import torch
    H, W = 60, 100
    a = _hot((H, W), detector="MOG2", find_blobs=True, area_min=10)
    b = _hot((H, W), detector="MOG2", find_blobs=True, area_min=10)
    clip = SyntheticClip(H, W, 1, seed=8)
    for t in range(10):
        f = clip.next()
        a.process(f)
        d = torch.from_numpy(f).cuda()
        b.process_device(d.data_ptr(), W, W * H, torch.cuda.current_stream().cuda_stream)
        assert np.array_equal(a.mask(), b.mask())
        assert a.frame_info() == b.frame_info()


def test_shape_change_raises_value_error():
    hp = _hot((20, 30), detector="FD", find_blobs=False)
    hp.process(np.zeros((20, 30), np.uint8))
    with pytest.raises(ValueError):
        hp.process(np.zeros((21, 30), np.uint8))
