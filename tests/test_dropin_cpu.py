"""CPU: the GOFPID drop-in's host logic -- parameter validation (same exception types as the
reference, pipeline.py:214-306,355-373), helpers' known answers (reference tests/test_helpers.py)
and the OpenCV-parameter -> kernel-configuration mapping.  No GPU needed: the CUDA pipeline is
only created on the first frame."""
import numpy as np
import pytest

from pygofpid_b200 import GOFPID
from pygofpid_b200 import helpers as h

try:
    import cv2 as cv
except Exception:
    cv = None

PF = {'perimeter': None, 'anchor': 'center', 'perspective': None, 'perspective_coeff': 0.5,
      'presence_min': 3, 'distance_min': 0.25}


def make(**kw):
    kw.setdefault('post_filter', dict(PF))
    return GOFPID(**kw)


def test_initialize_returns_self_and_resets_state():
    g = make()
    assert g.initialize() is g
    assert g.tracked_blobs_ == [] and g.input_shape_ is None and g.blobs_ == []
    assert g.post_filter['perimeter'].shape == (4, 2) and g.post_filter['perspective'].shape == (4, 2)


def test_blur_errors_and_defaults():
    with pytest.raises(KeyError):
        make(blur={'ksize': (3, 3)}).initialize()
    g = make(blur={'fun': 'blur'}).initialize()
    assert g.blur['ksize'] == (3, 3) and g.blur['borderType'] == 4
    assert g._plan['blur'] == ('box', (3, 3))
    assert make(blur=None).initialize()._plan['blur'] is None


@pytest.mark.skipif(cv is None, reason="needs cv2 objects")
def test_opencv_parameter_objects_map_to_kernel_config():
    g = make(blur={'fun': cv.GaussianBlur, 'ksize': (5, 5), 'borderType': cv.BORDER_REPLICATE},
             mat_morph=[{'fun': cv.dilate, 'kernel': cv.getStructuringElement(cv.MORPH_ELLIPSE, (15, 15))},
                        {'fun': cv.erode}]).initialize()
    assert g._plan['blur'] == ('gaussian', (5, 5), 1.0)  # borderType lands in sigmaX (reference quirk)
    assert [m[0] for m in g._plan['morph']] == ['dilate', 'erode']
    assert g._plan['morph'][0][1].shape == (15, 15) and g._plan['morph'][1][1].shape == (5, 5)
    with pytest.raises(NotImplementedError):
        make(blur={'fun': cv.medianBlur}).initialize()
    with pytest.raises(NotImplementedError):
        make(mat_morph=[{'fun': cv.morphologyEx}]).initialize()


def test_detector_and_morph_errors():
    with pytest.raises(ValueError):
        make(frg_detect='blabla').initialize()
    with pytest.raises(KeyError):
        make(mat_morph=[{'kernel': np.ones((3, 3), np.uint8)}]).initialize()
    for det in ('MOG2', 'KNN', 'FD', 'ViBe'):
        make(frg_detect=det).initialize()


def test_tracking_and_alarm_errors():
    with pytest.raises(KeyError):
        make(tracking={'absence_max': 10}).initialize()
    with pytest.raises(KeyError):
        make(tracking={'factor': 1.5}).initialize()
    with pytest.raises(KeyError):
        make(alarm_def={'duration_min': 10}).initialize()
    with pytest.raises(KeyError):
        make(alarm_def={'keep_alarm': True}).initialize()


@pytest.mark.parametrize("drop", ['perimeter', 'anchor', 'perspective', 'perspective_coeff', 'presence_min',
                                  'distance_min'])
def test_post_filter_missing_keys(drop):
    pf = dict(PF)
    del pf[drop]
    with pytest.raises(KeyError):
        GOFPID(post_filter=pf).initialize()


def test_post_filter_bad_values():
    bad = [dict(PF, anchor='top'), dict(PF, perimeter=[[0, 0], [1, 1], [1, 0]]),
           dict(PF, perimeter=np.zeros((4, 2, 1))), dict(PF, perimeter=np.zeros((2, 2))),
           dict(PF, perimeter=np.zeros((4, 3))), dict(PF, perspective=[[0, 0]] * 4),
           dict(PF, perspective=np.zeros((3, 2)))]
    for pf in bad:
        with pytest.raises(ValueError):
            GOFPID(post_filter=pf).initialize()


# ---- helpers: known answers of the reference's tests/test_helpers.py ----------------------------
def test_find_point_and_line():
    pts = np.array([[10, 10], [10, 50], [50, 50], [50, 10]])
    th = np.array([3, 3])
    assert h.find_point((11, 9), pts, th) == 0 and h.find_point((49, 52), pts, th) == 2
    assert h.find_point((30, 30), pts, th) == -1
    assert h.find_line((10, 30), pts, th) == 0 and h.find_line((30, 50), pts, th) == 1
    assert h.find_line((30, 30), pts, th) == -1 and h.find_line((1, 1), pts[:1], th) == -1


def test_bottom_center_of_square_contour():
    c = np.array([[[0, 0]], [[0, 10]], [[10, 10]], [[10, 0]]], np.int32)
    assert np.array_equal(h.get_bottom(c), [5, 10]) and h.get_bottom(c).dtype == np.int16
    assert np.array_equal(h.get_center(c), [5, 5])


def test_coords_roundtrip_and_regression():
    shape = (240, 320, 3)
    c = np.array([[32, 24], [160, 120]])
    n = h.normalize_coords(c, shape)
    assert np.allclose(n, [[0.1, 0.1], [0.5, 0.5]])
    assert np.array_equal(h.unnormalize_coords(n, shape), c)
    assert c.dtype.kind == 'i'  # inputs are not modified
    slr = h.SimpleLinearRegression().fit([0, 10], [1, 21])
    assert slr.coeff == 2 and slr.intercept == 1
    assert np.allclose(slr.predict([0, 5]), [1, 11])
    assert np.allclose(slr.predict_clip([-5, 5], 0), [0, 11])
    with pytest.raises(ValueError):
        slr.fit([0, 1, 2], [0, 1, 2])
    with pytest.raises(ValueError):
        slr.fit([0, 1], [0, 1, 2])


def test_cdist_and_dist():
    a, b = np.array([[0, 0], [3, 4]]), np.array([[0, 0], [6, 8], [3, 4]])
    d = h.cdist_euclidean(a, b)
    assert d.shape == (2, 3) and np.allclose(d, [[0, 10, 5], [5, 5, 0]])
    assert h.dist_euclidean(3, 7) == 4 and np.isclose(h.dist_euclidean(np.array([0, 0]), np.array([3, 4])), 5)
    assert h.change_extension("clip.avi") == "clip.png"


@pytest.mark.skipif(cv is None, reason="needs cv2")
def test_point_in_polygon_matches_cv2():
    rs = np.random.RandomState(3)
    poly = np.array([[2, 3], [40, 5], [35, 30], [20, 18], [5, 28]], np.int32)
    for _ in range(2000):
        p = (int(rs.randint(0, 45)), int(rs.randint(0, 35)))
        assert h.point_in_polygon(poly, p) == cv.pointPolygonTest(poly, p, False)


def test_tracking_distance_matrix_equals_the_scalar_code():
    """The vectorised (blob, track) distance matrix of the tracking tail (reference
    pipeline.py:590-662) reproduces the scalar code value for value, int16 wrap-around included."""
    import warnings
    from pygofpid_b200.pipeline import GOFPID
    g = GOFPID.__new__(GOFPID)
    g.input_shape_ = (1080, 1920)
    g.tracking = {'factor': 1.5}
    g._perspective = {'height': np.clip(np.linspace(-20, 300, 1080), 0, None)}
    rs = np.random.RandomState(0)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for trial in range(10):
            nt, nb = rs.randint(1, 40), rs.randint(1, 40)
            tracks = [{'bottom': np.array([rs.randint(0, 1920), rs.randint(-5, 1100)], np.int16),
                       'center': np.array([rs.randint(0, 1920), rs.randint(0, 1080)], np.int16)}
                      for _ in range(nt)]
            blobs = [{'center': np.array([rs.randint(0, 1920), rs.randint(0, 1080)], np.int16)}
                     for _ in range(nb)]
            want = np.array([[g._tracking_distance(t, b) for t in tracks] for b in blobs], dtype=np.float64)
            got = g._tracking_distance_matrix(tracks, blobs)
            assert np.array_equal(got, want, equal_nan=True), trial


@pytest.mark.parametrize("name", ["fd_gray_open3", "mog2_gray_default", "mog2_bgr_default", "knn_gray_default",
                                  "mog2_gray_box_openclose", "mog2_gray_convert_bgr"])
def test_host_tail_replays_the_reference_tracks_and_alarms(name):
    """Steps 5-7 (tracking, post-filter, alarm: reference pipeline.py:590-799) on the host, fed with
    the blob lists the REAL reference produced (tests/golden/*.npz): the number of tracked blobs and
    the alarm of every frame are the reference's.  No GPU involved."""
    from golden_cases import golden_blobs, load
    g = load(name)
    H, W, C = (int(v) for v in g["shape"])
    pf = {'perimeter': None, 'anchor': 'bottom', 'perspective': None, 'perspective_coeff': 0.5,
          'presence_min': 3, 'distance_min': 0.25}
    gofpid = GOFPID(post_filter=pf).initialize()
    gofpid.input_shape_ = (H, W) if C == 1 else (H, W, C)
    gofpid._calib_first_frame()
    for t in range(int(g["n"])):
        blobs = golden_blobs(g, t)
        for b in blobs:
            b["center"] = np.asarray(b["center"], np.int16)
            b["bottom"] = np.asarray(b["bottom"], np.int16)
            b["anchor"] = b["bottom"]
        gofpid.blobs_ = blobs
        gofpid._track_blob()
        gofpid._post_filter()
        y = gofpid._maintain_alarm(gofpid._detect_blob())
        assert bool(y) == bool(g["alarm"][t]), (name, t)
        assert len(gofpid.tracked_blobs_) == int(g["ntracked"][t]), (name, t)


@pytest.mark.parametrize("trial", range(12))
def test_host_tail_matches_the_reference_on_busy_blob_streams(trial):
    """Up to 30 moving blobs with missed detections, three frame sizes, two post-filter settings
    (perimeter polygon + perspective boxes / defaults): alarm, number of tracks, summed presence and
    number of active filters per frame are what the REAL reference's tail answered
    (tests/golden/tail_streams.npz, generated by tests/golden/make_tail_golden.py)."""
    import os
    import warnings
    from golden_cases import GOLDEN_DIR, tail_post_filter, tail_stream
    want = np.load(os.path.join(GOLDEN_DIR, "tail_streams.npz"))
    H, W, kind, frames = tail_stream(trial)
    g = GOFPID(post_filter=tail_post_filter(kind)).initialize()
    g.input_shape_ = (H, W)
    g._calib_first_frame()
    for t, blobs in enumerate(frames):
        g.blobs_ = blobs
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            g._track_blob()
            g._post_filter()
            y = g._maintain_alarm(g._detect_blob())
        assert bool(y) == bool(want["alarm_%d" % trial][t]), t
        assert len(g.tracked_blobs_) == int(want["ntracked_%d" % trial][t]), t
        assert sum(tr['presence'] for tr in g.tracked_blobs_) == int(want["presence_%d" % trial][t]), t
        assert sum(len(tr['filtered']) for tr in g.tracked_blobs_) == int(want["nfiltered_%d" % trial][t]), t


@pytest.mark.parametrize("trial", range(12))
def test_native_tail_matches_the_reference_on_busy_blob_streams(trial):
    """The C++ tracker of libgofpid_b200 (gofpid_tracker: pure host code, no GPU needed) on the same streams:
    alarm, number of tracks, summed presence and number of active filters of every frame are what the REAL
    reference answered, and its tracks equal those of the Python tail field by field."""
    import os
    import warnings
    from golden_cases import GOLDEN_DIR, tail_post_filter, tail_stream
    want = np.load(os.path.join(GOLDEN_DIR, "tail_streams.npz"))
    H, W, kind, frames = tail_stream(trial)
    g = GOFPID(post_filter=tail_post_filter(kind)).initialize()
    g.input_shape_ = (H, W)
    g._calib_first_frame()
    tail = g._native_tail()
    py = GOFPID(post_filter=tail_post_filter(kind), tail='python').initialize()
    py.input_shape_ = (H, W)
    py._calib_first_frame()
    for t, blobs in enumerate(frames):
        y = tail.step_dicts(blobs)
        tracks = g.tracked_blobs_
        assert bool(y) == bool(want["alarm_%d" % trial][t]), t
        assert len(tracks) == len(tail) == int(want["ntracked_%d" % trial][t]), t
        assert sum(tr['presence'] for tr in tracks) == int(want["presence_%d" % trial][t]), t
        assert sum(len(tr['filtered']) for tr in tracks) == int(want["nfiltered_%d" % trial][t]), t
        py.blobs_ = blobs
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            py._track_blob()
            py._post_filter()
            yp = py._maintain_alarm(py._detect_blob())
        assert yp == y
        if t % 10 == 0 or t == len(frames) - 1:
            for a, b in zip(tracks, py.tracked_blobs_):
                for key in ('anchor', 'bottom', 'center', 'center_first', 'contour'):
                    assert np.array_equal(a[key], b[key]) and a[key].dtype == np.asarray(b[key]).dtype, (t, key)
                for key in ('presence', 'absence', 'filtered', 'already_in_alarm'):
                    assert a[key] == b[key], (t, key)
    state = tail.alarm_state()
    assert state['already_in_alarm'] == py._alarm_system['already_in_alarm'] and state['duration'] == py._alarm_system['duration']


@pytest.mark.parametrize("name", ["fd_gray_open3", "mog2_gray_default", "mog2_bgr_default", "knn_gray_default"])
def test_native_tail_replays_the_reference_tracks_and_alarms(name):
    """... and on the blob lists the REAL reference produced on the pipeline goldens."""
    from golden_cases import golden_blobs, load
    g = load(name)
    H, W, C = (int(v) for v in g["shape"])
    pf = {'perimeter': None, 'anchor': 'bottom', 'perspective': None, 'perspective_coeff': 0.5,
          'presence_min': 3, 'distance_min': 0.25}  # (the constructor's default dict is shared and mutated, as in the reference)
    gofpid = GOFPID(post_filter=pf).initialize()
    gofpid.input_shape_ = (H, W) if C == 1 else (H, W, C)
    gofpid._calib_first_frame()
    tail = gofpid._native_tail()
    for t in range(int(g["n"])):
        blobs = golden_blobs(g, t)
        for b in blobs:
            b["center"] = np.asarray(b["center"], np.int16)
            b["bottom"] = np.asarray(b["bottom"], np.int16)
            b["anchor"] = b["bottom"]
        assert tail.step_dicts(blobs) == bool(g["alarm"][t]), (name, t)
        assert len(tail) == int(g["ntracked"][t]), (name, t)


def test_native_tail_degenerate_geometry_follows_the_reference_types():
    """int16 wrap-around of squared centre differences, zero radii, empty frames: the C++ tracker and the
    Python tail (which evaluates the reference's expressions in NumPy's own types) agree."""
    import warnings
    rs = np.random.RandomState(4)
    H, W = 1080, 1920
    for trial in range(4):
        pf = {'perimeter': None, 'anchor': 'center' if trial % 2 else 'bottom', 'perspective': None,
              'perspective_coeff': 0.5, 'presence_min': 2, 'distance_min': 0.0 if trial == 3 else 0.2}
        trk = {'factor': (1.5, 40.0, 0.01, 3.0)[trial], 'absence_max': 3}
        a = GOFPID(post_filter=dict(pf), tracking=dict(trk)).initialize()
        b = GOFPID(post_filter=dict(pf), tracking=dict(trk), tail='python').initialize()
        for g in (a, b):
            g.input_shape_ = (H, W)
            g._calib_first_frame()
        tail = a._native_tail()
        for t in range(60):
            blobs = []
            for _ in range(rs.randint(0, 6)):
                x0, y0, w, h = rs.randint(0, W - 300), rs.randint(0, H - 300), rs.randint(2, 300), rs.randint(2, 300)
                c = np.array([[x0, y0], [x0 + w, y0], [x0 + w, y0 + h], [x0, y0 + h]], np.int32).reshape(-1, 1, 2)
                bottom = np.array([x0 + w // 2, y0 + h], np.int16)
                center = np.array([x0 + w // 2, y0 + h // 2], np.int16)
                blobs.append(dict(contour=c, anchor=center if trial % 2 else bottom, bottom=bottom, center=center))
            ya = tail.step_dicts(blobs)
            b.blobs_ = blobs
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                b._track_blob()
                b._post_filter()
                yb = b._maintain_alarm(b._detect_blob())
            assert ya == yb, (trial, t)
            ta, tb = a.tracked_blobs_, b.tracked_blobs_
            assert len(ta) == len(tb), (trial, t)
            for u, v in zip(ta, tb):
                assert u['filtered'] == v['filtered'] and u['presence'] == v['presence'] and u['absence'] == v['absence']
                assert np.array_equal(u['center'], v['center']) and np.array_equal(u['center_first'], v['center_first'])


@pytest.mark.skipif(cv is None, reason="display() draws with OpenCV")
@pytest.mark.parametrize("trial", [0, 3])
def test_display_draws_what_the_reference_draws(trial):
    """GOFPID.display (reference pipeline.py:801-853) draws perimeter, contours, anchors, tracking
    ellipses and perspective boxes into the frame in place: same pixels as the REAL reference on the
    same tracked blobs (canvases in tests/golden/tail_streams.npz)."""
    import os
    import warnings
    from golden_cases import DISPLAY_FRAMES, GOLDEN_DIR, tail_post_filter, tail_stream
    want = np.load(os.path.join(GOLDEN_DIR, "tail_streams.npz"))
    H, W, kind, frames = tail_stream(trial)
    g = GOFPID(post_filter=tail_post_filter(kind)).initialize()
    g.input_shape_ = (H, W)
    g._calib_first_frame()
    seen = 0
    for t, blobs in enumerate(frames):
        g.blobs_ = blobs
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            g._track_blob()
            g._post_filter()
            g._maintain_alarm(g._detect_blob())
        if t in DISPLAY_FRAMES:
            canvas = np.zeros((H, W, 3), np.uint8)
            g.display(canvas, {'tracking': True, 'perspective': True})
            ref = want["display_%d_%d" % (trial, t)]
            assert ref.any() and np.array_equal(canvas, ref), (trial, t, int((canvas != ref).sum()))
            seen += 1
    assert seen == len(DISPLAY_FRAMES)


def test_detect_clip_validates_its_input_before_touching_the_gpu():
    g = make(frg_detect='MOG2').initialize()
    with pytest.raises(ValueError):
        g.detect_clip(np.zeros((0, 8, 8), np.uint8))   # no frames
    with pytest.raises(ValueError):
        g.detect_clip(np.zeros((8, 8), np.uint8))      # a frame, not a clip
    g.input_shape_ = (8, 8)
    with pytest.raises(ValueError, match="Input shape has changed"):
        g.detect_clip(np.zeros((2, 9, 8), np.uint8))
