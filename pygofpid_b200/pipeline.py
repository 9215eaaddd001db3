"""`GOFPID`: good-old-fashioned perimeter intrusion detection, drop-in for pygofpid.pipeline.GOFPID.

Same constructor parameters, `initialize()` / `detect(X)` / `display(X, display)` methods,
`foreground_mask_` / `blobs_` / `tracked_blobs_` / `input_shape_` attributes and the same
exceptions for the same mistakes as the reference (pygofpid/pipeline.py:168-283, 434-493).

What changed is where the per-frame pixel work runs: steps 1-4 of `detect`

    1. spatial blur            (reference pipeline.py:462-467, cv.GaussianBlur / cv.blur)
    2. foreground detection    (pipeline.py:470, MOG2 / KNN / frame differencing)
    3. mask morphology         (pipeline.py:473-476, cv.erode / cv.dilate chain)
    4. blob creation           (pipeline.py:555-588, findContours + contourArea + moments)

are one CUDA pipeline (`HotPath`, libgofpid_b200.so) whose background model lives in HBM.
Steps 5-7 (tracking, post-filtering, alarm) consume the blob list on the host exactly as the
reference does; they touch a handful of blobs per frame and no pixels.

The OpenCV callables of the reference's parameter dicts (`cv.GaussianBlur`, `cv.blur`,
`cv.erode`, `cv.dilate`) are accepted as they are and mapped to kernel configurations by
identity / name; their string names work too when OpenCV is not installed.  There is no CPU
fallback for steps 1-4: anything the kernels do not cover raises NotImplementedError.
"""
import math

import numpy as np

from .helpers import SimpleLinearRegression, point_in_polygon, unnormalize_coords
from .segmentation import ViBe

try:  # OpenCV is only needed for the parameter objects users pass in, and for display()
    import cv2 as cv
except Exception:  # pragma: no cover
    cv = None

_BORDER_DEFAULT = 4  # cv.BORDER_DEFAULT == cv.BORDER_REFLECT_101
_COLOR_BGR2GRAY, _COLOR_RGB2GRAY = 6, 7


def _cv_name(fun):
    """'GaussianBlur' | 'blur' | 'erode' | 'dilate' for a cv2 callable or its name."""
    if isinstance(fun, str):
        return fun
    if cv is not None:
        for name in ("GaussianBlur", "blur", "erode", "dilate"):
            if fun is getattr(cv, name, None):
                return name
    return getattr(fun, "__name__", None)


def _rect_kernel(n):
    return np.ones((n, n), dtype=np.uint8)


def _default_blur():
    return {'fun': cv.GaussianBlur if cv is not None else 'GaussianBlur', 'ksize': (3, 3),
            'borderType': _BORDER_DEFAULT}


def _default_morph():
    return [{'fun': cv.erode if cv is not None else 'erode', 'kernel': _rect_kernel(5)},
            {'fun': cv.dilate if cv is not None else 'dilate', 'kernel': _rect_kernel(5)}]


def _to_gray(X, code):
    """cv.cvtColor(X, COLOR_BGR2GRAY | COLOR_RGB2GRAY) for uint8 (15-bit fixed point, exact)."""
    w = (3735, 19235, 9798) if code == _COLOR_BGR2GRAY else (9798, 19235, 3735)
    Xi = X.astype(np.uint32)
    return ((Xi[..., 0] * w[0] + Xi[..., 1] * w[1] + Xi[..., 2] * w[2] + 16384) >> 15).astype(np.uint8)


class GOFPID():
    """GOFPID: good old fashioned perimeter intrusion detection system (B200 hot path).

    Parameters
    ----------
    convert : OpenCV color conversion code | None, default=None
        Convert the input frame from colour to gray first.
    blur : dict | None, default={'fun': cv.GaussianBlur, 'ksize': (3, 3), \
            'borderType': cv.BORDER_DEFAULT}
        Spatial blur: cv.GaussianBlur or cv.blur and its parameters; None disables it.
        As in the reference the third value is passed POSITIONALLY, so for cv.GaussianBlur it
        is sigmaX (default 4.0) and for cv.blur it has no effect; borders are REFLECT_101.
    frg_detect : {'MOG2', 'KNN', 'ViBe', 'FD'}, default='MOG2'
        Foreground detection method.
    mat_morph : list of dict | None
        cv.erode / cv.dilate steps with their 'kernel'; default: 5x5 open.
    tracking, post_filter, alarm_def, verbose :
        as in the reference (pipeline.py:69-128).
    device : int, default=0
        CUDA device of this stream's pipeline.
    knn_seed : int, default=0
        State of the cv::RNG replica KNN draws from (0 = OpenCV's default state); the
        reference consumes the calling thread's cv.theRNG() instead.
    max_blobs, max_kept, max_points : int
        Per-frame capacities of the GPU path (connected components, blobs that pass the area filter,
        contour points; 0 points = sized from the frame).  A frame that exceeds them is not refused:
        its blobs are created again with larger buffers (slower for that frame).
    depth : int, default=0
        Frame slots of the GPU pipeline (0 = default: deep, for `detect_clip`); `detect` alone never
        has more than one frame in flight and works with depth=2 (saves device memory).
    tail : {'native', 'python'}, default='native'
        Who runs steps 5-7 (tracking, post-filter, alarm): the C++ tracker of libgofpid_b200
        (`gofpid_tracker`, same results, a few microseconds per frame; `blobs_` and `tracked_blobs_` are
        then built from its state when they are read) or the Python methods of this class.

    Attributes
    ----------
    foreground_mask_ : ndarray (H, W) uint8, fetched from the GPU when read.
    blobs_ : list of dict(contour, anchor, bottom, center)
    tracked_blobs_ : list of dict, see the reference.
    """

    def __init__(
        self,
        convert=None,
        blur=_default_blur(),
        frg_detect='MOG2',
        mat_morph=_default_morph(),
        tracking={'factor': 1.5, 'absence_max': 10},
        post_filter={'perimeter': None, 'anchor': 'bottom', 'perspective': None,
                     'perspective_coeff': 0.5, 'presence_min': 3, 'distance_min': 0.25},
        alarm_def={'keep_alarm': True, 'duration_min': 10},
        verbose=False,
        device=0,
        knn_seed=0,
        max_blobs=1 << 16,
        max_kept=4096,
        max_points=0,
        depth=0,
        tail='native',
    ):
        self.convert = convert
        self.blur = blur
        self.frg_detect = frg_detect
        self.mat_morph = mat_morph
        self.tracking = tracking
        self.post_filter = post_filter
        self.alarm_def = alarm_def
        self.verbose = verbose
        self.device = device
        self.knn_seed = knn_seed
        self.max_blobs = max_blobs
        self.max_kept = max_kept
        self.max_points = max_points
        self.depth = depth
        self.tail = tail

    # ------------------------------------------------------------------ configuration
    def initialize(self):
        """Check the parameters and set the pipeline up (reference pipeline.py:214-283).

        Raises KeyError for missing dictionary keys and ValueError for bad values, at the
        same places as the reference."""
        if self.blur:
            if 'fun' not in self.blur.keys():
                raise KeyError('Parameter blur has no key "fun".')
            self.blur.setdefault('ksize', (3, 3))
            self.blur.setdefault('borderType', _BORDER_DEFAULT)

        if self.frg_detect not in ('MOG2', 'KNN', 'FD', 'ViBe'):
            raise ValueError('Unknown method for foreground detection')

        if self.mat_morph:
            for d in self.mat_morph:
                if 'fun' not in d.keys():
                    raise KeyError('Parameter mat_morph has no key "fun".')
                d.setdefault('kernel', _rect_kernel(5))

        for key in ('factor', 'absence_max'):
            if key not in self.tracking.keys():
                raise KeyError('Parameter tracking has no key "%s".' % key)

        if 'anchor' not in self.post_filter.keys():
            raise KeyError('Parameter post_filter has no key "anchor".')
        if self.post_filter['anchor'] not in ['bottom', 'center']:
            raise ValueError('Parameter anchor must be "bottom" or "center".')
        for key in ('perspective_coeff', 'presence_min', 'distance_min'):
            if key not in self.post_filter.keys():
                raise KeyError('Parameter post_filter has no key "%s".' % key)
        self._check_perimeter()
        self._check_perspective()

        for key in ('keep_alarm', 'duration_min'):
            if key not in self.alarm_def.keys():
                raise KeyError('Parameter alarm_def has no key "%s".' % key)

        if self.tail not in ('native', 'python'):
            raise ValueError('Parameter tail must be "native" or "python".')
        self._plan = self._hot_path_plan()
        self._engine = None           # created on the first frame (the shape is needed)
        self._host_estimator = ViBe() if self.frg_detect == 'ViBe' else None
        self._mask_cache = None
        self._tail = None             # native tracker, created with the first frame `detect` sees
        self._raw = None              # (records, points) of the last frame when the native tail is in use
        self.blobs_ = []
        self.tracked_blobs_ = []
        self.input_shape_ = None
        self._alarm_system = {'already_in_alarm': False, 'duration': 0}
        return self

    # `blobs_` / `tracked_blobs_` are plain lists with the Python tail.  With the native tail they are
    # built when read: from the last frame's blob records, and from the tracker's state.
    @property
    def blobs_(self):
        if self._raw is not None:
            from .hotpath import blobs_from_records
            self._blobs_py = blobs_from_records(self._raw[0], self._raw[1], self.post_filter['anchor'])
            self._raw = None
        return self._blobs_py

    @blobs_.setter
    def blobs_(self, value):
        self._raw = None
        self._blobs_py = value

    @property
    def tracked_blobs_(self):
        if getattr(self, '_tail', None) is not None:
            return self._tail.tracks()
        return self._tracked_py

    @tracked_blobs_.setter
    def tracked_blobs_(self, value):
        self._tracked_py = value

    def _native_tail(self):
        """The C++ tracker of this stream (created after the first-frame calibration)."""
        if self._tail is None:
            from .tail import NativeTail
            self._tail = NativeTail(self.input_shape_[0], self._perspective['height'], self._perspective['area'],
                                    self.post_filter['perimeter'], self.tracking, self.post_filter, self.alarm_def)
        return self._tail

    def _hot_path_plan(self):
        """Translate the reference's OpenCV-flavoured parameters into kernel configuration."""
        blur = None
        if self.blur:
            name = _cv_name(self.blur.get('fun'))
            ksize = tuple(int(k) for k in self.blur.get('ksize'))
            if name == 'GaussianBlur':
                # fun(X, ksize, borderType): the third positional argument of GaussianBlur is sigmaX
                blur = ('gaussian', ksize, float(self.blur.get('borderType')))
            elif name == 'blur':
                blur = ('box', ksize)  # third positional argument of cv.blur is `dst`: ignored
            else:
                raise NotImplementedError(
                    "blur['fun'] must be cv.GaussianBlur or cv.blur (the CUDA hot path has no "
                    "CPU fallback for other callables): %r" % (self.blur.get('fun'),))
        morph = []
        for d in (self.mat_morph or []):
            name = _cv_name(d.get('fun'))
            if name not in ('erode', 'dilate'):
                raise NotImplementedError(
                    "mat_morph 'fun' must be cv.erode or cv.dilate: %r" % (d.get('fun'),))
            morph.append((name, np.asarray(d.get('kernel'))))
        return {'blur': blur, 'morph': morph}

    def _check_perimeter(self):
        pf = self.post_filter
        if 'perimeter' not in pf.keys():
            raise KeyError('Parameter post_filter has no key "perimeter".')
        per = pf['perimeter']
        if per is None:
            pf['perimeter'] = np.array([[0, 0], [0, 1], [1, 1], [1, 0]])
        elif isinstance(per, np.ndarray):
            if per.ndim != 2:
                raise ValueError('Parameter perimeter has not the good shape.')
            if per.shape[0] < 3:
                raise ValueError('Parameter perimeter has not enough points.')
            if per.shape[1] != 2:
                raise ValueError('Parameter perimeter has not the good shape.')
        else:
            raise ValueError('Unknown type for parameter "perimeter".')
        if pf.get('display_config'):
            pf['perimeter'] = self._config_perimeter()

    def _check_perspective(self):
        pf = self.post_filter
        if 'perspective' not in pf.keys():
            raise KeyError('Parameter post_filter has no key "perspective".')
        per = pf['perspective']
        if per is None:
            pf['perspective'] = np.array([[0.1, 0.4], [0.2, 0.8], [0.8, 0.2], [0.85, 0.3]])
        elif isinstance(per, np.ndarray):
            if per.shape != (4, 2):
                raise ValueError('Parameter perspective has not the good shape.')
        else:
            raise ValueError('Unknown type for parameter "perspective".')
        if pf.get('display_config'):
            pf['perspective'] = self._config_perspective()

    def _config_perimeter(self, window_name="Configure perimeter"):
        raise NotImplementedError("interactive configuration windows (reference pipeline.py:308-407) are "
                                  "outside the hot path: pass 'perimeter' as an array")

    def _config_perspective(self, window_name="Configure perspective"):
        raise NotImplementedError("interactive configuration windows (reference pipeline.py:308-407) are "
                                  "outside the hot path: pass 'perspective' as an array")

    # ------------------------------------------------------------------ per frame
    @property
    def foreground_mask_(self):
        """Foreground mask of the last frame, (H, W) uint8 (reference pipeline.py:476);
        copied from the GPU on first access after each `detect`."""
        if self._mask_cache is None:
            if getattr(self, '_engine', None) is None:  # the reference sets the attribute in detect()
                raise AttributeError("'GOFPID' object has no attribute 'foreground_mask_'")
            self._mask_cache = self._engine.mask()
        return self._mask_cache

    def detect(self, X):
        """Detect if there is an intrusion in the current frame.

        Parameters
        ----------
        X : ndarray of uint8, shape (H, W) or (H, W, C)

        Returns
        -------
        y : bool, intrusion detected.
        """
        if not self.input_shape_:
            self.input_shape_ = X.shape
            self._calib_first_frame()
        elif X.shape != self.input_shape_:
            raise ValueError('Input shape has changed.')

        # BGR / RGB -> gray is fused into the frame load of the GPU path; any other conversion code
        # (and the host-estimator path) converts on the host first
        fused = None
        if self.convert and self._host_estimator is None and X.ndim == 3 and X.shape[2] == 3 \
                and X.dtype == np.uint8:
            fused = {_COLOR_BGR2GRAY: "bgr2gray", _COLOR_RGB2GRAY: "rgb2gray"}.get(self.convert)
        if self.convert and fused is None:
            X = self._convert(X)

        # steps 1-4 on the GPU
        if self._host_estimator is None:
            self._hot_path(X, fused)
        else:
            self._hot_path_host_estimator(X)
        return self._run_tail()

    def _run_tail(self):
        """Steps 5-7 on the blobs of the frame that just went through the GPU."""
        if self.tail == 'native':
            recs, pts = self._raw
            y = self._native_tail().step_records(recs, len(recs), pts)
            if self.verbose and y and not self._alarm_system['already_in_alarm']:
                print("Intrusion detected")
            self._alarm_system = self._tail.alarm_state() if self.verbose else self._alarm_system
            return y
        self._track_blob()       # step 5
        self._post_filter()      # step 6
        y = self._detect_blob()  # step 7
        return self._maintain_alarm(y)

    def detect_clip(self, X):
        """A decoded clip (T, H, W[, C]) at once -- the read-detect loop of examples/video.py:45-50
        without one synchronous GPU round trip per frame: steps 1-4 of all frames go through the
        library's clip entry point (frames in flight, MOG2 carried through several frames per
        launch), steps 5-7 run per frame on the host as in `detect`.  Not part of the reference's
        API; returns what `detect` would have returned frame by frame (list of bool) and leaves
        `blobs_`, `tracked_blobs_` and `foreground_mask_` as after the last frame."""
        X = np.asarray(X)
        if X.ndim < 3 or X.shape[0] == 0:
            raise ValueError('a clip is an array of frames: (T, H, W) or (T, H, W, C)')
        fused = None
        if self.convert and self._host_estimator is None and X.ndim == 4 and X.shape[3] == 3 \
                and X.dtype == np.uint8:
            fused = {_COLOR_BGR2GRAY: "bgr2gray", _COLOR_RGB2GRAY: "rgb2gray"}.get(self.convert)
        if self._host_estimator is not None or (self.convert and fused is None) or X.dtype != np.uint8:
            return [self.detect(f) for f in X]  # host estimator / host-side conversion: frame by frame
        if not self.input_shape_:
            self.input_shape_ = X.shape[1:]
            self._calib_first_frame()
        elif X.shape[1:] != self.input_shape_:
            raise ValueError('Input shape has changed.')
        if self._engine is None:
            self._engine = self._new_engine(X.shape[1:], self.frg_detect, self._plan['blur'],
                                            self._plan['morph'], fused)
        from .hotpath import blobs_from_records
        frames = X[..., 0] if (X.ndim == 4 and X.shape[3] == 1) else X
        counts, recs, pts = self._engine.process_sequence(frames)
        self._mask_cache = None
        if self.tail == 'native':
            out = self._native_tail().run(counts, recs, pts)
            nb, npt = int(counts[-1, 0, 0]), int(counts[-1, 0, 1])
            self._raw = (recs[len(recs) - nb:], pts[len(pts) - npt:])  # blobs_ of the last frame, built when read
            if self.verbose and any(out):
                print("Intrusion detected")
            return out
        out, ob, op = [], 0, 0
        for t in range(X.shape[0]):
            nb, npt = int(counts[t, 0, 0]), int(counts[t, 0, 1])
            self.blobs_ = blobs_from_records(recs[ob:ob + nb], pts[op:op + npt], self.post_filter['anchor'])
            ob += nb
            op += npt
            self._track_blob()       # step 5
            self._post_filter()      # step 6
            y = self._detect_blob()  # step 7
            out.append(self._maintain_alarm(y))
        return out

    def _convert(self, X):
        if self.convert in (_COLOR_BGR2GRAY, _COLOR_RGB2GRAY) and X.ndim == 3 and X.shape[2] == 3 \
                and X.dtype == np.uint8:
            return _to_gray(X, self.convert)
        if cv is None:
            raise NotImplementedError("colour conversion code %r needs OpenCV" % (self.convert,))
        return cv.cvtColor(X, self.convert)  # a step before the hot path; other codes stay on cv2

    def _new_engine(self, shape, detector, blur, morph, convert=None):
        from .hotpath import HotPath
        if len(shape) == 3 and shape[2] == 1:
            shape = shape[:2]
        return HotPath(shape, blur=blur, detector=detector, morph=morph, area_min=100.0,
                       device=self.device, knn_seed=self.knn_seed, convert=convert,
                       max_blobs=self.max_blobs, max_kept=self.max_kept, max_points=self.max_points,
                       depth=self.depth)

    def _hot_path(self, X, convert=None):
        if self._engine is None:
            self._engine = self._new_engine(X.shape, self.frg_detect, self._plan['blur'],
                                            self._plan['morph'], convert)
        self._engine.process(X)
        self._mask_cache = None
        self._take_blobs()

    def _take_blobs(self):
        """blobs_ of the frame the engine just finished: dictionaries for the Python tail, the raw records
        for the native one (dictionaries are then only built if somebody reads blobs_)."""
        if self.tail == 'native':
            self._raw = self._engine.raw_blobs()
        else:
            self.blobs_ = self._engine.blobs(anchor=self.post_filter['anchor'])

    def _hot_path_host_estimator(self, X):
        """ViBe keeps its host-side model (see segmentation.ViBe); blur before it and
        morphology + blobs after it still run on the GPU."""
        from .hotpath import HotPath
        if self._engine is None:
            shape = X.shape[:2] if (X.ndim == 3 and X.shape[2] == 1) else X.shape
            self._pre = HotPath(shape, blur=self._plan['blur'], detector="FD", morph=(),
                                find_blobs=False, device=self.device) if self._plan['blur'] else None
            # masks are fed through frame differencing against an all-zero frame (|m - 0| > 0)
            self._engine = HotPath(shape[:2], blur=None, detector="FD", morph=self._plan['morph'],
                                   fd_threshold=0, area_min=100.0, device=self.device)
        if self._pre is not None:
            self._pre.process(X)
            X = self._pre.fd_model()  # the blurred frame is FD's state
        mask = self._host_estimator.apply(X)
        self._engine.reset()
        self._engine.process(np.zeros_like(mask))
        self._engine.process(mask)
        self._mask_cache = None
        self._take_blobs()

    # ------------------------------------------------------------------ calibration
    def _calib_first_frame(self):
        """Turn the normalised perimeter / perspective into pixels (reference pipeline.py:495-547)."""
        pf = self.post_filter
        pf['perimeter'] = unnormalize_coords(pf['perimeter'], self.input_shape_, dtype=np.int32)
        pf['perspective'] = unnormalize_coords(pf['perspective'], self.input_shape_, dtype=np.float64)
        self._calib_perspective()

    def _calib_perspective(self):
        """Area, height and width of a person as linear functions of the bottom ordinate,
        from the two calibration boxes."""
        p = self.post_filter['perspective']
        boxes = ((p[0], p[1]), (p[2], p[3]))
        bottoms = [max(a[1], b[1]) for a, b in boxes]
        heights = [abs(b[1] - a[1]) for a, b in boxes]
        widths = [abs(b[0] - a[0]) for a, b in boxes]
        areas = [h * w for h, w in zip(heights, widths)]
        rows = np.arange(0, self.input_shape_[0], 1)
        slr = SimpleLinearRegression()
        self._perspective = {}
        for key, values in (('area', areas), ('height', heights), ('width', widths)):
            self._perspective[key] = slr.fit(bottoms, values).predict_clip(rows, 0)

    def _predict_perspective(self, key, y):
        if np.ndim(y) == 0:  # one ordinate: the same clip and lookup without numpy's scalar overhead
            yi = int(y)
            last = self.input_shape_[0] - 1
            return self._perspective.get(key)[0 if yi < 0 else (last if yi > last else yi)]
        y = np.clip(y, a_min=0, a_max=self.input_shape_[0] - 1)
        return self._perspective.get(key)[y]

    # ------------------------------------------------------------------ step 5: tracking
    def _tracking_radii(self, bottom):
        """Major / minor radius of the elliptic search region of a tracked blob."""
        major = np.int32(self._predict_perspective('height', bottom[1]) * self.tracking['factor'])
        return major, np.int32(major / 2)

    def _tracking_distance(self, tracked, blob):
        """Normalised elliptic distance (<= 1: inside the search region) between a tracked
        blob and an instantaneous blob (reference pipeline.py:639-662)."""
        r_maj, r_min = self._tracking_radii(tracked['bottom'])
        if r_maj <= 0 or r_min <= 0:
            return 0
        dx = blob['center'][0] - tracked['center'][0]
        dy = blob['center'][1] - tracked['center'][1]
        return dx**2 / r_maj**2 + dy**2 / r_min**2

    def _tracking_distance_matrix(self, tracks, blobs):
        """`_tracking_distance` for every (blob, track) pair at once.  Array arithmetic in the very
        dtypes the reference's scalar code runs in (int16 centre differences and squares, wrapping
        like the scalars do; int32 radii; float64 quotients), so the matrix is identical to the
        double loop -- which costs ~130 ms per frame at 200 blobs."""
        tb = np.array([t['bottom'][1] for t in tracks])
        tc = np.array([t['center'] for t in tracks]).reshape(len(tracks), 2)
        bc = np.array([b['center'] for b in blobs]).reshape(len(blobs), 2)
        if tc.dtype != np.int16 or bc.dtype != np.int16:  # foreign blob dicts: stay on the scalar path
            return np.array([[self._tracking_distance(t, b) for t in tracks] for b in blobs], dtype=np.float64)
        major = (self._predict_perspective('height', tb) * self.tracking['factor']).astype(np.int32)
        minor = (major / 2).astype(np.int32)
        ok = (major > 0) & (minor > 0)
        with np.errstate(over='ignore', divide='ignore', invalid='ignore'):
            dx = bc[:, None, 0] - tc[None, :, 0]
            dy = bc[:, None, 1] - tc[None, :, 1]
            d = dx**2 / (major**2)[None, :] + dy**2 / (minor**2)[None, :]
        return np.where(ok[None, :], d, 0.0)

    @staticmethod
    def _new_track(blob):
        return {'contour': blob['contour'].copy(), 'anchor': blob['anchor'], 'bottom': blob['bottom'],
                'center': blob['center'], 'center_first': blob['center'], 'presence': 1, 'absence': 0,
                'filtered': set(['presence', 'distance']), 'already_in_alarm': False}

    def _relative_travel(self, track):
        """Distance between first and last centre, in units of sqrt(expected person area)."""
        c, c0 = track['center'], track['center_first']
        dx, dy = float(c[0]) - float(c0[0]), float(c[1]) - float(c0[1])  # float64, like the reference's norm
        area_min = self._predict_perspective('area', track['bottom'][1])
        if area_min <= 0:
            return 0
        return math.sqrt(dx * dx + dy * dy) / math.sqrt(area_min)

    def _pair(self, track, blob):
        track.update(contour=blob['contour'].copy(), anchor=blob['anchor'], bottom=blob['bottom'],
                     center=blob['center'], absence=0)
        track['presence'] += 1
        if track['presence'] >= self.post_filter['presence_min']:
            track['filtered'].discard('presence')
        if self._relative_travel(track) >= self.post_filter['distance_min']:
            track['filtered'].discard('distance')

    def _miss(self, j):
        track = self.tracked_blobs_[j]
        track['presence'] = 0
        track['filtered'].add('presence')
        track['absence'] += 1
        if track['absence'] >= self.tracking['absence_max']:
            self.tracked_blobs_.pop(j)

    def _track_blob(self):
        """Greedy nearest-track association in blob order (reference pipeline.py:590-628):
        each blob takes the closest pre-existing track if it lies in that track's search
        ellipse, otherwise it starts a new track; unmatched tracks age."""
        blobs, tracks = self.blobs_, self.tracked_blobs_
        n_tracks = len(tracks)
        if n_tracks == 0:
            tracks.extend(self._new_track(b) for b in blobs)
            return
        if not blobs:
            for j in range(n_tracks - 1, -1, -1):
                self._miss(j)
            return
        dist = self._tracking_distance_matrix(tracks, blobs)
        # (a blob's row is only read for that blob: nearest tracks of all rows at once)
        nearest = dist.argmin(axis=1).tolist()
        inside = (dist[np.arange(len(blobs)), nearest] <= 1).tolist()
        paired = [False] * n_tracks
        for i, blob in enumerate(blobs):
            if inside[i]:
                paired[nearest[i]] = True
                self._pair(tracks[nearest[i]], blob)
            else:
                tracks.append(self._new_track(blob))
        # tracks no blob was paired with age.  The reference marks a pair with dist = -1 and ages the
        # tracks whose column satisfies all(dist >= 0): a column that already held a negative value
        # (int16 wrap-around of the squared differences) or a NaN is not aged either
        nonneg = (dist >= 0).all(axis=0).tolist()
        for j in range(n_tracks - 1, -1, -1):
            if nonneg[j] and not paired[j]:
                self._miss(j)

    # ------------------------------------------------------------------ steps 6-7
    def _post_filter(self):
        """Perimeter and perspective filters on tracked blobs (reference pipeline.py:734-762)."""
        pf = self.post_filter
        for track in self.tracked_blobs_:
            if point_in_polygon(pf['perimeter'], track['anchor']) < 0:
                track['filtered'].add('perimeter')
            else:
                track['filtered'].discard('perimeter')
            xy = track['contour'].reshape(-1, 2)
            lo, hi = xy.min(axis=0), xy.max(axis=0)  # cv.boundingRect of the contour
            width = int(hi[0]) - int(lo[0]) + 1
            height = int(hi[1]) - int(lo[1]) + 1
            if width * height <= self._predict_perspective('area', track['bottom'][1]) * pf['perspective_coeff']:
                track['filtered'].add('perspective')
            else:
                track['filtered'].discard('perspective')

    def _detect_blob(self):
        alarm = False
        for track in self.tracked_blobs_:
            if not track['filtered']:
                alarm = True
                if self.alarm_def['keep_alarm']:
                    track['already_in_alarm'] = True
            elif track['already_in_alarm']:
                alarm = True
        return alarm

    def _maintain_alarm(self, sys_in_alarm):
        """Hold a raised alarm for `duration_min` frames (reference pipeline.py:782-799)."""
        state = self._alarm_system
        state['duration'] -= 1
        if sys_in_alarm and not state['already_in_alarm']:
            state['duration'] = self.alarm_def['duration_min']
        if not sys_in_alarm and state['duration'] >= 0:
            sys_in_alarm = True
        if self.verbose and sys_in_alarm and not state['already_in_alarm']:
            print("Intrusion detected")
        state['already_in_alarm'] = sys_in_alarm
        return sys_in_alarm

    # ------------------------------------------------------------------ on-screen display
    def display(self, X, display={'tracking': False, 'perspective': False}):
        """Draw the perimeter and the tracked blobs into `X` in place (reference
        pipeline.py:801-853).  Needs OpenCV's drawing functions."""
        if cv is None:
            raise NotImplementedError("display() draws with OpenCV, which is not installed")
        cv.drawContours(X, [self.post_filter['perimeter']], 0, (25, 200, 200))
        for track in self.tracked_blobs_:
            flt = track['filtered']
            if not flt or track['already_in_alarm']:
                color = (0, 0, 255)
            elif 'perimeter' in flt:
                color = (255, 0, 0)
            elif 'perspective' in flt:
                color = (25, 200, 200)
            elif 'presence' in flt or 'distance' in flt:
                color = (0, 255, 0)
            else:
                raise ValueError(f'Unknown filtering type {flt}')
            cv.drawContours(X, [track['contour']], 0, color)
            cv.circle(X, [int(v) for v in track['anchor']], 4, color, -1)
            if display.get('tracking'):
                r_maj, r_min = self._tracking_radii(track['bottom'])
                cv.ellipse(X, [int(v) for v in track['center']], [int(r_maj), int(r_min)], 0, 0, 360,
                           color, 1)
            if display.get('perspective'):
                h = int(self._predict_perspective('height', track['bottom'][1]))
                w = int(self._predict_perspective('width', track['bottom'][1]))
                bx, by = int(track['bottom'][0]), int(track['bottom'][1])
                cv.rectangle(X, [bx - w // 2, by], [bx + w // 2, by - h], color, 1)
