"""Host-side helpers of the GOFPID pipeline (geometry, coordinates, small numerics).

API mirror of pygofpid/helpers.py of the reference (same names, argument meaning and error
behaviour) so that code written against the reference keeps working; the implementations are
independent NumPy code and do not need OpenCV.  Blob creation itself runs on the GPU
(`find_contours` drives the CUDA labelling + border tracing of libgofpid_b200).
"""
import threading

import numpy as np


# ------------------------------------------------------------------------------ files
def change_extension(filename_input, ext_new="png"):
    """Replaces what follows the last dot -- every occurrence of it, and the whole name when there is
    no dot -- by `ext_new`, exactly as the reference does (helpers.py:7-10)."""
    return filename_input.replace(filename_input.split('.')[-1], ext_new)


def read_first_frame(video_filename):
    """First frame of a video, also written next to it as an image (helpers.py:13-23)."""
    import cv2 as cv  # video decoding is outside the hot path
    cap = cv.VideoCapture(video_filename)
    if not cap.isOpened():
        raise ValueError("Unable to open video %s." % video_filename)
    _, frame = cap.read()
    cap.release()
    cv.imwrite(change_extension(video_filename), frame)
    return frame


# ------------------------------------------------------------------------------ polygons
def point_in_polygon(polygon, point):
    """+1 inside, 0 on the boundary, -1 outside (what cv.pointPolygonTest(..., False) answers).

    Exact integer / float arithmetic on the polygon's vertices (even-odd rule)."""
    poly = np.asarray(polygon, dtype=np.float64).reshape(-1, 2).tolist()  # (plain floats: same doubles, faster)
    px, py = float(point[0]), float(point[1])
    n = len(poly)
    inside = False
    for i in range(n):
        x0, y0 = poly[i - 1]
        x1, y1 = poly[i]
        # on the segment?
        cross = (px - x0) * (y1 - y0) - (py - y0) * (x1 - x0)
        if cross == 0 and min(x0, x1) <= px <= max(x0, x1) and min(y0, y1) <= py <= max(y0, y1):
            return 0
        if (y0 > py) != (y1 > py):
            xi = x0 + (py - y0) * (x1 - x0) / (y1 - y0)
            if xi > px:
                inside = not inside
    return 1 if inside else -1


def _thick_square(point, thickness):
    tx, ty = thickness[0], thickness[1]
    return np.array([[point[0] - tx, point[1] - ty], [point[0] - tx, point[1] + ty],
                     [point[0] + tx, point[1] + ty], [point[0] + tx, point[1] - ty]])


def find_point(coord, points, thickness):
    """Index of the point whose `thickness` box contains `coord`, else -1 (helpers.py:45-75)."""
    for i, point in enumerate(points):
        if point_in_polygon(_thick_square(point, thickness), coord) >= 0:
            return i
    return -1


def find_line(coord, points, thickness):
    """Index of the closed polyline's segment selected by `coord`, else -1 (helpers.py:78-122).

    A segment is selected when `coord` falls in the diamond spanned by its end points and its
    mid point pushed out by `thickness` along the normal."""
    n = len(points)
    if n < 2:
        return -1
    pts = np.asarray(points)
    for i in range(n):
        a, b = pts[i % n], pts[(i + 1) % n]
        mid = (a + b) / 2
        d = b - a
        d = d / np.sqrt(float(d[0]) ** 2 + float(d[1]) ** 2)
        quad = np.array([[a[0], a[1]],
                         [mid[0] - d[1] * thickness[0], mid[1] + d[0] * thickness[1]],
                         [b[0], b[1]],
                         [mid[0] + d[1] * thickness[0], mid[1] - d[0] * thickness[1]]], dtype=np.int32)
        if point_in_polygon(quad, coord) >= 0:
            return i
    return -1


def contour_moments(contour):
    """(m00, m10, m01) of a closed integer contour, as cv.moments computes them."""
    pts = np.asarray(contour, dtype=np.float64).reshape(-1, 2)
    prev = np.roll(pts, 1, axis=0)
    dxy = prev[:, 0] * pts[:, 1] - pts[:, 0] * prev[:, 1]
    a00 = dxy.sum()
    a10 = (dxy * (prev[:, 0] + pts[:, 0])).sum()
    a01 = (dxy * (prev[:, 1] + pts[:, 1])).sum()
    if abs(a00) <= 1.1920928955078125e-07:
        return 0.0, 0.0, 0.0
    if a00 > 0:
        half, sixth = 0.5, 0.16666666666666666666666666666667
    else:
        half, sixth = -0.5, -0.16666666666666666666666666666667
    return a00 * half, a10 * sixth, a01 * sixth


def get_bottom(contour, dtype=np.int16):
    """Middle-bottom point of a contour: x of the centroid, largest y (helpers.py:153-159)."""
    m00, m10, _ = contour_moments(contour)
    x = int(m10 / m00)
    y = np.max(np.asarray(contour)[..., 1])
    return np.array([x, y], dtype=dtype)


def get_center(contour, dtype=np.int16):
    """Centroid of a contour from its polygon moments (helpers.py:162-168)."""
    m00, m10, m01 = contour_moments(contour)
    return np.array([int(m10 / m00), int(m01 / m00)], dtype=dtype)


_contour_engines = {}
_contour_lock = threading.Lock()


def find_contours(mask, device=0):
    """Outer contours of the 8-connected components of `mask` that are not enclosed by another
    component -- cv.findContours(mask, RETR_EXTERNAL, CHAIN_APPROX_NONE) (helpers.py:125-150) --
    computed on the GPU `device` (labelling + border tracing of libgofpid_b200).  One engine per
    (shape, device) is kept and shared under a lock."""
    from .hotpath import HotPath
    mask = np.ascontiguousarray(mask)
    if mask.ndim == 3 and mask.shape[2] == 1:
        mask = mask[..., 0]
    if mask.ndim != 2 or mask.dtype != np.uint8:
        raise ValueError("mask must be a 2-D uint8 array")
    key = (mask.shape, int(device))
    with _contour_lock:
        hp = _contour_engines.get(key)
        if hp is None:
            # |mask - 0| > 0 turns the mask into its own foreground mask; area_min <= 0 keeps all
            # (one synchronous frame at a time: two slots are enough)
            hp = HotPath(mask.shape, blur=None, detector="FD", morph=(), fd_threshold=0, area_min=-1.0,
                         device=device, depth=2)
            _contour_engines.clear()
            _contour_engines[key] = hp
        hp.reset()
        hp.process(np.zeros_like(mask))
        hp.process(mask)
        recs, pts = hp.raw_blobs()
        return tuple(pts[b.point_offset:b.point_offset + b.n_points].reshape(-1, 1, 2).copy() for b in recs)


# ------------------------------------------------------------------------------ coordinates
def normalize_coords(coords, shape, dtype=np.float64):
    """Image coordinates (x, y) -> normalised coordinates in [0, 1] (helpers.py:171-177)."""
    out = np.array(coords, dtype=dtype)
    out[..., 0] /= shape[1]
    out[..., 1] /= shape[0]
    return out


def unnormalize_coords(ncoords, shape, dtype=np.uint16):
    """Normalised coordinates in [0, 1] -> image coordinates (helpers.py:180-186)."""
    out = np.array(ncoords, dtype=np.float64)
    out[..., 0] *= shape[1]
    out[..., 1] *= shape[0]
    return out.astype(dtype)


def dist_euclidean(a, b):
    """Euclidean distance between scalars or vectors (helpers.py:189-193)."""
    if isinstance(a, (int, float)):
        return abs(a - b)
    return np.linalg.norm(a - b)


def cdist_euclidean(XA, XB):
    """Pairwise Euclidean distances, shape (len(XA), len(XB)) (helpers.py:196-207)."""
    XA, XB = np.asarray(XA, dtype=np.float64), np.asarray(XB, dtype=np.float64)
    if XA.ndim == 1:
        XA, XB = XA[:, None], XB[:, None]
    diff = XA[:, None, :] - XB[None, :, :]
    return np.atleast_2d(np.sqrt((diff * diff).sum(-1)))


class SimpleLinearRegression():
    """y = a * x + b through two samples (helpers.py:210-285)."""

    def __init__(self):
        self.coeff = None
        self.intercept = None

    def fit(self, x, y):
        if len(x) != 2:
            raise ValueError("Fit only two samples.")
        if len(x) != len(y):
            raise ValueError("Inputs must have the same size.")
        self.coeff = (y[1] - y[0]) / (x[1] - x[0])
        self.intercept = y[0] - self.coeff * x[0]
        return self

    def predict(self, x):
        return self.coeff * np.asarray(x) + self.intercept

    def predict_clip(self, x, a_min=None, a_max=None):
        return np.clip(self.predict(x), a_min=a_min, a_max=a_max)
