"""ctypes bindings of oracle/gofpid_oracle.c (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Each wrapper cites the reference call site the C function restates.
"""
import ctypes as ct
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgofpid_oracle.so")
_lib = None

u8p = ct.POINTER(ct.c_uint8)
u16p = ct.POINTER(ct.c_uint16)
f32p = ct.POINTER(ct.c_float)
i32p = ct.POINTER(ct.c_int32)


def build(force=False):
    """Compile the oracle library with gcc (idempotent)."""
    src = os.path.join(_HERE, "gofpid_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libgofpid_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ct.CDLL(_LIB_PATH)
        L.orc_gauss_kernel_q8.argtypes = [ct.c_int, ct.c_double, u16p]
        L.orc_gauss_kernel_q8.restype = ct.c_int
        L.orc_blur_q8.argtypes = [u8p, ct.c_int, ct.c_int, ct.c_int, u16p, ct.c_int, u16p,
                                  ct.c_int, u8p]
        L.orc_box_u8.argtypes = [u8p, ct.c_int, ct.c_int, ct.c_int, ct.c_int, ct.c_int, u8p]
        L.orc_fd.argtypes = [u8p, u8p, ct.c_int, ct.c_int, ct.c_int, ct.c_int, u8p]
        L.orc_mog2_create.argtypes = [ct.c_int, ct.c_int, ct.c_int]
        L.orc_mog2_create.restype = ct.c_void_p
        L.orc_mog2_destroy.argtypes = [ct.c_void_p]
        L.orc_mog2_apply.argtypes = [ct.c_void_p, u8p, u8p]
        L.orc_mog2_background.argtypes = [ct.c_void_p, u8p]
        for name, rt in (("weight", f32p), ("var", f32p), ("mean", f32p), ("nmodes", u8p)):
            f = getattr(L, "orc_mog2_" + name)
            f.argtypes = [ct.c_void_p]
            f.restype = rt
        L.orc_mog2_nframes.argtypes = [ct.c_void_p]
        L.orc_mog2_nframes.restype = ct.c_long
        L.orc_rng_fill_u8.argtypes = [ct.POINTER(ct.c_uint64), u8p, ct.c_long, ct.c_int]
        L.orc_knn_create.argtypes = [ct.c_int, ct.c_int, ct.c_int, ct.c_uint64]
        L.orc_knn_create.restype = ct.c_void_p
        L.orc_knn_destroy.argtypes = [ct.c_void_p]
        L.orc_knn_apply.argtypes = [ct.c_void_p, u8p, u8p]
        L.orc_knn_model.argtypes = [ct.c_void_p]
        L.orc_knn_model.restype = u8p
        L.orc_knn_idx.argtypes = [ct.c_void_p, ct.c_int]
        L.orc_knn_idx.restype = u8p
        L.orc_knn_nxt.argtypes = [ct.c_void_p, ct.c_int]
        L.orc_knn_nxt.restype = u8p
        L.orc_knn_rng.argtypes = [ct.c_void_p]
        L.orc_knn_rng.restype = ct.c_uint64
        L.orc_knn_counter.argtypes = [ct.c_void_p, ct.c_int]
        L.orc_knn_counter.restype = ct.c_int
        L.orc_knn_periods.argtypes = [ct.c_long, ct.c_int, ct.POINTER(ct.c_int)]
        L.orc_morph.argtypes = [u8p, ct.c_int, ct.c_int, u8p, ct.c_int, ct.c_int, ct.c_int, u8p]
        L.orc_ccl8.argtypes = [u8p, ct.c_int, ct.c_int, i32p]
        _lib = L
    return _lib


def _p(a, t=u8p):
    return a.ctypes.data_as(t)


def _hwc(X):
    X = np.ascontiguousarray(X, dtype=np.uint8)
    if X.ndim == 2:
        return X, X.shape[0], X.shape[1], 1
    return X, X.shape[0], X.shape[1], X.shape[2]


# --------------------------------------------------------------------------- blur
def gauss_kernel_q8(n, sigma):
    """8.8 fixed-point Gaussian taps cv.GaussianBlur uses for u8 (pipeline.py:463-467)."""
    q = np.zeros(n, np.uint16)
    if lib().orc_gauss_kernel_q8(n, float(sigma), _p(q, u16p)) != 0:
        raise ValueError("bad Gaussian kernel size %r" % (n,))
    return q


def gaussian_blur(X, ksize, sigma):
    """cv.GaussianBlur(X, ksize, sigma) for u8, BORDER_REFLECT_101 (pipeline.py:463)."""
    X, H, W, C = _hwc(X)
    qx = gauss_kernel_q8(ksize[0], sigma)
    qy = gauss_kernel_q8(ksize[1], sigma)
    out = np.empty((H, W) if X.ndim == 2 or C == 1 else (H, W, C), np.uint8)
    lib().orc_blur_q8(_p(X), H, W, C, _p(qx, u16p), len(qx), _p(qy, u16p), len(qy), _p(out))
    return out


def box_blur(X, ksize):
    """cv.blur(X, ksize) for u8 (pipeline.py:463 with fun=cv.blur)."""
    X, H, W, C = _hwc(X)
    out = np.empty((H, W) if X.ndim == 2 or C == 1 else (H, W, C), np.uint8)
    lib().orc_box_u8(_p(X), H, W, C, int(ksize[0]), int(ksize[1]), _p(out))
    return out


# --------------------------------------------------------------------------- FD
class FrameDifferencing:
    """segmentation.py:21-70 restated (threshold floored for u8, uint8-wrapping mean)."""

    def __init__(self, threshold=5):
        self.thr = int(np.floor(threshold))
        self.prev = None

    def apply(self, X):
        X, H, W, C = _hwc(X)
        out = np.zeros((H, W), np.uint8)
        if self.prev is not None:
            lib().orc_fd(_p(X), _p(self.prev), H, W, C, self.thr, _p(out))
        self.prev = X.copy()
        return out


# --------------------------------------------------------------------------- MOG2
class MOG2:
    """cv.createBackgroundSubtractorMOG2() defaults + .apply (pipeline.py:232,470)."""

    def __init__(self, H, W, C):
        self.H, self.W, self.C = H, W, C
        self.h = lib().orc_mog2_create(H, W, C)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_mog2_destroy(self.h)
            self.h = None

    def apply(self, X):
        X, H, W, C = _hwc(X)
        assert (H, W, C) == (self.H, self.W, self.C)
        mask = np.empty((H, W), np.uint8)
        lib().orc_mog2_apply(self.h, _p(X), _p(mask))
        return mask

    def background(self):
        out = np.empty((self.H, self.W) if self.C == 1 else (self.H, self.W, self.C), np.uint8)
        lib().orc_mog2_background(self.h, _p(out))
        return out

    def planes(self):
        """Model as arrays: weight (H,W,5), var (H,W,5), mean (H,W,5,C), nmodes (H,W)."""
        n = self.H * self.W
        L = lib()
        w = np.ctypeslib.as_array(L.orc_mog2_weight(self.h), (n * 5,)).reshape(self.H, self.W, 5)
        v = np.ctypeslib.as_array(L.orc_mog2_var(self.h), (n * 5,)).reshape(self.H, self.W, 5)
        m = np.ctypeslib.as_array(L.orc_mog2_mean(self.h), (n * 5 * self.C,)).reshape(
            self.H, self.W, 5, self.C)
        k = np.ctypeslib.as_array(L.orc_mog2_nmodes(self.h), (n,)).reshape(self.H, self.W)
        return w.copy(), v.copy(), m.copy(), k.copy()


# --------------------------------------------------------------------------- KNN
def rng_fill_u8(state, total, rng_range):
    """cv::RNG::fill(UNIFORM, 0, range) on a continuous u8 array; returns (array, state)."""
    st = ct.c_uint64(state)
    arr = np.empty(total, np.uint8)
    lib().orc_rng_fill_u8(ct.byref(st), _p(arr), total, rng_range)
    return arr, st.value


def knn_periods(n, history=500):
    u = (ct.c_int * 3)()
    lib().orc_knn_periods(n, history, u)
    return tuple(u)


class KNN:
    """cv.createBackgroundSubtractorKNN() defaults + .apply (pipeline.py:234,470)."""

    def __init__(self, H, W, C, seed=0xffffffff):
        self.H, self.W, self.C = H, W, C
        self.h = lib().orc_knn_create(H, W, C, seed)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_knn_destroy(self.h)
            self.h = None

    def apply(self, X):
        X, H, W, C = _hwc(X)
        assert (H, W, C) == (self.H, self.W, self.C)
        mask = np.empty((H, W), np.uint8)
        lib().orc_knn_apply(self.h, _p(X), _p(mask))
        return mask

    def state(self):
        """samples (H,W,21,C+1), idx (3,H,W), nxt (3,H,W), counters (3,), rng state."""
        n = self.H * self.W
        L = lib()
        nd = self.C + 1
        s = np.ctypeslib.as_array(L.orc_knn_model(self.h), (n * 21 * nd,)).reshape(
            self.H, self.W, 21, nd).copy()
        idx = np.stack([np.ctypeslib.as_array(L.orc_knn_idx(self.h, t), (n,)).reshape(
            self.H, self.W).copy() for t in range(3)])
        nxt = np.stack([np.ctypeslib.as_array(L.orc_knn_nxt(self.h, t), (n,)).reshape(
            self.H, self.W).copy() for t in range(3)])
        cnt = np.array([L.orc_knn_counter(self.h, t) for t in range(3)])
        return s, idx, nxt, cnt, L.orc_knn_rng(self.h)


# --------------------------------------------------------------------------- morphology
def morph(mask, kernel, dilate):
    """cv.erode / cv.dilate (pipeline.py:475), grey-level, arbitrary SE, border ignored."""
    mask = np.ascontiguousarray(mask, np.uint8)
    se = np.ascontiguousarray(kernel, np.uint8)
    out = np.empty_like(mask)
    lib().orc_morph(_p(mask), mask.shape[0], mask.shape[1], _p(se), se.shape[0], se.shape[1],
                    1 if dilate else 0, _p(out))
    return out


def erode(mask, kernel):
    return morph(mask, kernel, False)


def dilate(mask, kernel):
    return morph(mask, kernel, True)


# --------------------------------------------------------------------------- labelling
def ccl8(mask):
    """8-connected labels, canonical: 1 + raster index of the component's first pixel."""
    mask = np.ascontiguousarray(mask, np.uint8)
    labels = np.empty(mask.shape, np.int32)
    lib().orc_ccl8(_p(mask), mask.shape[0], mask.shape[1], _p(labels, i32p))
    return labels


def component_stats(labels):
    """Per-component dict keyed by canonical label: area, bbox (x0,y0,x1,y1 inclusive)."""
    out = {}
    ys, xs = np.nonzero(labels)
    lab = labels[ys, xs]
    order = np.argsort(lab, kind="stable")
    lab, ys, xs = lab[order], ys[order], xs[order]
    if lab.size == 0:
        return out
    starts = np.flatnonzero(np.r_[True, lab[1:] != lab[:-1]])
    ends = np.r_[starts[1:], lab.size]
    for s, e in zip(starts, ends):
        out[int(lab[s])] = dict(area=int(e - s), x0=int(xs[s:e].min()), y0=int(ys[s:e].min()),
                                x1=int(xs[s:e].max()), y1=int(ys[s:e].max()))
    return out
