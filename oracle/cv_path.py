"""The reference's own hot path, call by call, on OpenCV (TEST / BASELINE INFRASTRUCTURE ONLY).

PyGOFPID's foreground-extraction arithmetic *is* OpenCV; this module issues the same cv2 calls
in the same order as ``GOFPID.detect`` lines pipeline.py:461-479 (+ _find_blob :555-570,
_create_blob :572-588, helpers.find_contours :125-150, get_bottom/get_center :153-168,
FrameDifferencing.apply segmentation.py:40-70).  It is what ``bench.py`` times as the CPU
baseline (kind "port": the reference's glue restated, the reference's arithmetic library
unchanged) and what the differential tests compare the CUDA path with on sizes where the C
restatement would be slow.  It must never be imported by ``pygofpid_b200``.
"""
import numpy as np

try:  # cv2 ships in the image; keep the import soft so the C oracle stays usable without it
    import cv2 as cv
except Exception:  # pragma: no cover
    cv = None


class CvFrameDifferencing:
    """segmentation.py:21-70: |X_t - X_{t-1}| > threshold on cv2."""

    def __init__(self, threshold=5):
        self.threshold = threshold
        self._prev = None

    def apply(self, X):
        if self._prev is None:
            F = np.zeros_like(X)
        else:
            F = cv.threshold(cv.absdiff(X, self._prev), self.threshold, 255, cv.THRESH_BINARY)[1]
        self._prev = X
        if F.ndim > 2:
            F = np.mean(F, axis=-1, dtype=np.uint8, keepdims=False)
        return F


class CvHotPath:
    """blur -> foreground detection -> morphology -> blobs, exactly as the reference wires cv2.

    ``blur``: None or dict(fun=cv.GaussianBlur|cv.blur, ksize, borderType) -- note the third
    value is passed *positionally* (pipeline.py:463-467), so it is sigmaX for GaussianBlur.
    ``frg_detect``: 'MOG2' | 'KNN' | 'FD'.  ``mat_morph``: list of dict(fun, kernel) or None.
    """

    def __init__(self, blur="default", frg_detect="MOG2", mat_morph="default", anchor="bottom",
                 area_min=100, fd_threshold=5):
        if blur == "default":
            blur = dict(fun=cv.GaussianBlur, ksize=(3, 3), borderType=cv.BORDER_DEFAULT)
        if mat_morph == "default":
            se = cv.getStructuringElement(cv.MORPH_RECT, (5, 5))
            mat_morph = [dict(fun=cv.erode, kernel=se), dict(fun=cv.dilate, kernel=se)]
        self.blur, self.mat_morph, self.anchor, self.area_min = blur, mat_morph, anchor, area_min
        if frg_detect == "MOG2":
            self.detector = cv.createBackgroundSubtractorMOG2()
        elif frg_detect == "KNN":
            self.detector = cv.createBackgroundSubtractorKNN()
        elif frg_detect == "FD":
            self.detector = CvFrameDifferencing(fd_threshold)
        else:
            raise ValueError("Unknown method for foreground detection")
        self.stage_ms = {}

    def blurred(self, X):
        if self.blur:
            X = self.blur["fun"](X, self.blur["ksize"], self.blur["borderType"])
        return X

    def mask(self, X):
        m = self.detector.apply(self.blurred(X))
        if self.mat_morph:
            for d in self.mat_morph:
                m = d["fun"](m, d["kernel"])
        return m

    @staticmethod
    def contours(mask):
        return cv.findContours(mask, mode=cv.RETR_EXTERNAL, method=cv.CHAIN_APPROX_NONE)[0]

    def blobs(self, mask):
        out = []
        for c in self.contours(mask):
            if cv.contourArea(c) < self.area_min:
                continue
            mom = cv.moments(c)
            x = int(mom["m10"] / mom["m00"])
            bottom = np.array([x, np.max(c[..., 1])], dtype=np.int16)
            mom = cv.moments(c)
            center = np.array([int(mom["m10"] / mom["m00"]), int(mom["m01"] / mom["m00"])],
                              dtype=np.int16)
            out.append(dict(contour=c, anchor=bottom if self.anchor == "bottom" else center,
                            bottom=bottom, center=center))
        return out

    def step(self, X):
        """One frame through the whole hot path: returns (mask, blobs)."""
        m = self.mask(X)
        return m, self.blobs(m)
