"""Foreground estimators with the reference's `apply(X) -> mask` protocol
(pygofpid/segmentation.py:11-18).

`FrameDifferencing` runs on the GPU (fd_kernel of libgofpid_b200).  `ViBe` is outside the CUDA
hot path (it is not part of the accelerated configurations: its update draws from one
sequential NumPy random stream in raster order, which has no parallel equivalent with the same
results), so it stays a host-side NumPy estimator.
"""
from abc import ABCMeta, abstractmethod

import numpy as np


class ForegroundEstimator(metaclass=ABCMeta):
    BACKGROUND = 0
    FOREGROUND = 255

    @abstractmethod
    def apply(self, X):
        pass


class FrameDifferencing(ForegroundEstimator):
    """F = |X_t - X_{t-1}| > threshold, on the GPU (reference: segmentation.py:21-70).

    The first frame gives an all-background mask; 3-channel frames give per-channel masks
    averaged in uint8 arithmetic exactly as the reference's `np.mean(..., dtype=np.uint8)`."""

    def __init__(self, threshold=5, device=0):
        self.threshold = threshold
        self.device = device
        self._engine = None
        self._shape = None

    def apply(self, X):
        from .hotpath import HotPath
        X = np.asarray(X)
        if self._engine is None or X.shape != self._shape:
            shape = X.shape[:2] if (X.ndim == 3 and X.shape[2] == 1) else X.shape
            self._engine = HotPath(shape, blur=None, detector="FD", morph=(), find_blobs=False,
                                   fd_threshold=self.threshold, device=self.device, depth=2)
            self._shape = X.shape
        self._engine.process(X)
        return self._engine.mask()


class ViBe(ForegroundEstimator):
    """VIsual Background Extractor (Barnich & Van Droogenbroeck, 2010), host side.

    Same parameters and the same per-pixel procedure as the reference (segmentation.py:73-181):
    a pixel is background when at least `n_samples_close` of its `n_samples_per_pixel` model
    samples lie within `sphere_radius`; background pixels refresh one of their own samples and
    one sample of a random 8-neighbour, each with probability ~1/subsampling_factor."""

    def __init__(self, n_samples_per_pixel=20, sphere_radius=20, n_samples_close=2,
                 subsampling_factor=16, seed=42):
        self.n_samples_per_pixel = n_samples_per_pixel
        self.sphere_radius = sphere_radius
        self.n_samples_close = n_samples_close
        self.subsampling_factor = subsampling_factor
        self.seed = seed
        self._model = None
        self._rnd = None

    def _is_background(self, value, samples):
        close = 0
        for k in range(self.n_samples_per_pixel):
            if close >= self.n_samples_close:
                break
            diff = np.asarray(value, dtype=np.float64) - samples[..., k]
            if float(np.sqrt(np.sum(diff * diff))) < self.sphere_radius:
                close += 1
        return close >= self.n_samples_close

    def apply(self, X):
        X = np.asarray(X)
        H, W = X.shape[:2]
        if self._model is None:
            self._model = np.zeros(X.shape + (self.n_samples_per_pixel,))
        if self._rnd is None:
            self._rnd = np.random.RandomState(seed=self.seed)
        rnd, F = self._rnd, np.zeros((H, W), dtype=np.uint8)
        sub, ns = self.subsampling_factor - 1, self.n_samples_per_pixel - 1
        for y in range(H):
            for x in range(W):
                if not self._is_background(X[y, x], self._model[y, x]):
                    F[y, x] = self.FOREGROUND
                    continue
                if rnd.randint(0, sub) == 0:
                    self._model[y, x, ..., rnd.randint(0, ns)] = X[y, x]
                if rnd.randint(0, sub) == 0:
                    yn = min(max(y + rnd.choice([-1, 0, 1]), 0), H - 1)
                    xn = min(max(x + rnd.choice([-1, 0, 1]), 0), W - 1)
                    self._model[yn, xn, ..., rnd.randint(0, ns)] = X[y, x]
        self._F = F
        return F
