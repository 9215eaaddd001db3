"""Deterministic synthetic video clips (the inputs of bench.py and of the parity tests).

Smooth sinusoidal background, i.i.d. Gaussian sensor noise, moving brighter rectangles with
integer velocities that wrap around, and an optional darker vertical band that sweeps across
the frame (exercises the shadow branch of MOG2 / KNN).  Everything derives from
`np.random.RandomState(seed)`, so a clip is fully described by its arguments.
"""
import numpy as np


class SyntheticClip:
    def __init__(self, H, W, C=1, seed=0, n_obj=6, noise_sigma=4.0, obj_size=None, shadow_band=True,
                 obj_gain=70.0, rng="legacy"):
        """rng='legacy': all randomness from one RandomState stream, frames in order (the parity tests
        and the golden vectors).  rng='pcg': the noise of frame t comes from its own
        np.random.default_rng([seed, t]) in float32 -- several times faster and frames can be
        produced in any order / in parallel (`frame_at`); what bench.py uses for its large clips."""
        self.H, self.W, self.C = int(H), int(W), int(C)
        self.seed, self.rng = int(seed), rng
        self.rs = np.random.RandomState(seed)
        yy, xx = np.mgrid[0:H, 0:W]
        bg = 96.0 + 48.0 * np.sin(xx / 37.0) + 32.0 * np.cos(yy / 23.0)
        if C == 3:
            bg = np.stack([bg, 0.9 * bg + 10.0, 0.8 * bg + 20.0], axis=-1)
        elif C != 1:
            raise ValueError("C must be 1 or 3")
        self.bg = bg.astype(np.float32)
        self.noise_sigma = float(noise_sigma)
        self.obj_gain = float(obj_gain)
        self.shadow_band = shadow_band
        if obj_size is None:
            obj_size = (max(6, W // 24), max(7, W // 8), max(6, H // 20), max(7, H // 6))
        w0, w1, h0, h1 = obj_size
        self.objs = []
        for _ in range(n_obj):
            self.objs.append((self.rs.randint(0, W), self.rs.randint(0, H),
                              self.rs.randint(w0, max(w0 + 1, w1)), self.rs.randint(h0, max(h0 + 1, h1)),
                              self.rs.randint(-6, 7), self.rs.randint(-6, 7)))
        self.t = 0

    def next(self):
        f = self.frame_at(self.t)
        self.t += 1
        return f

    def frame_at(self, t):
        if self.rng == "pcg":
            noise = np.random.default_rng([self.seed, int(t)]).standard_normal(self.bg.shape, dtype=np.float32)
        elif t != self.t:
            raise ValueError("rng='legacy' produces frames in order only")
        else:
            noise = self.rs.standard_normal(self.bg.shape).astype(np.float32)
        f = self.bg + noise * self.noise_sigma
        for (x, y, w, h, vx, vy) in self.objs:
            x0, y0 = (x + vx * t) % self.W, (y + vy * t) % self.H
            f[y0:y0 + h, x0:x0 + w] += self.obj_gain
        if self.shadow_band:
            b0 = (3 * t) % self.W
            f[:, b0:b0 + max(4, self.W // 64)] *= 0.7
        return np.clip(f, 0, 255).astype(np.uint8)

    def frames(self, n):
        return [self.next() for _ in range(n)]
