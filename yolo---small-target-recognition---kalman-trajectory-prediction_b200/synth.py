"""Seeded synthetic inputs: infrared-like frames with small moving blobs, and
multi-target detection sequences for the tracker.  numpy only; used by bench.py, the tests, the
calibration tool and the golden-vector generator so that every side sees identical data.
"""
from __future__ import annotations

import numpy as np


class IRStream:
    """One synthetic 8-bit infrared video stream: a fixed noise background (gray replicated to 3
    channels, like the project's IR footage) plus K bright Gaussian blobs of 3-12 px moving at <= 3 px/frame."""

    def __init__(self, seed=0, h=512, w=640, n_targets=20):
        g = np.random.default_rng(seed)
        self.h, self.w = h, w
        self.bg = g.integers(0, 96, (h, w), dtype=np.uint8)
        self.pos = np.stack([g.uniform(20, w - 20, n_targets), g.uniform(20, h - 20, n_targets)], 1)
        self.vel = g.uniform(-3, 3, (n_targets, 2))
        self.size = g.uniform(3, 12, n_targets)
        self.amp = g.uniform(120, 159, n_targets)
        self.t = 0

    def frame(self):
        """Next frame, (h, w, 3) uint8 BGR (all channels equal)."""
        img = self.bg.astype(np.float32)
        for (cx, cy), s, a in zip(self.pos, self.size, self.amp):
            r = int(3 * s) + 1
            x0, x1 = max(0, int(cx) - r), min(self.w, int(cx) + r + 1)
            y0, y1 = max(0, int(cy) - r), min(self.h, int(cy) + r + 1)
            if x0 >= x1 or y0 >= y1:
                continue
            yy, xx = np.mgrid[y0:y1, x0:x1]
            img[y0:y1, x0:x1] += a * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * (s / 2.0) ** 2))
        self.pos += self.vel
        for d, lim in ((0, self.w), (1, self.h)):
            out = (self.pos[:, d] < 10) | (self.pos[:, d] > lim - 10)
            self.vel[out, d] *= -1
        self.t += 1
        g = np.clip(img, 0, 255).astype(np.uint8)
        return np.ascontiguousarray(np.stack([g, g, g], -1))


def noise_frames(seed, b, h, w):
    """(b, h, w, 3) uint8 uniform-noise gray frames (SURVEY.md 8d, C1 recipe)."""
    g = np.random.default_rng(seed).integers(0, 256, (b, h, w), dtype=np.uint8)
    return np.ascontiguousarray(np.stack([g, g, g], -1))


class DetectionSequence:
    """Synthetic per-frame detections for one stream: targets on straight lines with measurement
    noise, Bernoulli misses (occlusion bursts) and clutter.  Returns float32 rows [x1,y1,x2,y2,conf]."""

    def __init__(self, seed=0, n_targets=8, w=640, h=512, p_detect=0.8, clutter=0.3, burst=(40, 70)):
        g = np.random.default_rng(seed)
        self.g, self.w, self.h = g, w, h
        self.pos = np.stack([g.uniform(40, w - 40, n_targets), g.uniform(40, h - 40, n_targets)], 1)
        self.vel = g.normal(0, 1.5, (n_targets, 2))
        self.size = g.uniform(4, 24, (n_targets, 2))
        self.p_detect, self.clutter, self.burst = p_detect, clutter, burst
        self.t = 0

    def step(self):
        g = self.g
        self.pos += self.vel
        rows = []
        in_burst = self.burst[0] <= self.t < self.burst[1]
        for i, ((cx, cy), (bw, bh)) in enumerate(zip(self.pos, self.size)):
            hidden = in_burst and i % 2 == 0
            if hidden or g.random() > self.p_detect:
                continue
            nx, ny = g.normal(0, 1.0, 2)
            rows.append([cx + nx - bw / 2, cy + ny - bh / 2, cx + nx + bw / 2, cy + ny + bh / 2, g.uniform(0.2, 0.95)])
        for _ in range(g.poisson(self.clutter)):
            cx, cy = g.uniform(0, self.w), g.uniform(0, self.h)
            bw, bh = g.uniform(4, 24, 2)
            rows.append([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2, g.uniform(0.15, 0.5)])
        self.t += 1
        if rows:
            order = g.permutation(len(rows))
            return np.asarray(rows, np.float32)[order]
        return np.zeros((0, 5), np.float32)
