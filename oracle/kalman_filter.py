"""Oracle (numpy float64, CPU) for the Ultralytics ByteTrack/BoT-SORT Kalman filters -- test infrastructure.

Restates ultralytics/trackers/utils/kalman_filter.py: KalmanFilterXYAH (:39-286) and
KalmanFilterXYWH (:289-493) as batched functions over N tracks with dense 8x8 covariances.
Third-party arithmetic restated: scipy.linalg.cho_factor/cho_solve/solve_triangular (scipy>=1.4.1,
pyproject.toml:70; call sites kalman_filter.py:228-231, :283) -- Cholesky solve of the 4x4
innovation covariance, here via numpy.linalg.

State: [x, y, a|w, h, vx, vy, va|vw, vh]; std weights 1/20 (position) and 1/160 (velocity).
"""
from __future__ import annotations

import numpy as np

W_POS, W_VEL = 1.0 / 20, 1.0 / 160
_F = np.eye(8)
for _i in range(4):
    _F[_i, 4 + _i] = 1.0
_H = np.eye(4, 8)


def _scales(kind, mean):
    """The (N,4) length scale each std is proportional to: h for XYAH, (w,h,w,h) for XYWH."""
    mean = np.atleast_2d(mean)
    if kind == "xyah":
        return np.stack([mean[:, 3]] * 4, 1)
    return np.stack([mean[:, 2], mean[:, 3], mean[:, 2], mean[:, 3]], 1)


def initiate(kind, measurement):
    """initiate (:62-97 / :304-361). measurement (4,) -> mean (8,), covariance (8,8)."""
    z = np.asarray(measurement, np.float64)
    s = _scales(kind, np.r_[z, np.zeros(4)])[0]
    std = np.r_[2 * W_POS * s, 10 * W_VEL * s]
    if kind == "xyah":
        std[2], std[6] = 1e-2, 1e-5
    return np.r_[z, np.zeros(4)], np.diag(std ** 2)


def predict(kind, mean, cov):
    """predict / multi_predict (:99-134, :165-203 / :363-398, :435-470). mean (N,8) or (8,), cov (N,8,8) or (8,8)."""
    single = np.ndim(mean) == 1
    m = np.atleast_2d(np.asarray(mean, np.float64))
    P = np.asarray(cov, np.float64).reshape(-1, 8, 8)
    s = _scales(kind, m)
    std = np.concatenate([W_POS * s, W_VEL * s], 1)
    if kind == "xyah":
        std[:, 2], std[:, 6] = 1e-2, 1e-5
    Q = np.zeros_like(P)
    idx = np.arange(8)
    Q[:, idx, idx] = std ** 2
    m2 = m @ _F.T
    P2 = _F[None] @ P @ _F.T[None] + Q
    return (m2[0], P2[0]) if single else (m2, P2)


def project(kind, mean, cov):
    """project (:136-163 / :400-433). Single track."""
    m = np.asarray(mean, np.float64)
    s = _scales(kind, m)[0]
    std = W_POS * s
    if kind == "xyah":
        std[2] = 1e-1
    return _H @ m, _H @ np.asarray(cov, np.float64) @ _H.T + np.diag(std ** 2)


def update(kind, mean, cov, measurement):
    """update (:205-238): K = P H^T S^-1 via Cholesky, mean += K innov, P -= K S K^T."""
    m = np.asarray(mean, np.float64)
    P = np.asarray(cov, np.float64)
    pm, S = project(kind, m, P)
    L = np.linalg.cholesky(S)
    B = (P @ _H.T).T                                        # (4,8)
    K = np.linalg.solve(L.T, np.linalg.solve(L, B)).T       # (8,4)
    innov = np.asarray(measurement, np.float64) - pm
    return m + innov @ K.T, P - K @ S @ K.T


def gating_distance(kind, mean, cov, measurements, only_position=False, metric="maha"):
    """gating_distance (:240-286)."""
    pm, S = project(kind, mean, cov)
    z = np.asarray(measurements, np.float64)
    if only_position:
        pm, S, z = pm[:2], S[:2, :2], z[:, :2]
    d = z - pm
    if metric == "gaussian":
        return np.sum(d * d, axis=1)
    if metric != "maha":
        raise ValueError("Invalid distance metric")
    L = np.linalg.cholesky(S)
    y = np.linalg.solve(L, d.T)
    return np.sum(y * y, axis=0)
