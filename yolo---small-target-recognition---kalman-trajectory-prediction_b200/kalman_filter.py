"""Mirror of ``ultralytics/trackers/utils/kalman_filter.py`` (KalmanFilterXYAH :39-286, KalmanFilterXYWH
:289-493) over the batched CUDA kernels of csrc/kf_ultra.cu.

Same method names, argument meaning and return shapes as the reference.  numpy inputs are staged to the
GPU and back (drop-in use); CUDA tensors are processed in place / returned as CUDA tensors (pipeline
use).  State is float32 on the device (the reference is float64 numpy).
"""
from __future__ import annotations

import numpy as np

from . import _lib


class _KalmanFilterBase:
    _kind = 0
    ndim = 4

    def __init__(self):
        _lib.require_cuda()
        self.lib = _lib.load()
        self._std_weight_position = 1.0 / 20
        self._std_weight_velocity = 1.0 / 160

    # ---- helpers
    @staticmethod
    def _dev(a, shape=None):
        import torch

        if isinstance(a, torch.Tensor):
            t = a.to("cuda", torch.float32).contiguous()
        else:
            t = torch.as_tensor(np.ascontiguousarray(a, np.float32), device="cuda")
        return t.reshape(shape) if shape is not None else t

    @staticmethod
    def _back(t, like, shape=None):
        import torch

        if isinstance(like, torch.Tensor):
            return t.reshape(shape) if shape is not None else t
        a = t.cpu().numpy().astype(np.float64)
        return a.reshape(shape) if shape is not None else a

    # ---- reference API
    def initiate(self, measurement):
        """(4,) -> mean (8,), covariance (8, 8)   [kalman_filter.py:62-97 / :304-361]"""
        import torch

        z = self._dev(measurement, (-1, 4))
        n = z.shape[0]
        mean = torch.empty((n, 8), dtype=torch.float32, device="cuda")
        cov = torch.empty((n, 8, 8), dtype=torch.float32, device="cuda")
        _lib.check(self.lib.b2_kf_initiate(self._kind, _lib.ptr(z), _lib.ptr(mean), _lib.ptr(cov), n, _lib.stream_ptr()))
        single = np.ndim(measurement) == 1
        return (self._back(mean, measurement, (8,) if single else None),
                self._back(cov, measurement, (8, 8) if single else None))

    def multi_predict(self, mean, covariance):
        """(N, 8), (N, 8, 8) -> predicted copies   [kalman_filter.py:165-203 / :435-470]"""
        m = self._dev(mean, (-1, 8)).clone()
        c = self._dev(covariance, (-1, 8, 8)).clone()
        _lib.check(self.lib.b2_kf_predict(self._kind, _lib.ptr(m), _lib.ptr(c), m.shape[0], _lib.stream_ptr()))
        return self._back(m, mean), self._back(c, covariance)

    def predict(self, mean, covariance):
        """(8,), (8, 8)   [kalman_filter.py:99-134 / :363-398]"""
        m, c = self.multi_predict(np.asarray(mean)[None] if not hasattr(mean, "is_cuda") else mean[None],
                                  np.asarray(covariance)[None] if not hasattr(covariance, "is_cuda") else covariance[None])
        return m[0], c[0]

    def project(self, mean, covariance):
        """(8,), (8, 8) -> (4,), (4, 4)   [kalman_filter.py:136-163 / :400-433]"""
        import torch

        m = self._dev(mean, (-1, 8))
        c = self._dev(covariance, (-1, 8, 8))
        n = m.shape[0]
        pm = torch.empty((n, 4), dtype=torch.float32, device="cuda")
        pc = torch.empty((n, 4, 4), dtype=torch.float32, device="cuda")
        _lib.check(self.lib.b2_kf_project(self._kind, _lib.ptr(m), _lib.ptr(c), _lib.ptr(pm), _lib.ptr(pc), n, _lib.stream_ptr()))
        single = np.ndim(mean) == 1
        return self._back(pm, mean, (4,) if single else None), self._back(pc, mean, (4, 4) if single else None)

    def update(self, mean, covariance, measurement, mask=None):
        """(8,), (8, 8), (4,) -> corrected (mean, covariance)   [kalman_filter.py:205-238]; batched over a leading N."""
        import torch

        m = self._dev(mean, (-1, 8)).clone()
        c = self._dev(covariance, (-1, 8, 8)).clone()
        z = self._dev(measurement, (-1, 4))
        mk = None if mask is None else torch.as_tensor(np.asarray(mask, np.uint8), device="cuda")
        _lib.check(self.lib.b2_kf_update(self._kind, _lib.ptr(m), _lib.ptr(c), _lib.ptr(z), _lib.ptr(mk), m.shape[0], _lib.stream_ptr()))
        single = np.ndim(mean) == 1
        return self._back(m, mean, (8,) if single else None), self._back(c, covariance, (8, 8) if single else None)

    def gating_distance(self, mean, covariance, measurements, only_position=False, metric="maha"):
        """(8,), (8, 8), (M, 4) -> (M,) squared distances   [kalman_filter.py:240-286]; (N, M) for batched state."""
        import torch

        if metric not in ("maha", "gaussian"):
            raise ValueError("Invalid distance metric")
        m = self._dev(mean, (-1, 8))
        c = self._dev(covariance, (-1, 8, 8))
        z = self._dev(measurements, (-1, 4))
        n, k = m.shape[0], z.shape[0]
        out = torch.empty((n, k), dtype=torch.float32, device="cuda")
        _lib.check(self.lib.b2_kf_gating(self._kind, _lib.ptr(m), _lib.ptr(c), n, _lib.ptr(z), k, int(bool(only_position)),
                                         0 if metric == "maha" else 1, _lib.ptr(out), _lib.stream_ptr()))
        single = np.ndim(mean) == 1
        return self._back(out, mean, (k,) if single else None)


class KalmanFilterXYAH(_KalmanFilterBase):
    """Centre x, y, aspect ratio a, height h (+ velocities)."""
    _kind = 0


class KalmanFilterXYWH(_KalmanFilterBase):
    """Centre x, y, width w, height h (+ velocities)."""
    _kind = 1
