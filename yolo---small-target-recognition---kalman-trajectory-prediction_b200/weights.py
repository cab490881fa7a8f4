"""Weights for the detect path: seeded synthetic state_dicts, BN folding, bf16 packing.

The reference ships no checkpoint, and its default initialisation is degenerate
for inference (every anchor scores ~1e-4), so benchmarks and parity fixtures use a
seeded, numpy-only recipe that produces a ``state_dict`` with exactly the key names and shapes of
``DetectionModel(cfg).state_dict()`` (ultralytics/nn/tasks.py:374-428) -- it loads into the
reference model unchanged, and a real trained ``state_dict`` loads into this package the same way.

BN folding follows ``fuse_conv_and_bn`` (ultralytics/utils/torch_utils.py:255-286), invoked by
``AutoBackend(fuse=True)`` on the reference's predict path (engine/predictor.py:397-405).
"""
from __future__ import annotations

import os
import zlib

import numpy as np

from .cfg import BN_EPS, REG_MAX, conv_list

_SILU_M2 = 0.3558   # E[silu(z)^2], z ~ N(0,1): keeps pre-activation variance ~1 layer to layer


def _rng(seed, key):
    return np.random.default_rng([int(seed), zlib.crc32(key.encode())])


def synthetic_state_dict(spec, seed=0, calib="auto"):
    """Seeded random weights (numpy float32) keyed like the reference ``state_dict``.

    ``calib``: ``"auto"`` loads ``calib/<name>_nc<nc>_seed<seed>.npz`` if it exists (BN running
    statistics and head gains measured once by ``tools/calibrate_synthetic.py``), ``None`` leaves
    BN statistics at identity, or a dict of arrays.
    """
    sd = {}
    det = spec["layers"][-1]
    nc = det["nc"]
    for prefix, c1, c2, k, s, bn in conv_list(spec):
        fan_in = c1 * k * k
        if bn:
            g = _rng(seed, prefix + ".w")
            sd[prefix + ".conv.weight"] = (g.standard_normal((c2, c1, k, k)) / np.sqrt(_SILU_M2 * fan_in)).astype(np.float32)
            g = _rng(seed, prefix + ".bn")
            gamma = g.uniform(0.9, 1.1, c2)
            if ".m." in prefix and prefix.endswith(".cv2"):
                gamma *= 0.5                       # damp residual branches (conditioning under bf16)
            sd[prefix + ".bn.weight"] = gamma.astype(np.float32)
            sd[prefix + ".bn.bias"] = (0.1 * g.standard_normal(c2)).astype(np.float32)
            sd[prefix + ".bn.running_mean"] = np.zeros(c2, np.float32)
            sd[prefix + ".bn.running_var"] = np.ones(c2, np.float32)
            sd[prefix + ".bn.num_batches_tracked"] = np.zeros((), np.int64)
        else:
            g = _rng(seed, prefix + ".w")
            is_box = ".cv2." in prefix
            gain = 0.6 if is_box else 1.0
            sd[prefix + ".weight"] = (gain * g.standard_normal((c2, c1, 1, 1)) / np.sqrt(_SILU_M2 * fan_in)).astype(np.float32)
            if is_box:
                # DFL logits biased towards short distances: small boxes, as for IR small targets
                b = np.tile(-0.45 * np.arange(REG_MAX, dtype=np.float64), 4)
            else:
                level = int(prefix.split(".")[-2])
                b = np.full(nc, -3.2 + 0.15 * level)   # puts O(1e2) anchors per frame above conf=0.15
            sd[prefix + ".bias"] = b.astype(np.float32)
    sd[f"model.{det['i']}.dfl.conv.weight"] = np.arange(REG_MAX, dtype=np.float32).reshape(1, REG_MAX, 1, 1)
    if calib == "auto":
        path = calib_path(spec, seed)
        calib = dict(np.load(path)) if os.path.exists(path) else None
    if calib:
        for k_, v in calib.items():
            if k_ in sd:
                sd[k_] = np.asarray(v, sd[k_].dtype).reshape(sd[k_].shape)
    return sd


def calib_path(spec, seed):
    return os.path.join(os.path.dirname(__file__), "calib", f"{spec['name']}_nc{spec['nc']}_seed{seed}.npz")


def to_numpy_state_dict(sd):
    """Accept a torch ``state_dict`` (or a checkpoint's ``model.state_dict()``) and return numpy arrays."""
    out = {}
    for k, v in sd.items():
        if hasattr(v, "detach"):
            v = v.detach().float().cpu().numpy() if v.dtype.is_floating_point else v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    return out


def fold_conv_bn(sd, prefix):
    """(w', b') with BN folded: w' = w * gamma/sqrt(var+eps), b' = beta - mean*gamma/sqrt(var+eps)."""
    w = np.asarray(sd[prefix + ".conv.weight"], np.float32)
    scale = np.asarray(sd[prefix + ".bn.weight"], np.float32) / np.sqrt(
        np.asarray(sd[prefix + ".bn.running_var"], np.float32) + np.float32(BN_EPS))
    b = np.asarray(sd[prefix + ".bn.bias"], np.float32) - np.asarray(sd[prefix + ".bn.running_mean"], np.float32) * scale
    if prefix + ".conv.bias" in sd:
        b = b + np.asarray(sd[prefix + ".conv.bias"], np.float32) * scale
    return (w * scale[:, None, None, None]).astype(np.float32), b.astype(np.float32)


def folded(sd, prefix, bn):
    if bn:
        return fold_conv_bn(sd, prefix)
    return np.asarray(sd[prefix + ".weight"], np.float32), np.asarray(sd[prefix + ".bias"], np.float32)


def f32_to_bf16_bits(a):
    """Round-to-nearest-even fp32 -> bf16 bit patterns (uint16)."""
    u = np.ascontiguousarray(a, np.float32).view(np.uint32)
    r = ((u >> np.uint32(16)) & np.uint32(1)) + np.uint32(0x7FFF)
    return ((u + r) >> np.uint32(16)).astype(np.uint16)


def bf16_bits_to_f32(b):
    return (np.asarray(b, np.uint16).astype(np.uint32) << np.uint32(16)).view(np.float32)


def pack_ohwi(w):
    """[O, I, kh, kw] -> [O, kh, kw, I] (GEMM-K index = (kh*k + kw)*I + c, K-major rows per output channel)."""
    return np.ascontiguousarray(np.transpose(w, (0, 2, 3, 1)))


def pack_stem(w):
    """[C0, 3, kh, kw] fp32 (RGB input order) -> [C0][32] bf16 bits, GEMM-K index (kh*3+kw)*3 + c, zero padded."""
    c0 = w.shape[0]
    k = np.zeros((c0, 32), np.float32)
    k[:, :27] = np.transpose(w, (0, 2, 3, 1)).reshape(c0, 27)
    return f32_to_bf16_bits(k)
