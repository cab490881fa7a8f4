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


# ---- the "blob highway": a designed small-target detector embedded in the random network ------------------------------
# Random weights make a detector whose candidates are, by construction, the anchors closest to the confidence threshold:
# they flicker from frame to frame, bf16 rounding moves them in and out of the detection set, and the tracker downstream sees
# hundreds of unrelated boxes.  The recipe therefore reserves a few channels along the P2 path
#   model.0 -> model.1 -> C2f(model.2).cv1/.cv2 -> Concat -> C2f(P2 head).cv1/.cv2 -> Detect.cv3[0][0..2]
# (three from the first C2f on) for a hand-made bright-blob detector; every other weight stays seeded-random, so the arithmetic (dense convolutions of the
# same shapes) is unchanged.  Stages (pre-activation -> SiLU):
#   0  stem      h = G0 * mean3x3(pixel)                      local brightness (roughly linear range of SiLU)
#   1  model.1   h = G1 * mean3x3(h) + b1                     b1 puts the background at about -5.5 sigma (set by the calibration
#                                                             pass from the median / MAD of the pre-activation) -> blobs only
#   2  cv1       u_k = C - A_k h, k = 0, 1, 2                 three inverted copies with steep / medium / shallow slopes ...
#   3  cv2       v_k = C - u_k                                ... so that v_k ~ min(A_k h, C): a concave three-segment response
#   4..7         pass-through (1x1 taps / centre taps)
#   class 0 logit = sum_k W_k v_k + B                         ~ logarithmic in the blob energy h: small blobs clear conf = 0.15,
#                                                             large ones stay far from sigmoid saturation, and the peak anchor
#                                                             of a blob leads its neighbours by a margin that does not depend
#                                                             on the blob's brightness (stable NMS order under bf16)
# All other class logits are shrunk and biased far below the threshold by the calibration pass; the box branch keeps its
# random features around a DFL bias peaked at bin HW_BOX_BIN (boxes of ~8 * bin pixels at P2).
HW_G0, HW_G1 = 12.0, 3.0
HW_BG_SIGMAS = 5.5
HW_C, HW_A = 4.0, (1.0, 0.4, 0.1)
HW_LOGIT_BG, HW_LOGIT_AT = -4.2, ((2.8, 1.0), (10.0, 3.9), (30.0, 6.4))   # class-0 logit of the background; (h, logit) anchors
HW_BOX_BIN, HW_BOX_SLOPE = 3.0, 0.6


def _silu(x):
    return x / (1.0 + np.exp(-x))


def _highway_response(h, a):
    """Scalar response of stages 2..7 to a stage-1 activation h (float64)."""
    v = _silu(HW_C - _silu(HW_C - a * h))
    for _ in range(4):
        v = _silu(v)
    return v


def highway_path(spec):
    """[(state_dict prefix, kernel size, input channel of the highway)] for the eight stages, in execution order."""
    layers = {L["i"]: L for L in spec["layers"]}
    det = spec["layers"][-1]
    path, i = [], det["f"][0]
    while True:
        path.append(i)
        L = layers[i]
        if i == 0:
            break
        f = L["f"]
        if L["type"] == "Concat":
            srcs = [i - 1 if j == -1 else j for j in f]
            i = [j for j in srcs if layers[j]["type"] != "Upsample"][0]
        else:
            i = i - 1 if f == -1 else f
    out, ch = [], 0
    for i in reversed(path):
        L = layers[i]
        p = f"model.{i}"
        if L["type"] == "Conv":
            out.append((p, L["k"], ch)); ch = 0
        elif L["type"] == "C2f":
            out.append((p + ".cv1", 1, ch)); out.append((p + ".cv2", 1, 0)); ch = 0
        elif L["type"] == "Concat":
            srcs = [i - 1 if j == -1 else j for j in L["f"]]
            off = 0
            for j in srcs:
                if layers[j]["type"] != "Upsample":
                    break
                off += layers[j]["c_out"]
            ch += off
        else:
            raise ValueError(f"unexpected {L['type']} on the P2 path")
    dp = f"model.{det['i']}"
    out.append((dp + ".cv3.0.0", 3, ch)); out.append((dp + ".cv3.0.1", 3, 0))
    assert len(out) == 8, out
    return out


def highway_channels(spec):
    """{bn prefix: [channels]} whose BN statistics stay at identity (the calibration pass must not rescale them)."""
    st = highway_path(spec)
    return {p: ([0] if k < 2 else list(range(len(HW_A)))) for k, (p, _, _) in enumerate(st)}


def _install_highway(spec, sd, b1=-8.6):
    """b1: background bias of stage 1 (refined by the calibration pass)."""
    st = highway_path(spec)

    def row(prefix, r, incol, taps, beta):
        w = sd[prefix + ".conv.weight"]
        w[r] = 0
        w[r, incol] = taps
        sd[prefix + ".bn.weight"][r] = 1.0
        sd[prefix + ".bn.bias"][r] = beta

    p0, k0, _ = st[0]
    sd[p0 + ".conv.weight"][0] = HW_G0 / (3 * k0 * k0)
    sd[p0 + ".bn.weight"][0], sd[p0 + ".bn.bias"][0] = 1.0, 0.0
    p1, k1, c1 = st[1]
    row(p1, 0, c1, np.full((k1, k1), HW_G1 / (k1 * k1), np.float32), b1)
    p2, _, c2 = st[2]
    K = range(len(HW_A))
    for r in K:
        row(p2, r, c2, np.full((1, 1), -HW_A[r], np.float32), HW_C)
    p3, _, _ = st[3]
    for r in K:
        row(p3, r, r, np.full((1, 1), -1.0, np.float32), HW_C)
    for k in range(4, 8):
        pk, kk, ck = st[k]
        taps = np.zeros((kk, kk), np.float32)
        taps[kk // 2, kk // 2] = 1.0
        for r in K:
            row(pk, r, ck + r, taps, 0.0)
    # final linear map: one equation per anchor of HW_LOGIT_AT in the gains
    A = np.array([[_highway_response(h, a) for a in HW_A] for h, _ in HW_LOGIT_AT])
    g = np.linalg.solve(A, np.array([l - HW_LOGIT_BG for _, l in HW_LOGIT_AT]))
    det = spec["layers"][-1]
    dp = f"model.{det['i']}"
    w, b = sd[dp + ".cv3.0.2.weight"], sd[dp + ".cv3.0.2.bias"]
    w[0] = 0
    w[0, :len(g), 0, 0] = g
    b[0] = HW_LOGIT_BG


def synthetic_state_dict(spec, seed=0, calib="auto", bake=True):
    """Seeded synthetic weights (numpy float32) keyed like the reference ``state_dict``: random everywhere except the two
    blob-highway channels described above.

    ``calib``: ``"auto"`` loads ``calib/<name>_nc<nc>_seed<seed>.npz`` if it exists (BN running statistics, head gains and the
    background bias of the highway, measured once by ``tools/calibrate_synthetic.py``), ``None`` leaves them uncalibrated, or
    a dict of arrays.
    ``bake`` (default): the calibrated BN scale gamma / sqrt(var + eps) is multiplied into the conv weights, which are then
    rounded to bf16-representable values, and the BN is left as the identity scale (gamma = 1, var = 1 - eps, mean = 0; beta
    kept).  The fp32 reference and the bf16 engine then hold bit-identical weights -- what separates them is the rounding of
    the stored activations alone (SURVEY.md H1 iv).
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
                # DFL logits peaked at bin HW_BOX_BIN: boxes of about 2 * bin * stride pixels
                b = np.tile(-HW_BOX_SLOPE * np.abs(np.arange(REG_MAX, dtype=np.float64) - HW_BOX_BIN), 4)
            else:
                b = np.full(nc, -9.0)              # the random classes never fire (gains shrunk by the calibration pass)
            sd[prefix + ".bias"] = b.astype(np.float32)
    sd[f"model.{det['i']}.dfl.conv.weight"] = np.arange(REG_MAX, dtype=np.float32).reshape(1, REG_MAX, 1, 1)
    _install_highway(spec, sd)
    if calib == "auto":
        path = calib_path(spec, seed)
        calib = dict(np.load(path)) if os.path.exists(path) else None
    if calib:
        for k_, v in calib.items():
            if k_ in sd and not k_.endswith(".2.bias"):        # head biases are the recipe's, not measured
                sd[k_] = np.asarray(v, sd[k_].dtype).reshape(sd[k_].shape)
        # the highway rows are the recipe's as well (the calibrated head gains above cover whole weight matrices); only the
        # background bias of stage 1 is measured
        b1_key = highway_path(spec)[1][0] + ".bn.bias"
        _install_highway(spec, sd, float(calib[b1_key][0]) if b1_key in calib else -8.6)
    # BN running_var of the highway channels: exactly 1 - eps, so that the folded scale is exactly gamma = 1
    for prefix, chans in highway_channels(spec).items():
        for c in chans:
            sd[prefix + ".bn.running_var"][c] = 1.0 - BN_EPS
            sd[prefix + ".bn.running_mean"][c] = 0.0
    if bake:
        for prefix, c1, c2, k, s, bn in conv_list(spec):
            if bn:
                w, b = fold_conv_bn(sd, prefix)
                sd[prefix + ".conv.weight"] = bf16_bits_to_f32(f32_to_bf16_bits(w)).reshape(w.shape)
                sd[prefix + ".bn.weight"] = np.ones(c2, np.float32)
                sd[prefix + ".bn.bias"] = b
                sd[prefix + ".bn.running_mean"] = np.zeros(c2, np.float32)
                sd[prefix + ".bn.running_var"] = np.full(c2, 1.0 - BN_EPS, np.float32)
            else:
                w = sd[prefix + ".weight"]
                sd[prefix + ".weight"] = bf16_bits_to_f32(f32_to_bf16_bits(w)).reshape(w.shape)
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


# ---------------------------------------------------------------------------------------------------------------------
# channel padding: the tcgen05 conv path wants every tensor's channel count to be a multiple of 16
# ---------------------------------------------------------------------------------------------------------------------
def needs_padding(spec, multiple=16):
    return any(((c1 % multiple) and p != "model.0") or c2 % multiple for p, c1, c2, _, _, bn in conv_list(spec) if bn) or \
        any(c1 % multiple for _, c1, _, _, _, bn in conv_list(spec) if not bn)


def pad_channels(spec, sd, multiple=16):
    """Re-parameterise (spec, state_dict) as the SAME function with every hidden width rounded up to ``multiple``.

    The project's own model, ``yolov8-small.yaml`` at its default scale (train_small_targets.py:20), has C2f hidden widths of
    12 / 24 channels (cfg/models/v8/yolov8-small.yaml:12-16 with width 0.375).  Padding channels get zero conv rows, BN
    (gamma, beta, mean, var) = (1, 0, 0, 1 - eps) and zero columns in every consumer, so they carry SiLU(0) = 0 through
    shortcuts, max-pools, upsamples and concatenations; the real channels compute exactly what they did.  A C2f's ``cv1``
    output is two padded halves (chunk(2) of block.py:316 stays a split in the middle), a Concat is the concatenation of its
    padded inputs.  Returns (spec_padded, sd_padded); Detect's outputs (64 DFL bins, nc classes) are unchanged."""
    import copy

    def up(c):
        return (c + multiple - 1) // multiple * multiple

    sp = copy.deepcopy(spec)
    L0 = {L["i"]: L for L in spec["layers"]}
    out = {}
    omap, ophys = {}, {}                    # layer -> physical index of each logical output channel, physical channel count

    def src(i, f):
        return i - 1 if f == -1 else f

    def conv_bn(prefix, rows, n_rows, cols, n_cols):
        """rows / cols: physical index of each logical output / input channel."""
        w = np.asarray(sd[prefix + ".conv.weight"], np.float32)
        wp = np.zeros((n_rows, n_cols) + w.shape[2:], np.float32)
        wp[np.ix_(rows, cols)] = w
        out[prefix + ".conv.weight"] = wp
        for key, fill in ((".bn.weight", 1.0), (".bn.bias", 0.0), (".bn.running_mean", 0.0), (".bn.running_var", 1.0 - BN_EPS)):
            v = np.full(n_rows, fill, np.float32)
            v[rows] = np.asarray(sd[prefix + key], np.float32)
            out[prefix + key] = v
        out[prefix + ".bn.num_batches_tracked"] = np.asarray(sd.get(prefix + ".bn.num_batches_tracked", 0))

    def blocks(n_blocks, c, pc):
        return np.concatenate([j * pc + np.arange(c) for j in range(n_blocks)])

    for L, Lp in zip(spec["layers"], sp["layers"]):
        i, t, f = L["i"], L["type"], L["f"]
        p = f"model.{i}"
        if t in ("Conv", "C2f", "SPPF"):
            s_in = src(i, f)
            cols, n_cols = (np.arange(L["c1"]), L["c1"]) if i == 0 else (omap[s_in], ophys[s_in])
            c2, pc2 = L["c2"], up(L["c2"])
            if t == "Conv":
                conv_bn(p, np.arange(c2), pc2, cols, n_cols)
            elif t == "C2f":
                c, n = L["c"], L["n"]
                pc = up(c)
                conv_bn(p + ".cv1", blocks(2, c, pc), 2 * pc, cols, n_cols)
                for j in range(n):
                    conv_bn(f"{p}.m.{j}.cv1", np.arange(c), pc, np.arange(c), pc)
                    conv_bn(f"{p}.m.{j}.cv2", np.arange(c), pc, np.arange(c), pc)
                conv_bn(p + ".cv2", np.arange(c2), pc2, blocks(2 + n, c, pc), (2 + n) * pc)
                Lp["c"] = pc
            else:
                c_ = L["c1"] // 2
                pc_ = up(c_)
                conv_bn(p + ".cv1", np.arange(c_), pc_, cols, n_cols)
                conv_bn(p + ".cv2", np.arange(c2), pc2, blocks(4, c_, pc_), 4 * pc_)
                if 2 * pc_ != n_cols:
                    raise NotImplementedError("SPPF: padded hidden width must stay half of the padded input")
            omap[i], ophys[i] = np.arange(c2), pc2
            Lp["c1"], Lp["c2"], Lp["c_out"] = n_cols, pc2, pc2
        elif t == "Upsample":
            s_in = src(i, f)
            omap[i], ophys[i] = omap[s_in], ophys[s_in]
            Lp["c_out"] = ophys[i]
        elif t == "Concat":
            idx, off = [], 0
            for x in f:
                s_in = src(i, x)
                idx.append(off + omap[s_in]); off += ophys[s_in]
            omap[i], ophys[i] = np.concatenate(idx), off
            Lp["c_out"] = off
        elif t == "Detect":
            cb, cc, nc = L["c2_box"], L["c3_cls"], L["nc"]
            pcb, pcc = up(cb), up(cc)
            for l, x in enumerate(f):
                cols, n_cols = omap[x], ophys[x]
                for br, ch, pch, n_out in (("cv2", cb, pcb, 4 * REG_MAX), ("cv3", cc, pcc, nc)):
                    conv_bn(f"{p}.{br}.{l}.0", np.arange(ch), pch, cols, n_cols)
                    conv_bn(f"{p}.{br}.{l}.1", np.arange(ch), pch, np.arange(ch), pch)
                    w = np.asarray(sd[f"{p}.{br}.{l}.2.weight"], np.float32)
                    wp = np.zeros((n_out, pch, 1, 1), np.float32)
                    wp[:, :ch] = w
                    out[f"{p}.{br}.{l}.2.weight"] = wp
                    out[f"{p}.{br}.{l}.2.bias"] = np.asarray(sd[f"{p}.{br}.{l}.2.bias"], np.float32)
            out[f"{p}.dfl.conv.weight"] = np.asarray(sd[f"{p}.dfl.conv.weight"], np.float32)
            Lp["ch"], Lp["c2_box"], Lp["c3_cls"] = [ophys[x] for x in f], pcb, pcc
        else:
            raise NotImplementedError(t)
    return sp, out


# ---------------------------------------------------------------------------------------------------
# `.pt` checkpoints (ultralytics/nn/tasks.py:1404-1521 torch_safe_load / load_checkpoint) without the ultralytics package
# ---------------------------------------------------------------------------------------------------
class _CkptStub:
    """Stands in for any class of the `ultralytics` package while unpickling: keeps the pickled attribute dict.  nn.Module
    subclasses arrive as their `__dict__` (`_parameters`, `_buffers`, `_modules`, plus plain attributes such as `yaml`, `names`)."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, tuple) and len(state) == 2:            # (dict, slots)
            state = {**(state[0] or {}), **(state[1] or {})}
        if isinstance(state, dict):
            self.__dict__.update(state)


def _walk_state(obj, prefix, out):
    """state_dict() of a pickled module tree whose ultralytics containers are stubs and whose leaves are real torch modules."""
    for name, p in (getattr(obj, "_parameters", None) or {}).items():
        if p is not None:
            out[prefix + name] = p.detach()
    persistent_off = getattr(obj, "_non_persistent_buffers_set", None) or set()
    for name, b in (getattr(obj, "_buffers", None) or {}).items():
        if b is not None and name not in persistent_off:
            out[prefix + name] = b
    for name, child in (getattr(obj, "_modules", None) or {}).items():
        if child is not None:
            _walk_state(child, prefix + name + ".", out)


def load_checkpoint(path):
    """Read an Ultralytics `.pt` checkpoint (``model.save(...)`` / the trainer's ``best.pt``) -> ``(model_dict, state_dict, names)``:
    the model YAML dict the network was built from (``DetectionModel.yaml``: scale, nc, backbone, head -- what :func:`cfg.resolve`
    takes), the float32 numpy state_dict with the reference's key names (``model.0.conv.weight`` ...) and the class-name dict.
    Follows load_checkpoint (tasks.py:1487-1521): the EMA weights win over ``model``, half weights are widened to fp32.  The
    ``ultralytics`` package is NOT needed: its classes are unpickled as attribute-dict stubs, tensors by torch itself."""
    import pickle
    import types

    import torch

    stubs = {}

    class _Unpickler(pickle.Unpickler):
        def find_class(self, mod, name):
            root = mod.split(".")[0]
            if root in ("ultralytics", "__main__", "models", "utils"):      # the last two: how YOLOv5-era pickles name their classes
                if root in ("models", "utils"):
                    raise TypeError(f"{path} looks like a YOLOv5 checkpoint ({mod}.{name}); it is not forwards compatible (tasks.py:1450-1459)")
                key = (mod, name)
                if key not in stubs:
                    stubs[key] = type(name, (_CkptStub,), {"__module__": "b200dt._ckpt." + mod})
                return stubs[key]
            return super().find_class(mod, name)

    pm = types.ModuleType("b2dt_ckpt_pickle")
    pm.Unpickler = _Unpickler
    pm.load = lambda f, **kw: _Unpickler(f, **kw).load()
    pm.__name__ = "pickle"
    with open(path, "rb") as fh:
        ckpt = torch.load(fh, map_location="cpu", pickle_module=pm, weights_only=False)
    if not isinstance(ckpt, dict):                                          # torch.save(model, ...) of a whole YOLO object (tasks.py:1475-1481)
        ckpt = {"model": getattr(ckpt, "model", ckpt)}
    model = ckpt.get("ema") or ckpt["model"]
    yaml_d = getattr(model, "yaml", None)
    if not isinstance(yaml_d, dict) or "backbone" not in yaml_d or "head" not in yaml_d:
        raise ValueError(f"{path}: the pickled model carries no YAML dict (model.yaml); cannot rebuild the layer list")
    sd = {}
    _walk_state(model, "", sd)
    if not sd:
        raise ValueError(f"{path}: no parameters found in the pickled model")
    names = getattr(model, "names", None)
    if isinstance(names, (list, tuple)):
        names = dict(enumerate(names))
    out = {k: v.float().numpy() for k, v in sd.items()}
    return dict(yaml_d), out, names
