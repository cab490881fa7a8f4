"""Oracle (numpy, CPU) for the YOLOv8-P2 forward -- test infrastructure, see oracle/__init__.py.

Restates, from the reference sources, the graph that
``DetectionModel('yolov8{n,s,m,l,x}-p2.yaml')`` / ``yolov8-small.yaml`` builds and the
arithmetic its fused (BN-folded) inference path performs:

  graph construction   ultralytics/nn/tasks.py:1524-1700 (parse_model: depth/width scaling,
                       make_divisible(min(c2,max_ch)*width, 8), repeats -> C2f n)
  topologies           ultralytics/cfg/models/v8/yolov8-p2.yaml:9-57, yolov8-small.yaml:12-60
  Conv (+BN fold)      ultralytics/nn/modules/conv.py:30-93, ultralytics/utils/torch_utils.py:255-286
  Bottleneck / C2f     ultralytics/nn/modules/block.py:470-495, :294-326
  SPPF                 ultralytics/nn/modules/block.py:216-241
  Upsample / Concat    yolov8-p2.yaml:33-54, ultralytics/nn/modules/conv.py:673-683
  Detect (legacy head) ultralytics/nn/modules/head.py:80-126
  layer routing        ultralytics/nn/tasks.py:159-188 (_predict_once)

Third-party arithmetic restated here: torch.nn.functional.conv2d / BatchNorm2d(eval) / SiLU /
MaxPool2d / Upsample(nearest) (PyTorch, pinned torch>=1.8 in the reference's pyproject.toml:71;
2.11.0 in this image) -- plain cross-correlation with zero padding, fp32.

Two numeric modes:
  "fp32"  the reference's arithmetic (pinned against the reference run in-process,
          tests/golden/net_*.npz).
  "bf16"  the same graph with the rounding points of the B200 engine: BN-folded weights rounded
          to bf16, the stem input taken as bf16(255 x) (exact uint8 pixel values), fp32
          accumulation, every stored activation rounded to bf16.
"""
from __future__ import annotations

import math

import numpy as np

# --------------------------------------------------------------------------------------
# topologies (ultralytics/cfg/models/v8/yolov8-p2.yaml:9-57, yolov8-small.yaml:12-60)
# each row: (from, repeats, module, args)
# --------------------------------------------------------------------------------------
_P2_LAYERS = [
    (-1, 1, "Conv", (64, 3, 2)), (-1, 1, "Conv", (128, 3, 2)), (-1, 3, "C2f", (128, True)),
    (-1, 1, "Conv", (256, 3, 2)), (-1, 6, "C2f", (256, True)), (-1, 1, "Conv", (512, 3, 2)),
    (-1, 6, "C2f", (512, True)), (-1, 1, "Conv", (1024, 3, 2)), (-1, 3, "C2f", (1024, True)),
    (-1, 1, "SPPF", (1024, 5)),
    (-1, 1, "Upsample", ()), ((-1, 6), 1, "Concat", ()), (-1, 3, "C2f", (512,)),
    (-1, 1, "Upsample", ()), ((-1, 4), 1, "Concat", ()), (-1, 3, "C2f", (256,)),
    (-1, 1, "Upsample", ()), ((-1, 2), 1, "Concat", ()), (-1, 3, "C2f", (128,)),
    (-1, 1, "Conv", (128, 3, 2)), ((-1, 15), 1, "Concat", ()), (-1, 3, "C2f", (256,)),
    (-1, 1, "Conv", (256, 3, 2)), ((-1, 12), 1, "Concat", ()), (-1, 3, "C2f", (512,)),
    (-1, 1, "Conv", (512, 3, 2)), ((-1, 9), 1, "Concat", ()), (-1, 3, "C2f", (1024,)),
    ((18, 21, 24, 27), 1, "Detect", ()),
]
_P2_SCALES = {"n": (0.33, 0.25, 1024), "s": (0.33, 0.50, 1024), "m": (0.67, 0.75, 768),
              "l": (1.00, 1.00, 512), "x": (1.00, 1.25, 512)}

_SMALL_LAYERS = [
    (-1, 1, "Conv", (32, 3, 2)), (-1, 1, "Conv", (64, 3, 2)), (-1, 3, "C2f", (64, True)),
    (-1, 1, "Conv", (128, 3, 2)), (-1, 6, "C2f", (128, True)), (-1, 1, "Conv", (256, 3, 2)),
    (-1, 6, "C2f", (256, True)), (-1, 1, "Conv", (512, 3, 2)), (-1, 3, "C2f", (512, True)),
    (-1, 1, "SPPF", (512, 5)),
    (-1, 1, "Upsample", ()), ((-1, 6), 1, "Concat", ()), (-1, 3, "C2f", (256,)),
    (-1, 1, "Upsample", ()), ((-1, 4), 1, "Concat", ()), (-1, 3, "C2f", (128,)),
    (-1, 1, "Upsample", ()), ((-1, 2), 1, "Concat", ()), (-1, 3, "C2f", (64,)),
    (15, 1, "Conv", (128, 3, 2)), ((-1, 12), 1, "Concat", ()), (-1, 3, "C2f", (256,)),
    (-1, 1, "Conv", (256, 3, 2)), ((-1, 9), 1, "Concat", ()), (-1, 3, "C2f", (512,)),
    ((18, 15, 21, 24), 1, "Detect", ()),
]
_SMALL_SCALES = {"n": (0.50, 0.375, 1024), "s": (0.67, 0.625, 1024), "m": (1.00, 0.875, 768),
                 "l": (1.33, 1.125, 512), "x": (1.67, 1.375, 512)}

BN_EPS = 1e-3  # ultralytics/utils/torch_utils.py:488-498 (initialize_weights)
REG_MAX = 16   # ultralytics/nn/modules/head.py:91


def make_divisible(x, d):
    return int(math.ceil(x / d) * d)


def build_spec(name="yolov8n-p2", nc=None):
    """Resolve a model name to a list of layer dicts with concrete channel counts.

    Follows parse_model (nn/tasks.py:1615-1700) and yaml_model_load's scale-letter handling
    (:1703-1724): 'yolov8s-p2' -> family p2, scale 's'; 'yolov8-small' has no scale letter so the
    first scale ('n') is taken (nn/tasks.py:1545-1549).
    """
    base = name.replace(".yaml", "")
    if base.endswith("-p2"):
        scale = base[len("yolov8"):-len("-p2")] or "n"
        rows, scales, nc_default = _P2_LAYERS, _P2_SCALES, 80
    elif base.startswith("yolov8") and base.endswith("-small"):
        scale = base[len("yolov8"):-len("-small")] or "n"
        rows, scales, nc_default = _SMALL_LAYERS, _SMALL_SCALES, 1
    else:
        raise ValueError(f"unknown model {name!r}")
    nc = nc_default if nc is None else nc
    depth, width, max_ch = scales[scale]
    ch, layers = [3], []
    for i, (f, n, m, args) in enumerate(rows):
        n = max(round(n * depth), 1) if n > 1 else n
        L = {"i": i, "f": f, "type": m}
        if m in ("Conv", "C2f", "SPPF"):
            c1 = ch[f]
            c2 = make_divisible(min(args[0], max_ch) * width, 8)
            L.update(c1=c1, c2=c2)
            if m == "Conv":
                L.update(k=args[1], s=args[2])
            elif m == "C2f":
                L.update(n=n, shortcut=bool(args[1]) if len(args) > 1 else False, c=int(c2 * 0.5))
            else:
                L.update(k=args[1])
        elif m == "Upsample":
            c2 = ch[f]
        elif m == "Concat":
            c2 = sum(ch[x] for x in f)
        elif m == "Detect":
            chs = [ch[x] for x in f]
            L.update(ch=chs, nc=nc, c2_box=max(16, chs[0] // 4, REG_MAX * 4),
                     c3_cls=max(chs[0], min(nc, 100)))
            c2 = None
        L["c_out"] = c2
        layers.append(L)
        if i == 0:
            ch = []
        ch.append(c2)
    return {"name": base, "scale": scale, "nc": nc, "layers": layers,
            "strides": [4, 8, 16, 32]}


# --------------------------------------------------------------------------------------
# numerics
# --------------------------------------------------------------------------------------
_CONV_BACKEND = "numpy"


def bf16_round(a):
    """Round-to-nearest-even fp32 -> bf16, returned as fp32."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32)
    r = ((u >> np.uint32(16)) & np.uint32(1)) + np.uint32(0x7FFF)
    return ((u + r) & np.uint32(0xFFFF0000)).view(np.float32)


def silu(x):
    if _CONV_BACKEND == "aten":
        import torch

        return torch.nn.functional.silu(torch.from_numpy(np.ascontiguousarray(x, np.float32))).numpy()
    with np.errstate(over="ignore"):
        return (x / (1.0 + np.exp(-x, dtype=np.float32))).astype(np.float32)


def set_conv_backend(name):
    """"numpy" (default: explicit tap loop + matmul) or "aten": the same cross-correlation evaluated by the
    third-party primitive the reference itself calls on CPU (torch.nn.functional.conv2d -> ATen/oneDNN,
    ultralytics/nn/modules/conv.py:83-93).  "aten" is used only for the timed CPU-baseline leg of bench.py, so
    that the baseline runs at the reference's real CPU speed; tests check both backends agree."""
    global _CONV_BACKEND
    assert name in ("numpy", "aten")
    _CONV_BACKEND = name


def conv2d(x, w, b, s=1, p=0):
    """Cross-correlation, NCHW, fp32 (torch.nn.functional.conv2d semantics, groups=1)."""
    if _CONV_BACKEND == "aten":
        import torch

        with torch.no_grad():
            y = torch.nn.functional.conv2d(torch.from_numpy(np.ascontiguousarray(x, np.float32)), torch.from_numpy(np.ascontiguousarray(w, np.float32)),
                                           None if b is None else torch.from_numpy(np.ascontiguousarray(b, np.float32)), stride=s, padding=p)
        return y.numpy()
    B, C, H, W = x.shape
    O, C2, k, _ = w.shape
    assert C == C2, (x.shape, w.shape)
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    xp = np.pad(x, ((0, 0), (0, 0), (p, p), (p, p))) if p else x
    out = np.zeros((B, O, Ho * Wo), np.float32)
    for kh in range(k):
        for kw in range(k):
            patch = xp[:, :, kh:kh + s * (Ho - 1) + 1:s, kw:kw + s * (Wo - 1) + 1:s]
            out += np.matmul(w[:, :, kh, kw], np.ascontiguousarray(patch).reshape(B, C, Ho * Wo))
    if b is not None:
        out += b[None, :, None]
    return out.reshape(B, O, Ho, Wo)


def maxpool5(x):
    """MaxPool2d(5, stride 1, pad 2) with -inf padding (torch semantics)."""
    B, C, H, W = x.shape
    xp = np.full((B, C, H + 4, W + 4), -np.inf, np.float32)
    xp[:, :, 2:-2, 2:-2] = x
    out = xp[:, :, 0:H, 0:W].copy()
    for dh in range(5):
        for dw in range(5):
            np.maximum(out, xp[:, :, dh:dh + H, dw:dw + W], out=out)
    return out


def upsample2(x):
    return x.repeat(2, axis=2).repeat(2, axis=3)


def fold_conv_bn(sd, prefix):
    """fuse_conv_and_bn (utils/torch_utils.py:255-286): w' = w*g/sqrt(var+eps), b' = beta - g*mu/sqrt(var+eps)."""
    w = np.asarray(sd[prefix + ".conv.weight"], np.float32)
    g = np.asarray(sd[prefix + ".bn.weight"], np.float32)
    beta = np.asarray(sd[prefix + ".bn.bias"], np.float32)
    mu = np.asarray(sd[prefix + ".bn.running_mean"], np.float32)
    var = np.asarray(sd[prefix + ".bn.running_var"], np.float32)
    scale = g / np.sqrt(var + np.float32(BN_EPS))
    return (w * scale[:, None, None, None]).astype(np.float32), (beta - mu * scale).astype(np.float32)


class Net:
    """The fused inference graph of DetectionModel, evaluated with numpy."""

    def __init__(self, spec, state_dict, mode="fp32"):
        assert mode in ("fp32", "bf16")
        self.spec, self.sd, self.mode = spec, state_dict, mode
        self.trace = {}      # module path -> activation (filled when record=True)
        self._cache = {}

    # -- helpers ------------------------------------------------------------------
    def _q(self, a):
        return bf16_round(a) if self.mode == "bf16" else a

    def _cba(self, x, prefix, k, s, stem=False, record=False):
        """Conv.forward_fuse (conv.py:83-93): SiLU(conv(x)+b) with BN folded."""
        if prefix not in self._cache:
            w, b = fold_conv_bn(self.sd, prefix)
            if self.mode == "bf16":
                w = bf16_round(w)
            self._cache[prefix] = (w, b)
        w, b = self._cache[prefix]
        y = self._q(silu(conv2d(x, w, b, s, k // 2)))
        if record:
            self.trace[prefix] = y
        return y

    def _plain(self, x, prefix, record=False):
        """Final nn.Conv2d(c, out, 1) with bias of each Detect branch (head.py:93-96), no activation."""
        w = np.asarray(self.sd[prefix + ".weight"], np.float32)
        b = np.asarray(self.sd[prefix + ".bias"], np.float32)
        if self.mode == "bf16":
            w = bf16_round(w)
        y = self._q(conv2d(x, w, b, 1, 0))
        if record:
            self.trace[prefix] = y
        return y

    # -- modules ------------------------------------------------------------------
    def _c2f(self, x, L, record):
        p = f"model.{L['i']}"
        c = L["c"]
        y0 = self._cba(x, p + ".cv1", 1, 1, record=record)
        ys = [y0[:, :c], y0[:, c:]]
        for j in range(L["n"]):
            t = self._cba(ys[-1], f"{p}.m.{j}.cv1", 3, 1, record=record)
            if L["shortcut"]:
                # x + cv2(cv1(x)) (block.py:493-495); engine adds in fp32 before the single bf16 store
                w, b = self._get(f"{p}.m.{j}.cv2")
                t2 = self._q(ys[-1] + silu(conv2d(t, w, b, 1, 1)))
                if record:
                    self.trace[f"{p}.m.{j}.cv2"] = t2
            else:
                t2 = self._cba(t, f"{p}.m.{j}.cv2", 3, 1, record=record)
            ys.append(t2)
        return self._cba(np.concatenate(ys, 1), p + ".cv2", 1, 1, record=record)

    def _get(self, prefix):
        if prefix not in self._cache:
            w, b = fold_conv_bn(self.sd, prefix)
            if self.mode == "bf16":
                w = bf16_round(w)
            self._cache[prefix] = (w, b)
        return self._cache[prefix]

    def _sppf(self, x, L, record):
        p = f"model.{L['i']}"
        y = [self._cba(x, p + ".cv1", 1, 1, record=record)]
        for _ in range(3):
            y.append(maxpool5(y[-1]))
        return self._cba(np.concatenate(y, 1), p + ".cv2", 1, 1, record=record)

    def _detect(self, xs, L, record):
        """Detect.forward up to the per-level cat (head.py:116-121): list of (B, 64+nc, H, W)."""
        p = f"model.{L['i']}"
        outs = []
        for l, x in enumerate(xs):
            a = self._cba(x, f"{p}.cv2.{l}.0", 3, 1, record=record)
            a = self._cba(a, f"{p}.cv2.{l}.1", 3, 1, record=record)
            a = self._plain(a, f"{p}.cv2.{l}.2", record=record)
            c = self._cba(x, f"{p}.cv3.{l}.0", 3, 1, record=record)
            c = self._cba(c, f"{p}.cv3.{l}.1", 3, 1, record=record)
            c = self._plain(c, f"{p}.cv3.{l}.2", record=record)
            outs.append(np.concatenate([a, c], 1))
        return outs

    # -- graph --------------------------------------------------------------------
    def forward(self, x, record=False):
        """x: (B,3,H,W) float32 in [0,1] (RGB).  Returns the per-level raw head maps."""
        x = np.asarray(x, np.float32)
        if self.mode == "bf16":
            # the engine's stem consumes bf16(255 x) (exact for uint8-derived inputs) and scales by 1/255 in fp32
            x = (bf16_round(x * np.float32(255.0)) / np.float32(255.0)).astype(np.float32)
        ys = []
        for L in self.spec["layers"]:
            f = L["f"]
            if isinstance(f, tuple):
                inp = [x if j == -1 else ys[j] for j in f]
            else:
                inp = x if f == -1 else ys[f]
            t = L["type"]
            if t == "Conv":
                x = self._cba(inp, f"model.{L['i']}", L["k"], L["s"], stem=(L["i"] == 0), record=record)
            elif t == "C2f":
                x = self._c2f(inp, L, record)
            elif t == "SPPF":
                x = self._sppf(inp, L, record)
            elif t == "Upsample":
                x = upsample2(inp)
            elif t == "Concat":
                x = np.concatenate(inp, 1)
            elif t == "Detect":
                x = self._detect(inp, L, record)
            if record and t in ("C2f", "SPPF", "Conv"):
                self.trace[f"layer.{L['i']}"] = x
            ys.append(x)
        return x


def conv_flops(spec, H, W):
    """Algorithmic conv FLOPs per image, SURVEY.md section 8(d): sum 2*Ho*Wo*Cout*Cin*k*k on the fused graph."""
    total = 0
    hw = {}
    cur = (H, W)
    for L in spec["layers"]:
        f = L["f"]
        src = cur if (f == -1) else hw[f if isinstance(f, int) else f[0]] if not isinstance(f, tuple) else None
        if isinstance(f, tuple):
            src = cur if f[0] == -1 else hw[f[0]]
        t = L["type"]
        h, w = src
        if t == "Conv":
            h, w = (h + 2 * (L["k"] // 2) - L["k"]) // L["s"] + 1, (w + 2 * (L["k"] // 2) - L["k"]) // L["s"] + 1
            total += 2 * h * w * L["c2"] * L["c1"] * L["k"] ** 2
        elif t == "C2f":
            c = L["c"]
            total += 2 * h * w * (L["c1"] * 2 * c + (2 + L["n"]) * c * L["c2"] + L["n"] * 2 * 9 * c * c)
        elif t == "SPPF":
            c_ = L["c1"] // 2
            total += 2 * h * w * (L["c1"] * c_ + 4 * c_ * L["c2"])
        elif t == "Upsample":
            h, w = 2 * h, 2 * w
        elif t == "Detect":
            for l, fi in enumerate(L["f"]):
                hh, ww = hw[fi]
                cb, cc, ci = L["c2_box"], L["c3_cls"], L["ch"][l]
                total += 2 * hh * ww * (9 * ci * cb + 9 * cb * cb + cb * 4 * REG_MAX
                                         + 9 * ci * cc + 9 * cc * cc + cc * L["nc"])
        hw[L["i"]] = (h, w)
        cur = (h, w)
    return total
