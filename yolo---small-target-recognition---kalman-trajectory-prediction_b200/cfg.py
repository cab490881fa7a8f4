"""Model graph description for the YOLOv8-P2 detect path.

Host-side mirror of the reference's YAML -> layer list resolution
(``yaml_model_load`` ultralytics/nn/tasks.py:1703-1724 and ``parse_model`` :1524-1700): the same
``[from, repeats, module, args]`` rows, the same depth/width/max_channels scaling, restricted to the
modules that appear on the hot path (Conv, C2f, SPPF, nn.Upsample, Concat, Detect).  The resolved
layer list is what :mod:`engine` lowers to a launch plan for the C-ABI library.

Built-in families (reference files they correspond to):
  ``yolov8{n,s,m,l,x}-p2``   ultralytics/cfg/models/v8/yolov8-p2.yaml
  ``yolov8[nsmlx]-small``    ultralytics/cfg/models/v8/yolov8-small.yaml (the project's own variant)
A path to an Ultralytics-style ``*.yaml`` is accepted too (``load_yaml``).
"""
from __future__ import annotations

import math
import os
import re

REG_MAX = 16      # head.py:91
BN_EPS = 1e-3     # torch_utils.py:488-498

_FAMILIES = {
    "p2": {
        "nc": 80,
        "scales": {"n": (0.33, 0.25, 1024), "s": (0.33, 0.50, 1024), "m": (0.67, 0.75, 768),
                   "l": (1.00, 1.00, 512), "x": (1.00, 1.25, 512)},
        "rows": [
            [-1, 1, "Conv", [64, 3, 2]], [-1, 1, "Conv", [128, 3, 2]], [-1, 3, "C2f", [128, True]],
            [-1, 1, "Conv", [256, 3, 2]], [-1, 6, "C2f", [256, True]], [-1, 1, "Conv", [512, 3, 2]],
            [-1, 6, "C2f", [512, True]], [-1, 1, "Conv", [1024, 3, 2]], [-1, 3, "C2f", [1024, True]],
            [-1, 1, "SPPF", [1024, 5]],
            [-1, 1, "nn.Upsample", [None, 2, "nearest"]], [[-1, 6], 1, "Concat", [1]], [-1, 3, "C2f", [512]],
            [-1, 1, "nn.Upsample", [None, 2, "nearest"]], [[-1, 4], 1, "Concat", [1]], [-1, 3, "C2f", [256]],
            [-1, 1, "nn.Upsample", [None, 2, "nearest"]], [[-1, 2], 1, "Concat", [1]], [-1, 3, "C2f", [128]],
            [-1, 1, "Conv", [128, 3, 2]], [[-1, 15], 1, "Concat", [1]], [-1, 3, "C2f", [256]],
            [-1, 1, "Conv", [256, 3, 2]], [[-1, 12], 1, "Concat", [1]], [-1, 3, "C2f", [512]],
            [-1, 1, "Conv", [512, 3, 2]], [[-1, 9], 1, "Concat", [1]], [-1, 3, "C2f", [1024]],
            [[18, 21, 24, 27], 1, "Detect", ["nc"]],
        ],
    },
    "small": {
        "nc": 1,
        "scales": {"n": (0.50, 0.375, 1024), "s": (0.67, 0.625, 1024), "m": (1.00, 0.875, 768),
                   "l": (1.33, 1.125, 512), "x": (1.67, 1.375, 512)},
        "rows": [
            [-1, 1, "Conv", [32, 3, 2]], [-1, 1, "Conv", [64, 3, 2]], [-1, 3, "C2f", [64, True]],
            [-1, 1, "Conv", [128, 3, 2]], [-1, 6, "C2f", [128, True]], [-1, 1, "Conv", [256, 3, 2]],
            [-1, 6, "C2f", [256, True]], [-1, 1, "Conv", [512, 3, 2]], [-1, 3, "C2f", [512, True]],
            [-1, 1, "SPPF", [512, 5]],
            [-1, 1, "nn.Upsample", [None, 2, "nearest"]], [[-1, 6], 1, "Concat", [1]], [-1, 3, "C2f", [256]],
            [-1, 1, "nn.Upsample", [None, 2, "nearest"]], [[-1, 4], 1, "Concat", [1]], [-1, 3, "C2f", [128]],
            [-1, 1, "nn.Upsample", [None, 2, "nearest"]], [[-1, 2], 1, "Concat", [1]], [-1, 3, "C2f", [64]],
            [15, 1, "Conv", [128, 3, 2]], [[-1, 12], 1, "Concat", [1]], [-1, 3, "C2f", [256]],
            [-1, 1, "Conv", [256, 3, 2]], [[-1, 9], 1, "Concat", [1]], [-1, 3, "C2f", [512]],
            [[18, 15, 21, 24], 1, "Detect", ["nc"]],
        ],
    },
}


def make_divisible(x, divisor):
    return int(math.ceil(x / divisor) * divisor)


def load_yaml(path):
    """Read an Ultralytics model YAML; the scale letter is taken from the file name (tasks.py:1716-1722)."""
    import yaml

    stem = os.path.basename(path)
    unified = re.sub(r"(\d+)([nslmx])(.+)?$", r"\1\3", stem)
    real = path if os.path.exists(path) else os.path.join(os.path.dirname(path), unified)
    with open(real, "r", encoding="utf-8") as fh:
        d = yaml.safe_load(fh)
    m = re.search(r"yolo(e-)?[v]?\d+([nslmx])", os.path.splitext(stem)[0])
    d["scale"] = m.group(2) if m else ""
    d["yaml_file"] = path
    return d


def model_dict(name):
    """Resolve ``name`` (built-in family name or YAML path) to an Ultralytics-style model dict."""
    if os.path.exists(str(name)):
        return load_yaml(str(name))
    base = os.path.basename(str(name)).replace(".yaml", "")
    m = re.fullmatch(r"yolov8([nsmlx]?)-(p2|small)", base)
    if not m:
        if str(name).endswith(".yaml"):
            return load_yaml(str(name))
        raise FileNotFoundError(f"unknown model {name!r}: expected yolov8[nsmlx]-p2 / yolov8[nsmlx]-small or a YAML path")
    fam = _FAMILIES[m.group(2)]
    return {"nc": fam["nc"], "scales": dict(fam["scales"]), "scale": m.group(1),
            "backbone": [list(r) for r in fam["rows"]], "head": [], "yaml_file": base + ".yaml"}


def _literal(a):
    import ast

    try:
        return ast.literal_eval(a)
    except (ValueError, SyntaxError):
        return a


def resolve(name_or_dict, nc=None, ch=3):
    """parse_model for the hot-path module set: returns the concrete layer list.

    Each layer: ``{i, f, type, c1, c2, ...}``; ``Detect`` carries ``ch`` (input channels per level),
    ``c2_box`` / ``c3_cls`` (hidden widths, head.py:92) and ``nc``.
    """
    d = model_dict(name_or_dict) if not isinstance(name_or_dict, dict) else dict(name_or_dict)
    if nc is not None:
        d["nc"] = nc
    nc = d["nc"]
    depth, width, max_ch = d.get("depth_multiple", 1.0), d.get("width_multiple", 1.0), float("inf")
    scales = d.get("scales")
    scale = d.get("scale", "")
    if scales:
        if not scale:
            scale = tuple(scales.keys())[0]   # tasks.py:1545-1549: no scale letter -> first scale
        depth, width, max_ch = scales[scale]
    chs, layers = [ch], []
    for i, (f, n, m, args) in enumerate(d["backbone"] + d["head"]):
        args = [nc if a == "nc" else a for a in args]
        # parse_model (tasks.py:1580-1584) literal_evals string arguments: a YAML dict pickled inside a checkpoint carries 'None'
        args = ["nearest" if a == "nearest" else (None if a == "None" else (a if not isinstance(a, str) or a in ("nc",) else _literal(a))) for a in args]
        n = max(round(n * depth), 1) if n > 1 else n
        f = tuple(f) if isinstance(f, (list, tuple)) else f
        L = {"i": i, "f": f, "type": m.replace("nn.", "")}
        if m in ("Conv", "C2f", "SPPF"):
            c1, c2 = chs[f], args[0]
            if c2 != nc:
                c2 = make_divisible(min(c2, max_ch) * width, 8)
            L.update(c1=c1, c2=c2)
            if m == "Conv":
                k = args[1] if len(args) > 1 else 1
                s = args[2] if len(args) > 2 else 1
                L.update(k=k, s=s)
            elif m == "C2f":
                L.update(n=n, shortcut=bool(args[1]) if len(args) > 1 else False, c=int(c2 * 0.5))
            else:
                L.update(k=args[1] if len(args) > 1 else 5)
        elif m == "nn.Upsample":
            if list(args[:3]) != [None, 2, "nearest"]:
                raise NotImplementedError(f"Upsample{args}: only nearest x2 is on the hot path")
            c2 = chs[f]
        elif m == "Concat":
            c2 = sum(chs[x] for x in f)
        elif m == "Detect":
            cin = [chs[x] for x in f]
            L.update(ch=cin, nc=nc, c2_box=max(16, cin[0] // 4, REG_MAX * 4), c3_cls=max(cin[0], min(nc, 100)))
            c2 = None
        else:
            raise NotImplementedError(f"module {m!r} is outside the detect hot path (SURVEY.md section 8)")
        L["c_out"] = c2
        layers.append(L)
        if i == 0:
            chs = []
        chs.append(c2)
    det = layers[-1]
    if det["type"] != "Detect":
        raise ValueError("last layer must be Detect")
    return {"name": os.path.basename(str(d.get("yaml_file", "model"))).replace(".yaml", ""),
            "scale": scale, "nc": nc, "layers": layers, "names": {i: f"{i}" for i in range(nc)}}


def conv_list(spec):
    """All conv units in execution order: (state_dict prefix, c1, c2, k, s, has_bn_act)."""
    out = []
    for L in spec["layers"]:
        p = f"model.{L['i']}"
        t = L["type"]
        if t == "Conv":
            out.append((p, L["c1"], L["c2"], L["k"], L["s"], True))
        elif t == "C2f":
            c = L["c"]
            out.append((p + ".cv1", L["c1"], 2 * c, 1, 1, True))
            for j in range(L["n"]):
                out.append((f"{p}.m.{j}.cv1", c, c, 3, 1, True))
                out.append((f"{p}.m.{j}.cv2", c, c, 3, 1, True))
            out.append((p + ".cv2", (2 + L["n"]) * c, L["c2"], 1, 1, True))
        elif t == "SPPF":
            c_ = L["c1"] // 2
            out.append((p + ".cv1", L["c1"], c_, 1, 1, True))
            out.append((p + ".cv2", 4 * c_, L["c2"], 1, 1, True))
        elif t == "Detect":
            for l, ci in enumerate(L["ch"]):
                cb, cc = L["c2_box"], L["c3_cls"]
                out.append((f"{p}.cv2.{l}.0", ci, cb, 3, 1, True))
                out.append((f"{p}.cv2.{l}.1", cb, cb, 3, 1, True))
                out.append((f"{p}.cv2.{l}.2", cb, 4 * REG_MAX, 1, 1, False))
                out.append((f"{p}.cv3.{l}.0", ci, cc, 3, 1, True))
                out.append((f"{p}.cv3.{l}.1", cc, cc, 3, 1, True))
                out.append((f"{p}.cv3.{l}.2", cc, L["nc"], 1, 1, False))
    return out


def level_shapes(spec, H, W):
    """(stride, h, w) of every Detect level for an H x W input (strides from the graph, tasks.py:415-419)."""
    hw, cur = {}, (H, W)
    for L in spec["layers"]:
        f = L["f"]
        src = cur if f == -1 else (hw[f] if isinstance(f, int) else (cur if f[0] == -1 else hw[f[0]]))
        h, w = src
        if L["type"] == "Conv":
            p = L["k"] // 2
            h, w = (h + 2 * p - L["k"]) // L["s"] + 1, (w + 2 * p - L["k"]) // L["s"] + 1
        elif L["type"] == "Upsample":
            h, w = 2 * h, 2 * w
        hw[L["i"]] = (h, w)
        cur = (h, w)
    det = spec["layers"][-1]
    return [(H // hw[fi][0], hw[fi][0], hw[fi][1]) for fi in det["f"]]


def conv_flops(spec, H, W):
    """Algorithmic conv FLOPs per image: sum of 2*Ho*Wo*Cout*Cin*k^2 over the fused graph."""
    hw, cur, total = {}, (H, W), 0
    convs = {c[0]: c for c in conv_list(spec)}
    for L in spec["layers"]:
        f = L["f"]
        src = cur if f == -1 else (hw[f] if isinstance(f, int) else (cur if f[0] == -1 else hw[f[0]]))
        h, w = src
        p = f"model.{L['i']}"
        if L["type"] == "Conv":
            pd = L["k"] // 2
            h, w = (h + 2 * pd - L["k"]) // L["s"] + 1, (w + 2 * pd - L["k"]) // L["s"] + 1
            total += 2 * h * w * L["c1"] * L["c2"] * L["k"] ** 2
        elif L["type"] == "Upsample":
            h, w = 2 * h, 2 * w
        elif L["type"] in ("C2f", "SPPF"):
            for name, c1, c2, k, s, _ in convs.values():
                if name.startswith(p + "."):
                    total += 2 * h * w * c1 * c2 * k * k
        elif L["type"] == "Detect":
            for l, fi in enumerate(L["f"]):
                hh, ww = hw[fi]
                for name, c1, c2, k, s, _ in convs.values():
                    if name.startswith(f"{p}.cv2.{l}.") or name.startswith(f"{p}.cv3.{l}."):
                        total += 2 * hh * ww * c1 * c2 * k * k
        hw[L["i"]] = (h, w)
        cur = (h, w)
    return total
