"""One-off calibration of the seeded synthetic weights (runs only where /root/reference exists).

Random-init weights are degenerate for inference.  This tool loads the numpy recipe
(`weights.synthetic_state_dict(calib=None)`) into the REFERENCE DetectionModel, and in one eval-mode
forward over seeded synthetic IR frames sets, layer by layer in execution order,
  * every BatchNorm's running_mean = 0 and running_var = E[y^2] of its own input (variance-only
    calibration, better conditioned under bf16 than mean subtraction) -- except the blob-highway
    channels (weights.highway_channels), whose statistics stay at identity,
  * the background bias of the highway's second stage: -(median + 5.5 * 1.4826 * MAD) of its
    pre-activation, so that only blobs come out positive,
  * the gain of the final 1x1 convs of each Detect branch: box logits get std BOX_STD around the recipe's
    DFL bias, the random class logits std CLS_STD around a bias of -9 (they never reach conf=0.15; class 0
    of the P2 level is the highway's output and is left alone).
Only those vectors are written to <package>/calib/<model>_nc<nc>_seed<seed>.npz; conv weights stay
regenerable from the seed.  Usage: python tools/calibrate_synthetic.py yolov8n-p2 [--imgsz H W]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
os.environ.setdefault("YOLO_CONFIG_DIR", "/tmp/ycfg")
os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count()))

import numpy as np
import torch

import b200dt  # noqa: F401
from b200dt import cfg, synth, weights

CONF = 0.15
BOX_STD, CLS_STD = 0.03, 0.1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("model")
    ap.add_argument("--nc", type=int, default=None)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--imgsz", type=int, nargs=2, default=(512, 640))
    ap.add_argument("--batch", type=int, default=4)
    a = ap.parse_args()
    from ultralytics.nn.tasks import DetectionModel

    spec = cfg.resolve(a.model, nc=a.nc)
    sd = weights.synthetic_state_dict(spec, seed=a.seed, calib=None, bake=False)
    m = DetectionModel(a.model + ".yaml", ch=3, nc=spec["nc"], verbose=False).eval()
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    H, W = a.imgsz
    frames = np.stack([synth.IRStream(seed=100 + i, h=H, w=W).frame() for i in range(a.batch)])
    x = torch.from_numpy(np.ascontiguousarray(frames[..., ::-1].transpose(0, 3, 1, 2))).float() / 255

    protected = {p + ".bn": ch for p, ch in weights.highway_channels(spec).items()}
    stage1 = weights.highway_path(spec)[1][0] + ".bn"
    names = {id(mod): n for n, mod in m.named_modules()}

    def bn_hook(mod, inp):
        y = inp[0]
        name = names[id(mod)]
        mod.running_mean.zero_()
        var = (y * y).mean((0, 2, 3))
        for c in protected.get(name, []):
            var[c] = 1.0 - mod.eps
        mod.running_var.copy_(var)
        if name == stage1:
            h = y[:, 0].flatten()
            med = h.median()
            mad = (h - med).abs().median()
            mod.bias[0] = -float(med + weights.HW_BG_SIGMAS * 1.4826 * mad)

    hooks = [mod.register_forward_pre_hook(bn_hook) for mod in m.modules() if isinstance(mod, torch.nn.BatchNorm2d)]
    det = m.model[-1]

    def head_hook(kind, level):
        def fn(mod, inp, out):
            noise = out - mod.bias.view(1, -1, 1, 1)
            if kind == "box":
                g = BOX_STD / float(noise.std())
                mod.weight.mul_(g)
                return noise * g + mod.bias.view(1, -1, 1, 1)
            rows = slice(1, None) if level == 0 else slice(None)
            if noise[:, rows].numel():
                g = CLS_STD / float(noise[:, rows].std())
                mod.weight[rows] *= g
                out = out.clone()
                out[:, rows] = noise[:, rows] * g + mod.bias.view(1, -1, 1, 1)[:, rows]
            return out
        return fn

    for l in range(det.nl):
        hooks.append(det.cv2[l][2].register_forward_hook(head_hook("box", l)))
        hooks.append(det.cv3[l][2].register_forward_hook(head_hook("cls", l)))
    with torch.no_grad():
        m(x)
    for h in hooks:
        h.remove()
    out = {}
    for k, v in m.state_dict().items():
        if k.endswith("running_var") or k.endswith("running_mean") or ".2.weight" in k or ".2.bias" in k or k == stage1 + ".bias":
            out[k] = v.numpy().astype(np.float32)
    path = weights.calib_path(spec, a.seed)
    np.savez_compressed(path, **out)
    with torch.no_grad():
        y, _ = m(x)
    sc = y[:, 4:].amax(1)
    print(f"wrote {path} ({os.path.getsize(path)/1024:.0f} KiB); candidates > {CONF}: {(sc > CONF).sum(1).tolist()}  "
          f"max score {float(sc.max()):.3f}; wh median {float(y[:, 2:4].median()):.1f}")


if __name__ == "__main__":
    main()
