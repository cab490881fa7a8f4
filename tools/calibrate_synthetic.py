"""One-off calibration of the seeded synthetic weights (runs only where /root/reference exists).

Random-init weights are degenerate for inference.  This tool loads the numpy recipe
(`weights.synthetic_state_dict(calib=None)`) into the REFERENCE DetectionModel, and in one eval-mode
forward over seeded synthetic IR frames sets, layer by layer in execution order,
  * every BatchNorm's running_mean = 0 and running_var = E[y^2] of its own input (variance-only
    calibration, better conditioned under bf16 than mean subtraction), and
  * the gain and bias of the final 1x1 convs of each Detect branch so that box logits have std 0.5
    around the recipe's bias ramp and a chosen fraction of anchors per level clears conf=0.15.
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
FRACTIONS = (0.003, 0.01, 0.02, 0.04)   # anchors per level allowed above CONF (P2..P5)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("model")
    ap.add_argument("--nc", type=int, default=None)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--imgsz", type=int, nargs=2, default=(512, 640))
    ap.add_argument("--batch", type=int, default=2)
    a = ap.parse_args()
    from ultralytics.nn.tasks import DetectionModel

    spec = cfg.resolve(a.model, nc=a.nc)
    sd = weights.synthetic_state_dict(spec, seed=a.seed, calib=None)
    m = DetectionModel(a.model + ".yaml", ch=3, nc=spec["nc"], verbose=False).eval()
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    H, W = a.imgsz
    frames = np.stack([synth.IRStream(seed=100 + i, h=H, w=W).frame() for i in range(a.batch)])
    x = torch.from_numpy(np.ascontiguousarray(frames[..., ::-1].transpose(0, 3, 1, 2))).float() / 255

    def bn_hook(mod, inp):
        y = inp[0]
        mod.running_mean.zero_()
        mod.running_var.copy_((y * y).mean((0, 2, 3)))

    hooks = [mod.register_forward_pre_hook(bn_hook) for mod in m.modules() if isinstance(mod, torch.nn.BatchNorm2d)]
    det = m.model[-1]
    logit_thr = float(np.log(CONF / (1 - CONF)))

    def head_hook(kind, level):
        def fn(mod, inp, out):
            noise = out - mod.bias.view(1, -1, 1, 1)
            if kind == "box":
                g = 0.5 / float(noise.std())
                mod.weight.mul_(g)
                return noise * g + mod.bias.view(1, -1, 1, 1)
            g = 1.0 / float(noise.std())
            mod.weight.mul_(g)
            noise = noise * g
            mx = noise.amax(1).flatten()
            q = torch.quantile(mx, 1 - FRACTIONS[level])
            mod.bias.fill_(logit_thr - float(q))
            return noise + mod.bias.view(1, -1, 1, 1)
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
        if k.endswith("running_var") or k.endswith("running_mean") or ".2.weight" in k or ".2.bias" in k:
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
