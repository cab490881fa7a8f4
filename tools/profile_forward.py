"""One detect step (stem + forward + candidates + NMS) of the bench workload at a reduced stream count, for ncu.

    B2_NCU_OPS=3,57,59 ncu --set full --clock-control none --import-source on --profile-from-start off \
        -o gpurun_out/prof python tools/profile_forward.py --streams 64
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import b200dt
from b200dt import _lib, cfg, synth, weights
from b200dt.predictor import DetectPipeline


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=64)
    ap.add_argument("--model", default="yolov8s-p2")
    ap.add_argument("--passes", type=int, default=2)
    a = ap.parse_args()
    spec = cfg.resolve(a.model, nc=80)
    sd = weights.synthetic_state_dict(spec, seed=0)
    H, W = 512, 640
    vids = [synth.IRStream(seed=1000 + s, h=H, w=W) for s in range(min(a.streams, 8))]
    fr = [v.frame() for v in vids]
    frames = torch.from_numpy(np.stack([fr[s % len(fr)] for s in range(a.streams)])).cuda()
    pipe = DetectPipeline(spec, sd, a.streams, H, W)
    lib = _lib.load()
    for i in range(a.passes):
        dets, cnt = pipe(frames, 0.15, 0.6, orig_hw=(H, W))
        torch.cuda.synchronize()
    # eager pass: the launches named in B2_NCU_OPS (indices into the plan, 0 = stem) are bracketed by
    # cudaProfilerStart/Stop inside b2_engine_profile_u8, so `ncu --profile-from-start off` captures only those
    rows = pipe.engine.profile_u8(frames)
    sel = [int(v) for v in os.environ.get("B2_NCU_OPS", "").split(",") if v.strip()]
    for i in sel:
        print(i, rows[i])
    print("detections per image (first 8):", cnt[:8].tolist())


if __name__ == "__main__":
    main()
