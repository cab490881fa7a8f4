"""A few full detect+track steps of the bench workload (for `ncu --metrics gpu__time_duration.sum`) followed by one C3-size
tracker frame (256 streams x 4096 tracks x 40 detections).  Prints nothing but the emitted-track count."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import b200dt  # noqa: F401
from b200dt import synth
from b200dt.pipeline import DetectTrackPipeline
from b200dt.tracker import TrackerBank


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=256)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--c3", type=int, default=1)
    a = ap.parse_args()
    H, W = 512, 640
    vids = [synth.IRStream(seed=1000 + s, h=H, w=W) for s in range(min(a.streams, 16))]
    pipe = DetectTrackPipeline("yolov8s-p2", a.streams, (H, W), 640, 0.15, 0.6, 300, capacity=2048, max_lost_frames=150, min_hits=1, iou_threshold=0.1)
    for t in range(a.steps):
        fr = [v.frame() for v in vids]
        frames = torch.from_numpy(np.stack([fr[s % len(fr)] for s in range(a.streams)])).cuda()
        rows, counts = pipe.step_device(frames)
    torch.cuda.synchronize()
    print("mean emitted tracks per stream:", float(counts.float().mean()))
    if a.c3:
        S, C, D = 256, 4096, 1024
        bank = TrackerBank(S, C, D, 150, 1, 0.1)
        for r in range(C // D):
            idx = torch.arange(D, device="cuda") + r * D
            x = (idx % 64).float() * 10.0
            y = (idx // 64).float() * 10.0
            boxes = torch.stack([x, y, x + 6, y + 6], 1)[None].repeat(S, 1, 1).contiguous()
            bank.update(boxes, torch.full((S,), D, dtype=torch.int32, device="cuda"), with_trajectory=False)
        dets = torch.zeros((S, D, 6), device="cuda")
        g = torch.Generator(device="cuda").manual_seed(0)
        pick = torch.randint(0, C, (S, 40), device="cuda", generator=g)
        px, py = (pick % 64).float() * 10.0, (pick // 64).float() * 10.0
        dets[:, :40, 0], dets[:, :40, 1], dets[:, :40, 2], dets[:, :40, 3], dets[:, :40, 4] = px, py, px + 6, py + 6, 0.9
        cnt = torch.full((S,), 40, dtype=torch.int32, device="cuda")
        torch.cuda.profiler.start()
        for _ in range(2):
            bank.update(dets, cnt, with_trajectory=False)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
