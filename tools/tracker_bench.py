"""CUDA-event timing of the Kalman track bank: C3 size (256 streams x 4096 tracks x 40 detections per frame) and the bench
pipeline's size (256 streams x capacity 2048, ~300 detections per frame).  usage: python tools/tracker_bench.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import b200dt  # noqa: F401
from b200dt.tracker import TrackerBank


def time_cuda(fn, iters=10):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for s, e in ev:
        s.record(); fn(); e.record()
    torch.cuda.synchronize()
    ts = sorted(s.elapsed_time(e) for s, e in ev)
    return ts[len(ts) // 2]


def grid_boxes(S, n, pitch, size, jitter, g):
    idx = torch.arange(n, device="cuda")
    x = (idx % 64).float() * pitch
    y = (idx // 64).float() * pitch
    b = torch.stack([x, y, x + size, y + size], 1)[None].repeat(S, 1, 1)
    if jitter:
        b = b + torch.randn((S, n, 1), device="cuda", generator=g) * jitter
    return b.contiguous()


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    out = {}
    # C3: every slot live, 40 detections per stream hitting existing tracks
    S, C, D = 256, 4096, 1024
    bank = TrackerBank(S, C, D, 150, 1, 0.1)
    for r in range(C // D):
        boxes = grid_boxes(S, D, 10.0, 6.0, 0.0, g)
        boxes[:, :, 1] += r * 160.0; boxes[:, :, 3] += r * 160.0
        bank.update(boxes, torch.full((S,), D, dtype=torch.int32, device="cuda"), with_trajectory=False)
    dets = torch.zeros((S, D, 6), device="cuda")
    dets[:, :40, :4] = grid_boxes(S, 40, 10.0, 6.0, 0.5, g)
    cnt = torch.full((S,), 40, dtype=torch.int32, device="cuda")
    out["c3_update_ms"] = time_cuda(lambda: bank.update(dets, cnt, with_trajectory=False))
    out["c3_predict_ms"] = time_cuda(lambda: bank.predict_only())
    bank.close()
    # pipeline size: ~600 live tracks per stream, 300 sparse detections per frame
    S, C, D = 256, 2048, 300
    bank = TrackerBank(S, C, D, 150, 1, 0.1)
    d2 = torch.zeros((S, D, 6), device="cuda")
    d2[:, :, :4] = grid_boxes(S, D, 9.0, 6.0, 0.0, g)
    c2 = torch.full((S,), D, dtype=torch.int32, device="cuda")
    for _ in range(3):
        bank.update(d2, c2, with_trajectory=False)
    out["pipe_sparse_update_ms"] = time_cuda(lambda: bank.update(d2, c2, with_trajectory=False))
    bank.close()
    # dense scene (what a random-weight detector emits): 300 large overlapping boxes per frame, ~25 k pairs with IoU >= 0.1
    bank = TrackerBank(S, C, D, 150, 1, 0.1)
    def dense():
        cx = torch.rand((S, D), device="cuda", generator=g) * 640.0
        cy = torch.rand((S, D), device="cuda", generator=g) * 512.0
        w = 30.0 + torch.rand((S, D), device="cuda", generator=g) * 60.0
        h = 30.0 + torch.rand((S, D), device="cuda", generator=g) * 80.0
        d = torch.zeros((S, D, 6), device="cuda")
        d[..., 0], d[..., 1], d[..., 2], d[..., 3], d[..., 4] = cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2, 0.9
        return d
    base = dense()
    for _ in range(3):
        bank.update((base + torch.randn((S, D, 1), device="cuda", generator=g) * 2.0).contiguous(), c2, with_trajectory=False)
    frames = [(base + torch.randn((S, D, 1), device="cuda", generator=g) * 2.0).contiguous() for _ in range(10)]
    it = iter(frames)
    out["pipe_dense_update_ms"] = time_cuda(lambda: bank.update(next(it), c2, with_trajectory=False))
    bank.close()
    print(out)


if __name__ == "__main__":
    main()
