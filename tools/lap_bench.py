"""Timing of the association kernels (b2_iou_cost, b2_linear_assignment) batched over S problems.  usage: python tools/lap_bench.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import b200dt  # noqa
from b200dt import _lib, byte_tracker as bt


def boxes(g, S, n, span):
    c = g.uniform(0, span, (S, n, 2)); s = g.uniform(6, 40, (S, n, 2))
    return np.concatenate([c - s / 2, c + s / 2], 2).astype(np.float32)


def main():
    g = np.random.default_rng(0)
    lib = _lib.load()
    out = []
    for S, n, m in [(256, 24, 24), (256, 64, 64), (256, 128, 128), (64, 300, 300), (1, 300, 300), (1, 24, 24)]:
        a = boxes(g, S, n, 600)
        b = np.concatenate([a[:, :min(n, m)] + g.normal(0, 2, (S, min(n, m), 4)).astype(np.float32), boxes(g, S, max(m - n, 0), 600)], 1)[:, :m]
        da, db = torch.as_tensor(a).cuda(), torch.as_tensor(np.ascontiguousarray(b)).cuda()
        sc = torch.rand((S, m), device="cuda")
        cost = torch.empty((S, n, m), device="cuda")
        x = torch.empty((S, n), dtype=torch.int32, device="cuda"); y = torch.empty((S, m), dtype=torch.int32, device="cuda")
        def run():
            _lib.check(lib.b2_iou_cost(_lib.ptr(da), _lib.ptr(db), _lib.ptr(sc), None, None, S, n, m, _lib.ptr(cost), _lib.stream_ptr()))
            _lib.check(lib.b2_linear_assignment(_lib.ptr(cost), None, None, S, n, m, 0.8, _lib.ptr(x), _lib.ptr(y), _lib.stream_ptr()))
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[5]
        out.append({"problems": S, "tracks": n, "detections": m, "ms": ms, "matched_mean": float((x >= 0).sum(1).float().mean()),
                    "problems_per_s": S / ms * 1e3})
        print(json.dumps(out[-1]), flush=True)


if __name__ == "__main__":
    main()
