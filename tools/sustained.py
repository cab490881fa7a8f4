"""Sustained detect+track step time of the bench workload: W warm-up steps (the board reaches its power-capped clocks after a
few hundred ms), then K timed steps.  One variant per process (B2DT_LIB / plan-time switches from the environment), so that
experiment builds can be compared on one box:  python tools/sustained.py [K] [W]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import b200dt  # noqa
from b200dt.pipeline import DetectTrackPipeline


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    S = 256
    fr = torch.from_numpy(bench.make_frames(S, 4)).cuda()
    pipe = DetectTrackPipeline(bench.MODEL, S, bench.FRAME_HW, 640, bench.CONF, bench.IOU, 300, capacity=2048, overlap_post=True,
                               max_tracks_out=256, **bench.TRACKER)
    for k in range(W):
        pipe.step_device(fr[k % 4])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(K):
        pipe.step_device(fr[k % 4])
    pipe.join(); e1.record(); torch.cuda.synchronize()
    tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("B2"))
    print(f"[{tag}] {e0.elapsed_time(e1) / K:.3f} ms/step over {K} steps ({S * K / e0.elapsed_time(e1) * 1e3:.0f} frames/s)", flush=True)


if __name__ == "__main__":
    main()
