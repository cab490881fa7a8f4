"""In-process A/B of whole detect+track steps: two pipelines built under different values of one environment switch
(read at plan time), stepped alternately on the same frames.  usage: python tools/ab_pipeline.py VAR=a,b [rounds]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import b200dt  # noqa
from b200dt.pipeline import DetectTrackPipeline


def main():
    var, vals = sys.argv[1].split("=")
    vals = vals.split(",")
    rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    S = 256
    fr = torch.from_numpy(bench.make_frames(S, 4)).cuda()
    pipes = {}
    for v in vals:
        os.environ[var] = v
        pipes[v] = DetectTrackPipeline(bench.MODEL, S, bench.FRAME_HW, 640, bench.CONF, bench.IOU, 300, capacity=2048, overlap_post=True, **bench.TRACKER)
        for k in range(3):
            pipes[v].step_device(fr[k % 4])
        pipes[v].join(); torch.cuda.synchronize()
    ts = {v: [] for v in vals}
    for r in range(rounds):
        for v in vals:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(4):
                pipes[v].step_device(fr[k])
            pipes[v].join(); e1.record(); torch.cuda.synchronize()
            ts[v].append(e0.elapsed_time(e1) / 4)
    for v in vals:
        t = sorted(ts[v])
        print(f"{var}={v}: median {t[len(t) // 2]:.3f} ms/step  min {t[0]:.3f}  max {t[-1]:.3f}")


if __name__ == "__main__":
    main()
