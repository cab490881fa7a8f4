"""Candidate-count distribution and NMS / association timings of the bench workload (256 streams)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import b200dt  # noqa
from b200dt.pipeline import DetectTrackPipeline


def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    pipe = DetectTrackPipeline(bench.MODEL, S, bench.FRAME_HW, 640, bench.CONF, bench.IOU, 300, capacity=2048, **bench.TRACKER)
    fr = torch.from_numpy(bench.make_frames(S, 4)).cuda()
    for k in range(12):
        pipe.step_device(fr[k % 4])
    torch.cuda.synchronize()
    post = pipe.detect.post
    c = post.cand_count.cpu().numpy()
    print("candidates per image: min %d mean %.1f median %d p90 %d max %d" % (c.min(), c.mean(), np.median(c), np.percentile(c, 90), c.max()))
    oc = post.out_count.cpu().numpy()
    print("kept per image: min %d mean %.1f max %d" % (oc.min(), oc.mean(), oc.max()))
    print("nms ms:", t(lambda: post.nms(bench.IOU, scale=None)))
    dets, counts = post.out, post.out_count
    print("tracker update ms:", t(lambda: pipe.bank.update(dets, counts, with_trajectory=False)))
    print("bank predict ms:", t(lambda: pipe.bank.predict_only()))
    print("candidates ms:", t(lambda: pipe.detect.candidates(bench.CONF)))


if __name__ == "__main__" and len(sys.argv) <= 2:
    main()


def sweep():
    S = 256
    pipe = DetectTrackPipeline(bench.MODEL, S, bench.FRAME_HW, 640, bench.CONF, bench.IOU, 300, capacity=2048, **bench.TRACKER)
    fr = torch.from_numpy(bench.make_frames(S, 4)).cuda()
    for k in range(4):
        pipe.step_device(fr[k % 4])
    torch.cuda.synchronize()
    post = pipe.detect.post
    orig = post.cand_count.clone()
    for cap in (100000, 2048, 1408, 704, 512, 256, 128):
        post.cand_count.copy_(orig.clamp(max=cap))
        print("cap", cap, "nms ms %.4f" % t(lambda: post.nms(bench.IOU, scale=None)), "kept mean %.1f" % post.out_count.float().mean().item())
    # only the heavy image / only the light ones
    heavy = int(orig.argmax())
    z = torch.zeros_like(orig); z[heavy] = orig[heavy]
    post.cand_count.copy_(z); print("heavy image alone (n=%d): %.4f ms" % (int(orig[heavy]), t(lambda: post.nms(bench.IOU, scale=None))))
    z = orig.clone(); z[orig > 1000] = 0
    post.cand_count.copy_(z); print("images with <= 1000 candidates (%d): %.4f ms" % (int((z > 0).sum()), t(lambda: post.nms(bench.IOU, scale=None))))


if len(sys.argv) > 2 and sys.argv[2] == "sweep":
    sweep()


def stamps():
    """B2_NMS_DEBUG=1: per-phase clock64 stamps of a few images."""
    S = 256
    pipe = DetectTrackPipeline(bench.MODEL, S, bench.FRAME_HW, 640, bench.CONF, bench.IOU, 300, capacity=2048, **bench.TRACKER)
    fr = torch.from_numpy(bench.make_frames(S, 4)).cuda()
    for k in range(4):
        pipe.step_device(fr[k % 4])
    torch.cuda.synchronize()
    post = pipe.detect.post
    post.nms(bench.IOU, scale=None); torch.cuda.synchronize()
    P = 1
    while P < post.cand_cap: P *= 2
    ws = post.ws.cpu().numpy()
    off = (-post.ws.data_ptr()) % 256
    c = post.cand_count.cpu().numpy()
    for b in list(np.argsort(c)[[0, 64, 128, 192, 240, 255]]):
        st = ws[off + b * P * 16: off + b * P * 16 + 128].view(np.int64)
        k = int(st[15])
        d = np.diff(st[:k])
        print("image %3d n=%4d kept=%3d stamps(cycles): %s" % (b, c[b], int(post.out_count[b]), " ".join(str(int(x)) for x in d)))


if len(sys.argv) > 2 and sys.argv[2] == "stamps":
    stamps()


def insitu():
    """Per-stage CUDA-event times of the post-processing on the varying frames of the bench loop."""
    S = 256
    pipe = DetectTrackPipeline(bench.MODEL, S, bench.FRAME_HW, 640, bench.CONF, bench.IOU, 300, capacity=2048, **bench.TRACKER)
    fr = torch.from_numpy(bench.make_frames(S, 4)).cuda()
    d = pipe.detect
    acc = {"forward": 0.0, "candidates": 0.0, "nms": 0.0, "tracker": 0.0}
    N = 16
    for k in range(N + 6):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record(); d.engine.forward_u8(fr[k % 4], pipe.top, pipe.left)
        ev[1].record(); d.candidates(pipe.conf)
        ev[2].record(); dets, counts = d.nms(pipe.iou, (pipe.h0, pipe.w0), False, "exact")
        ev[3].record(); pipe.bank.update(dets, counts, with_trajectory=False)
        ev[4].record(); torch.cuda.synchronize()
        if k >= 6:
            for i, key in enumerate(acc): acc[key] += ev[i].elapsed_time(ev[i + 1]) / N
    print({k: round(v, 4) for k, v in acc.items()})


if len(sys.argv) > 2 and sys.argv[2] == "insitu":
    insitu()
