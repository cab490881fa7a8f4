"""The HBM-judged kernels at their BASELINE sizes, launched a few times each inside one cudaProfilerStart/Stop range
(for `ncu --profile-from-start off`) after a warm-up pass: decode (C2), head_candidates + nms (C4 step), tracker sweep /
resolve (C3 bank, coasting frame and full frame), bank predict.  L2 is flushed before every launch of interest.

    python tools/hbm_probe.py                       # CUDA-event timings
    ncu --profile-from-start off --set full --clock-control none -k regex:'sweep|resolve|decode_kernel|head_cand|nms_kernel|bank_predict' \\
        -o gpurun_out/hbm_kernels python tools/hbm_probe.py --once
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import b200dt  # noqa: F401
from b200dt import ops, synth
from b200dt.pipeline import DetectTrackPipeline
from b200dt.tracker import TrackerBank

ONCE = "--once" in sys.argv


def main():
    import numpy as np

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    flush = lambda: flush_buf.fill_(1)
    cases = []
    # ---- decode at C2 size ----
    B2, nc2, ls2 = 64, 80, 144
    lh = lw = [160, 80, 40, 20]
    g = torch.Generator(device="cuda").manual_seed(1)
    logits = []
    for h, w in zip(lh, lw):
        t = torch.randn((B2, h * w, ls2), device="cuda", generator=g)
        t[..., 64:] -= 6.0
        logits.append(t.to(torch.bfloat16))
    post2 = ops.DetectPost(B2, lh, lw, [4, 8, 16, 32], nc2, ls2)
    cases.append(("decode_c2", lambda: post2.decode(logits, 0.15), B2 * post2.A * ls2 * 2))
    # ---- fused-head candidates + NMS at the C4 step size ----
    S = 256
    pipe = DetectTrackPipeline("yolov8s-p2", S, (512, 640), 640, 0.15, 0.6, 300, capacity=2048, max_tracks_out=256,
                               max_lost_frames=150, min_hits=1, iou_threshold=0.1)
    vids = [synth.IRStream(seed=1000 + s, h=512, w=640) for s in range(8)]
    fr = [v.frame() for v in vids]
    frames = torch.from_numpy(np.stack([fr[s % 8] for s in range(S)])).cuda()
    pipe.step_device(frames)
    eng, post = pipe.detect.engine, pipe.detect.post
    if eng.fused_head:
        cases.append(("head_candidates_c4", lambda: post.candidates_from_head(eng.head_dist, eng.head_cls, 0.15), S * eng.num_anchors * 8))
    cases.append(("nms_c4", lambda: post.nms(0.6), None))
    # ---- tracker bank at C3 size ----
    Sb, C, L, D = 256, 4608, 4096, 1024
    bank = TrackerBank(Sb, C, D, 150, 1, 0.1)
    for r in range(L // D):
        idx = torch.arange(D, device="cuda") + r * D
        x, y = (idx % 64).float() * 10.0, (idx // 64).float() * 10.0
        boxes = torch.stack([x, y, x + 6, y + 6], 1)[None].repeat(Sb, 1, 1).contiguous()
        bank.update(boxes, torch.full((Sb,), D, dtype=torch.int32, device="cuda"), with_trajectory=False)
    pb, ub = TrackerBank.bytes_per_track()
    zero = torch.zeros((Sb,), dtype=torch.int32, device="cuda")
    dets = torch.zeros((Sb, D, 6), device="cuda")
    cases.append(("tracker_coast_c3", lambda: bank.update(dets, zero, with_trajectory=False), Sb * L * pb))
    Dn = 40
    pick = torch.randint(0, L, (Sb, Dn), device="cuda", generator=g)
    px, py = (pick % 64).float() * 10.0, (pick // 64).float() * 10.0
    clutter = torch.rand((Sb, Dn), device="cuda", generator=g) < 0.3
    px = torch.where(clutter, torch.rand((Sb, Dn), device="cuda", generator=g) * 634.0, px + torch.randn((Sb, Dn), device="cuda", generator=g))
    py = torch.where(clutter, torch.rand((Sb, Dn), device="cuda", generator=g) * 634.0, py + torch.randn((Sb, Dn), device="cuda", generator=g))
    dets2 = torch.zeros((Sb, D, 6), device="cuda")
    dets2[:, :Dn, 0], dets2[:, :Dn, 1], dets2[:, :Dn, 2], dets2[:, :Dn, 3], dets2[:, :Dn, 4] = px, py, px + 6, py + 6, 0.9
    cnt = torch.full((Sb,), Dn, dtype=torch.int32, device="cuda")
    cases.append(("tracker_frame_c3", lambda: bank.update(dets2, cnt, with_trajectory=False), Sb * L * pb + Sb * Dn * (16 + ub)))
    cases.append(("bank_predict_c3", lambda: bank.predict_only(), Sb * L * 140))

    for _, fn, _ in cases:           # warm-up
        fn(); fn()
    torch.cuda.synchronize()
    out = {}
    torch.cuda.cudart().cudaProfilerStart()
    for name, fn, nbytes in cases:
        reps = 1 if ONCE else 10
        ts = []
        for _ in range(reps):
            flush()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        out[name] = {"ms": ms, "bytes": nbytes, "gbs": None if nbytes is None else nbytes / ms / 1e6}
    torch.cuda.cudart().cudaProfilerStop()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
