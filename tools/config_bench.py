"""CUDA-event timing of the other SURVEY.md 8d configurations: C2 (yolov8s-p2, 64 x 640x640 tensors, forward + decode + NMS) and
C5 (yolov8x-p2, 32 x 1280x1280 per GPU).  Prints one JSON line per configuration."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200dt  # noqa: F401
from b200dt import cfg, weights
from b200dt.predictor import DetectPipeline


def run(tag, name, B, HW, iters=8):
    spec = cfg.resolve(name, nc=80)
    pipe = DetectPipeline(spec, weights.synthetic_state_dict(spec, seed=0), B, HW[0], HW[1], 300)
    # input tensors: the bench's synthetic IR frames (noise background + ~20 bright blobs, the scenes the synthetic weights are
    # calibrated on) as BCHW RGB 0-1 tensors.  (Uniform-noise tensors make nearly every anchor a candidate with these weights --
    # 25 000 / 136 000 per image -- and the timing then measures a global-memory sort, not the configured path; --uniform keeps it.)
    if "--uniform" in sys.argv:
        g = torch.Generator(device="cuda").manual_seed(7)
        pool = [torch.rand((B, 3, HW[0], HW[1]), device="cuda", generator=g).to(torch.bfloat16) for _ in range(3)]
    else:
        import numpy as np
        from b200dt import synth
        vids = [synth.IRStream(seed=500 + b, h=HW[0], w=HW[1], n_targets=max(20, 20 * HW[0] * HW[1] // (512 * 640))) for b in range(min(B, 8))]
        pool = []
        for k in range(3):
            fr = np.stack([v.frame() for v in vids])[..., ::-1].copy()                      # BGR -> RGB
            t = torch.from_numpy(fr).cuda().permute(0, 3, 1, 2).float().div_(255.0).to(torch.bfloat16)
            pool.append(t.repeat((B + len(vids) - 1) // len(vids), 1, 1, 1)[:B].contiguous())
    for k in range(3):
        pipe.run_tensor(pool[k % 3], 0.15, 0.6)
    torch.cuda.synchronize()
    fwd, tot = [], []
    for k in range(iters):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(); pipe.engine.forward_tensor(pool[k % 3]); e[1].record()
        pipe.finish(0.15, 0.6, HW, None, False, "exact", None); e[2].record()
        torch.cuda.synchronize()
        fwd.append(e[0].elapsed_time(e[1])); tot.append(e[0].elapsed_time(e[2]))
    f, t = sorted(fwd)[len(fwd) // 2], sorted(tot)[len(tot) // 2]
    fl = pipe.engine.flops_per_image * B
    print(json.dumps({"config": tag, "model": name, "batch": B, "hw": HW, "forward_ms": f, "forward_decode_nms_ms": t, "images_per_s": B / t * 1e3,
                      "conv_tflops_forward": fl / f / 1e9, "gflop_per_image": pipe.engine.flops_per_image / 1e9,
                      "mean_candidates": float(pipe.post.cand_count.float().mean())}), flush=True)


if __name__ == "__main__":
    run("C2", "yolov8s-p2", 64, (640, 640))
    run("C5", "yolov8x-p2", 32, (1280, 1280))
