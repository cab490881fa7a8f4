"""A/B of the chained-launch patterns (engine.CHAIN_*): per-pattern conv time of one forward of the bench workload, measured in
ONE process with per-launch CUDA events (b2_engine_profile_u8), alternating the variants.  usage: python tools/ab_chain.py [masks..]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import b200dt  # noqa: F401
from b200dt import cfg, engine, synth, weights


def main():
    masks = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 4, 8, 15]
    S, H, W = 256, 512, 640
    spec = cfg.resolve("yolov8s-p2")
    sd = weights.synthetic_state_dict(spec, seed=0)
    vids = [synth.IRStream(seed=1000 + s, h=H, w=W) for s in range(8)]
    fr = [v.frame() for v in vids]
    frames = torch.from_numpy(np.stack([fr[s % 8] for s in range(S)])).cuda()
    out = {}
    for m in masks:
        os.environ["B2_CHAIN"] = str(m)
        eng = engine.Engine(spec, sd, S, H, W, fuse_head=True)
        for _ in range(2):
            eng.profile_u8(frames)
        best = None
        for _ in range(5):
            prof = eng.profile_u8(frames)
            conv = [p for p in prof if p["op"] == "conv"]
            t = sum(p["ms"] for p in conv)
            if best is None or t < best[0]:
                best = (t, sum(p["ms"] for p in prof), sum(p["flops"] for p in conv), prof)
        out[m] = {"conv_ms": best[0], "forward_ms": best[1], "conv_tflops": best[2] / best[0] / 1e9, "launches": len(best[3])}
        print(m, out[m], flush=True)
        if "--ops" in os.environ.get("AB_FLAGS", ""):
            for i, p in enumerate(best[3]):
                if "1x1" in p["desc"]:
                    print("   ", i, f"{p['ms'] * 1e3:7.1f} us", p["desc"])
        eng.close()
        del eng
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
