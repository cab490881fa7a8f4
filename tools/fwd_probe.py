"""Forward-only timing of the bench workload: K back-to-back forward graphs vs the eager per-launch sum, with and without
programmatic dependent launch, SM clock and power sampled while the graphs run.  usage: python tools/fwd_probe.py"""
import os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import b200dt  # noqa
from b200dt import cfg, engine, weights


def sample(stop, out):
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active", "--format=csv,noheader,nounits", "-lms", "20"],
                         stdout=subprocess.PIPE, text=True)
    while not stop.is_set():
        line = p.stdout.readline()
        if line:
            out.append(line.strip())
    p.terminate()


def main():
    S = 256
    fr = torch.from_numpy(bench.make_frames(S, 4)).cuda()
    spec = cfg.resolve(bench.MODEL)
    sd = weights.synthetic_state_dict(spec, seed=0)
    for pdl in ("1", "0"):
        os.environ["B2_CONV_PDL"] = pdl
        eng = engine.Engine(spec, sd, S, 512, 640, fuse_head=True)
        for k in range(3):
            eng.forward_u8(fr[k % 4], 0, 0)
        torch.cuda.synchronize()
        for K, gap in ((20, 0.0), (100, 0.0), (20, 0.02)):
            stop, smp = threading.Event(), []
            th = threading.Thread(target=sample, args=(stop, smp)); th.start()
            time.sleep(0.3)
            n0 = len(smp)
            ts = []
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if gap == 0.0:
                e0.record()
                for k in range(K):
                    eng.forward_u8(fr[k % 4], 0, 0)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / K
            else:
                for k in range(K):
                    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); eng.forward_u8(fr[k % 4], 0, 0); b_.record(); torch.cuda.synchronize()
                    ts.append(a.elapsed_time(b_)); time.sleep(gap)
                ms = sorted(ts)[len(ts) // 2]
            stop.set(); th.join()
            print(f"PDL={pdl} K={K} gap={gap}: {ms:.3f} ms/forward; smi during: {smp[n0:n0 + 6]} ... {smp[-3:]}", flush=True)
        prof = [eng.profile_u8(fr[k % 4]) for k in range(4)][1:]
        tot = sorted(sum(p["ms"] for p in pr) for pr in prof)
        print(f"PDL={pdl} eager per-launch sums: {tot}")
        eng.close(); del eng; torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
