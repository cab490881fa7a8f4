"""Stand-alone timing of conv2d_bf16 shapes (CUDA events, L2 flushed between iterations).
usage: python tools/conv_bench.py [--iters N] [--ab VAR=a,b,..] B,H,W,Cin,Cout,k,s[,res] ...
--ab: time the variants of one plan-time environment switch (B2_CONV_ACC, B2_CONV_MMAW, ...) alternately, launch by launch,
inside one process (boxes and clock states differ by more than the effects being measured)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200dt
from b200dt import ops

ACT = os.environ.get("CB_ACT", "1") == "1"


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    iters = 5
    if "--iters" in sys.argv:
        iters = int(sys.argv[sys.argv.index("--iters") + 1]); args = [a for a in args if a != str(iters)]
    ab_var, ab_vals = None, [None]
    if "--ab" in sys.argv:
        spec = sys.argv[sys.argv.index("--ab") + 1]; args = [a for a in args if a != spec]
        ab_var, vals = spec.split("="); ab_vals = vals.split(",")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = []
    for spec in args:
        v = [int(x) for x in spec.split(",")]
        B, H, W, Cin, Cout, k, s = v[:7]
        res = len(v) > 7 and v[7]
        x = torch.randn((B, H, W, Cin), device="cuda").to(torch.bfloat16)
        w = (torch.randn((Cout, k, k, Cin), device="cuda") / (Cin * k * k) ** 0.5).to(torch.bfloat16)
        b = torch.randn((Cout,), device="cuda")
        Ho, Wo = (H + 2 * (k // 2) - k) // s + 1, (W + 2 * (k // 2) - k) // s + 1
        y = torch.empty((B, Ho, Wo, max(Cout, 8)), device="cuda", dtype=torch.bfloat16)
        r = torch.randn((B, Ho, Wo, Cout), device="cuda").to(torch.bfloat16) if res else None
        ops.conv2d_bf16(x, w, b, k, s, ACT, out=y, residual=r)
        ts = {v: [] for v in ab_vals}
        for _ in range(iters):
            for v in ab_vals:
                if ab_var:
                    os.environ[ab_var] = v
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); ops.conv2d_bf16(x, w, b, k, s, ACT, out=y, residual=r); e1.record()
                torch.cuda.synchronize(); ts[v].append(e0.elapsed_time(e1))
        fl = 2 * B * Ho * Wo * Cout * Cin * k * k
        by = (B * H * W * Cin + B * Ho * Wo * Cout * (2 if res else 1)) * 2
        for v in ab_vals:
            ms = sorted(ts[v])[len(ts[v]) // 2]
            out.append({"shape": spec, "variant": f"{ab_var}={v}" if ab_var else "", "ms": ms, "tflops": fl / ms / 1e9, "gbs": by / ms / 1e6})
            print(out[-1], flush=True)
    return out

if __name__ == "__main__":
    main()
