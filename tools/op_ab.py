"""Per-launch A/B of one plan-time environment switch on the bench workload (eager per-launch CUDA events, median of five passes):
launches that differ by more than 5 % are flagged.  usage: [B2_CHAIN=0] python tools/op_ab.py B2_CONV_MMAW=-,1   ('-' = unset)"""
import sys,os,json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np, b200dt
from b200dt import cfg, engine, synth, weights
S,H,W=[int(v) for v in os.environ.get("OP_AB_SHAPE","256,512,640").split(",")]      # OP_AB_MODEL / OP_AB_SHAPE: another plan (e.g. yolov8x-p2, 32,1280,1280)
spec=cfg.resolve(os.environ.get("OP_AB_MODEL","yolov8s-p2")); sd=weights.synthetic_state_dict(spec,seed=0)
fr=[synth.IRStream(seed=1000+s,h=H,w=W).frame() for s in range(8)]
frames=torch.from_numpy(np.stack([fr[s%8] for s in range(S)])).cuda()
var,vals=sys.argv[1].split("=")
res={}
for v in vals.split(","):
    if v=="-": os.environ.pop(var,None)
    else: os.environ[var]=v
    eng=engine.Engine(spec,sd,S,H,W,fuse_head=True)
    for _ in range(2): eng.profile_u8(frames)
    ps=[eng.profile_u8(frames) for _ in range(5)]
    res[v]=[(p["desc"],sorted(q[i]["ms"] for q in ps)[2]) for i,p in enumerate(ps[0])]
    eng.close(); del eng; torch.cuda.empty_cache()
vs=vals.split(",")
a,b=res[vs[0]],res[vs[1]]
if len(a)==len(b):
    for i,((d,x),(_,y)) in enumerate(zip(a,b)):
        flag="  <<<" if y<0.95*x else ("  >>>" if y>1.05*x else "")
        print(f"{i:3d} {x*1e3:7.1f} {y*1e3:7.1f}  {d}{flag}")
else:      # the plans differ in length (a switch that changes the chained launches): list both, matched by description where unique
    from collections import Counter
    ca,cb=Counter(d for d,_ in a),Counter(d for d,_ in b)
    tb={d:y for d,y in b if cb[d]==1}
    for i,(d,x) in enumerate(a):
        y=tb.get(d) if ca[d]==1 else None
        print(f"{i:3d} {x*1e3:7.1f} "+(f"{y*1e3:7.1f}" if y is not None else "      -")+f"  {d}")
    for i,(d,y) in enumerate(b):
        if not (ca[d]==1 and cb[d]==1): print(f"  b{i:3d}         {y*1e3:7.1f}  {d}")
print("total",sum(x for _,x in a),sum(y for _,y in b), len(a), len(b))
