"""find the first op whose output differs between two identical forward passes"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from util import build_pair
from oracle import params as PR
from oracle.gen_golden import FULL_VIDEOS, full_cfg
from vilco_b200 import ops, lib as L
mode = sys.argv[1] if len(sys.argv) > 1 else "mixed"
ops.set_precision(mode)
cfg = full_cfg(22)
model, P = build_pair(cfg, seed=4)
videos = PR.synth_video_list(cfg, 2, **FULL_VIDEOS)
rec = []
names = ["linear", "conv3", "attention", "layernorm", "dwconv_ln", "channel_attention", "softmax_rows", "attn_scores", "attn_pv",
         "axpby", "maxpool3s2", "scale_add", "pack_feats"]
orig = {n: getattr(ops, n) for n in names}
def flat(o):
    if o is None: return []
    if isinstance(o, torch.Tensor): return [o]
    out = []
    for x in o: out += flat(x)
    return out
def wrap(n):
    f = orig[n]
    def g(*a, **k):
        o = f(*a, **k)
        desc = ""
        if n == "linear":
            desc = f"x{tuple(a[0].shape)} w{tuple(a[1].shape)} out{k.get('out_dtype', a[2] if len(a) > 2 else None)}"
        rec.append((n + " " + desc, [t.detach().float().clone() for t in flat(o)]))
        return o
    return g
for n in names:
    setattr(ops, n, wrap(n))
def run():
    rec.clear()
    model(videos, is_training=False, get_emb=True)
    torch.cuda.synchronize()
    return list(rec)
r1 = run(); r2 = run()
print("ops recorded", len(r1), len(r2))
shown = 0
for i, ((n1, o1), (n2, o2)) in enumerate(zip(r1, r2)):
    d = max((float((a - b).abs().max()) / (float(b.abs().max()) + 1e-12) for a, b in zip(o1, o2)), default=0.0)
    if d > 0:
        print(f"op {i:4d} {n1:70s} rel diff {d:.3e}")
        shown += 1
        if shown > 25: break
