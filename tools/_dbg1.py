import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from util import build_pair, rel_max
from oracle import params as PR
from oracle.gen_golden import FULL_VIDEOS, full_cfg
from vilco_b200 import ops
cfg = full_cfg(22)
model, P = build_pair(cfg, seed=4)
videos = PR.synth_video_list(cfg, 2, **FULL_VIDEOS)
def lg(vs):
    c, o, m = model(vs, is_training=False, get_emb=True)
    return torch.cat(c, 1).clone(), torch.cat(o, 1).clone()
for mode in ("mixed", "fp16x3"):
    ops.set_precision(mode)
    a = lg(videos); b = lg(videos)
    print(mode, "eager batched twice: logits", rel_max(a[0], b[0]), "offsets", rel_max(a[1], b[1]))
    s0 = lg(videos[:1]); s0b = lg(videos[:1])
    print(mode, "eager single twice:", rel_max(s0[0], s0b[0]), " batched[0] vs single:", rel_max(a[0][:1], s0[0]))
    eg = model.make_eval_graph(2, text_len=64)
    eg.load_inputs(videos); eg.replay(); torch.cuda.synchronize()
    o1 = [t.clone() for t in eg.out]
    eg.replay(); torch.cuda.synchronize()
    o2 = [t.clone() for t in eg.out]
    print(mode, "graph twice: scores", float((o1[1] - o2[1]).abs().max()))
    # graph's logits: run _device_forward eagerly on the graph's static inputs
    l1, f1, pm, pyr = model._device_forward(eg.feats, eg.mask, eg.text, eg.tmask, eg.tlens, False)
    l2, f2, _, _ = model._device_forward(eg.feats, eg.mask, eg.text, eg.tmask, eg.tlens, False)
    print(mode, "static-input forward twice:", rel_max(l1, l2))
    cat = torch.cat([l1[:, o:o + n] for o, n in zip(pyr.off, pyr.lens)], 1)
    print(mode, "static-input (text pad 64) vs eager batched (text pad 57):", rel_max(cat, a[0]))
