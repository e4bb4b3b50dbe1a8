"""Development probe (not a test): which stages of the NLQ network must stay in the exact operand mode for the single-plane
policy to meet the 1e-3 bar?  Every stage boundary is fp32, so stages are switched independently and the error against the
reference's NLQ goldens is measured for each choice.   python tests/nlq_policy_probe.py"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
from test_gpu_zz_nlq import _build  # noqa: E402
from util import rel_max  # noqa: E402

g = np.load(os.path.join(HERE, "golden", "nlq_small.npz"))
model, clips = _build()


def errs():
    out = []
    for i, clip in enumerate(clips):
        logits, offsets, _ = model([clip], is_training=False, get_emb=True)
        lg, of = torch.cat(logits, 1)[0].cpu().numpy(), torch.cat(offsets, 1)[0].cpu().numpy()
        rl = np.concatenate([g[f"logits_{i}_{l}"].reshape(-1, 1) for l in range(7)])
        ro = np.concatenate([g[f"offsets_{i}_{l}"].reshape(-1, 2) for l in range(7)])
        out += [rel_max(lg, rl), rel_max(of, ro)]
    return out


def timeit(n=10):
    for _ in range(2):
        model([clips[0]], is_training=False, get_emb=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        model([clips[0]], is_training=False, get_emb=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


STAGES = ["vid_embd", "txt_embd", "txt_stem", "vid_stem", "branch", "heads"]
model.operand_mode = "fp16x3"
print("all exact        ", ["%.1e" % e for e in errs()], "%.2f ms" % timeit())
model.operand_mode = "mixed"
model.exact_stages = ()
print("all single-plane ", ["%.1e" % e for e in errs()], "%.2f ms" % timeit())
for s in STAGES + [f"vid_stem.{i}" for i in range(4)] + [f"branch.{i}" for i in range(6)] + [f"txt_stem.{i}" for i in range(4)]:
    model.exact_stages = (s,)
    print(f"exact: {s:12s}", ["%.1e" % e for e in errs()])
for combo in (("vid_stem", "branch"), ("vid_stem", "txt_stem"), ("vid_stem", "branch", "txt_stem"), ("heads", "vid_stem"),
              ("txt_embd", "txt_stem"), ("vid_embd", "vid_stem")):
    model.exact_stages = combo
    print("exact:", combo, ["%.1e" % e for e in errs()], "%.2f ms" % timeit())
