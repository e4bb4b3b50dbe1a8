"""GPU sanity: a few dozen optimisation steps on one fixed synthetic batch must drive the loss down (small config, full
training semantics incl. dropout).  python tools/train_sanity.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import gen_golden as GG  # noqa: E402
from oracle import params as PR  # noqa: E402
import util  # noqa: E402
from vilco_b200.trainer import Trainer, make_optimizer  # noqa: E402

cfg = GG.small_cfg()
model, P = util.build_pair(cfg, 0)
model.train()
model.loss_normalizer_momentum = 1.0
videos = PR.synth_video_list(cfg, 4, seed=3, lens=[128, 100, 90, 128], text_lens=[40, 57, 33, 64], n_gt=[3, 2, 4, 1])
opt = make_optimizer(model, {"type": "AdamW", "learning_rate": 3e-4, "weight_decay": 0.05}, flat=True)
tr = Trainer(model, opt, clip_grad_l2norm=1.0)
hist = []
for it in range(60):
    lo = tr.step(videos)
    hist.append(float(lo["final_loss"].detach()))
    if it % 10 == 0 or it == 59:
        print(it, {k: round(float(v.detach()), 4) for k, v in lo.items()}, "grad norm", round(float(opt.grad_norm()), 3))
model.eval()
with torch.no_grad():
    res = model(videos, is_training=False)
print("first / last 5-step mean:", sum(hist[:5]) / 5, sum(hist[-5:]) / 5, "detections", [len(r["scores"]) for r in res])
print("OK" if sum(hist[-5:]) < 0.6 * sum(hist[:5]) else "BAD")
