"""One eager evaluation step (B clips) for ncu: python tools/profile_step.py [B] [precision] [nsteps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from vilco_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
if len(sys.argv) > 2:
    ops.set_precision(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
model = bench.build_model().cuda().eval()
vids = bench.synth_videos(B, 0)
model.packed_weights()        # weight packing (a one-off burst of small torch kernels) stays out of the per-step windows
torch.cuda.synchronize()
with torch.no_grad():
    for _ in range(n):
        res = model(vids, is_training=False)
torch.cuda.synchronize()
print("done", len(res), res[0]["segments"].shape)
