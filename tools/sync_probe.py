"""Lists the host-device synchronisation points of one training step (torch.cuda.set_sync_debug_mode("warn") with the Python
stack of each distinct site): python tools/sync_probe.py [clips].  Round 2: 34 -> 7 per step."""
import os, sys, warnings, traceback
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from vilco_b200.trainer import Trainer, make_optimizer
model = bench.build_model().cuda().train()
opt = make_optimizer(model, {"type": "AdamW", "learning_rate": 1e-4, "weight_decay": 0.05}, flat=True)
tr = Trainer(model, opt, clip_grad_l2norm=1.0)
vids = bench.synth_videos(int(sys.argv[1]) if len(sys.argv) > 1 else 8, seed=0)
for _ in range(3):
    tr.step(vids)
torch.cuda.synchronize()
seen = {}
def showwarning(message, category, filename, lineno, file=None, line=None):
    st = "".join(traceback.format_stack(limit=9)[:-1])
    key = st[-400:]
    seen[key] = seen.get(key, 0) + 1
    if seen[key] == 1:
        print("SYNC:", message, "\n", st[-900:], flush=True)
warnings.showwarning = showwarning
warnings.simplefilter("always")
torch.cuda.set_sync_debug_mode("warn")
tr.step(vids)
torch.cuda.set_sync_debug_mode("default")
print("distinct sync sites:", len(seen), "total", sum(seen.values()))
