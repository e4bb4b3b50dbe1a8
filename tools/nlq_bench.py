"""NLQ evaluation throughput on one B200: ego4d_nlq_v2_egovlp_1e-4.yaml (T = 2560, C = 384, 7 levels), synthetic clips, the
model's own operand policy.  python tools/nlq_bench.py [clips per step] [steps]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vilco_b200 import lib as L  # noqa: E402
from vilco_b200.modeling import make_meta_arch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
torch.manual_seed(0)
model = make_meta_arch("NlqLocPointTransformer", regression_range=[[0, 4], [2, 8], [4, 16], [8, 32], [16, 64], [32, 128], [64, 10000]],
                       test_cfg=dict(voting_thresh=0.9, pre_nms_topk=2000, max_seg_num=5, min_score=0.001, nms_sigma=0.75,
                                     duration_thresh=0.001)).cuda().eval()
with torch.no_grad():
    for n, p in model.named_parameters():            # lift the 1e-4 AffineDropPath scales so every branch contributes
        if n.endswith(".scale") and p.dim() == 3:
            p.fill_(1.0)
g = torch.Generator().manual_seed(1)
clips = [{"video_id": f"v{i}", "query_id": f"q{i}", "feats": torch.randn(256, 2560 - 17 * i, generator=g),
          "query_feats": torch.randn(512, 12, generator=g), "fps": 30.0, "duration": 1400.0, "feat_stride": 16.043,
          "feat_num_frames": 16.043} for i in range(B)]
for _ in range(3):
    res = model(clips, is_training=False)
torch.cuda.synchronize()
n0 = L.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    res = model(clips, is_training=False)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
out = {"workload": "NLQ evaluation incl. host batching, upload, decode + soft-NMS and result download (eager, no CUDA graph)",
       "clips_per_step": B, "ms_per_step": ms, "queries_per_s": B / ms * 1e3, "launches_per_step": (L.launch_count() - n0) // steps,
       "operand_mode": model.operand_mode, "segments_per_query": int(res[0]["segments"].shape[0])}
from vilco_b200.modeling.nlq import NlqEvalGraph  # noqa: E402
g = NlqEvalGraph(model, B, 12)
for _ in range(3):
    g.run(clips)
torch.cuda.synchronize()
e0.record()
for _ in range(steps):
    res2 = g.run(clips)
e1.record()
torch.cuda.synchronize()
ms_g = e0.elapsed_time(e1) / steps
out["graph_ms_per_step"], out["graph_queries_per_s"], out["graph_launches"] = ms_g, B / ms_g * 1e3, g.launches
out["graph_equals_eager"] = all(torch.equal(a["scores"], b["scores"]) and torch.equal(a["segments"], b["segments"]) for a, b in zip(res, res2))
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/r2_nlq_bench.json", "w"), indent=1)
