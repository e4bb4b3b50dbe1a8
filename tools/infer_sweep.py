"""Inference + soft-NMS over clip lengths 256 .. max_seq_len (BASELINE.json configs[2]).  The reference pads every clip
to max_seq_len in evaluation (meta_archs.py:1163-1164), so the cost is flat in T; this prints the measured videos/s per
length (CUDA-graph replay, device-resident inputs) and the detection counts.   python tools/infer_sweep.py [B]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
model = bench.build_model().cuda().eval()
g = model.make_eval_graph(B)
for T in (256, 512, 768, 1024):
    vids = bench.synth_videos(B, seed=T, pin=True)
    for v in vids:
        v["feats"] = v["feats"][:, :T].contiguous().pin_memory()
    res = g.run(vids)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    n_det = float(np.mean([len(r["scores"]) for r in res]))
    last = max(float(r["segments"].max()) for r in res)
    print(f"T={T:5d}: {ms:7.2f} ms / {B} clips = {B / ms * 1e3:7.1f} videos/s   detections/clip {n_det:.0f}, latest segment end {last:.1f} s")
