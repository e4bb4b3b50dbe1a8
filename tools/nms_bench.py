"""time the device part of batched NMS on the MQ worst case (B=8 videos x ~20k candidates, 22 classes)"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vilco_b200.utils.nms import _run
from vilco_b200 import lib as L
L.lib()
rs = np.random.RandomState(0)
B, nl, topk, K = int(os.environ.get("NMS_B", 8)), 10, 5000, 22
segs = torch.zeros(B, nl * topk, 2); scores = torch.zeros(B, nl * topk); labels = torch.zeros(B, nl * topk, dtype=torch.int32)
cnt = torch.zeros(B, nl, dtype=torch.int32)
per = [5000, 5000, 5000, 2816, 1408, 704, 352, 176, 88, 44]
for b in range(B):
    for l in range(nl):
        n = per[l]
        c = rs.uniform(0, 1024, n); ln = np.exp(rs.uniform(np.log(2.0), np.log(400.0), n))
        segs[b, l * topk:l * topk + n, 0] = torch.from_numpy((c - ln / 2).astype(np.float32))
        segs[b, l * topk:l * topk + n, 1] = torch.from_numpy((c + ln / 2).astype(np.float32))
        scores[b, l * topk:l * topk + n] = torch.from_numpy(np.sort(rs.uniform(0.009, 0.011, n).astype(np.float32))[::-1].copy())
        labels[b, l * topk:l * topk + n] = torch.from_numpy(rs.randint(0, K, n).astype(np.int32))
        cnt[b, l] = n
segs, scores, labels, cnt = segs.cuda(), scores.cuda(), labels.cuda(), cnt.cuda()
for _ in range(3):
    out = _run(segs, scores, labels, cnt, B, nl, topk, K, True, 2, 0.1, 0.99, 1e-4, 200)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    out = _run(segs, scores, labels, cnt, B, nl, topk, K, True, 2, 0.1, 0.99, 1e-4, 200)
e1.record(); torch.cuda.synchronize()
print(f"batched NMS B={B}, {sum(per)} candidates/video, K={K}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per call; kept {out[3].tolist()}")
