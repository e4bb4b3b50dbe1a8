"""CUDA-event timing + HBM roofline of the sliding-window attention kernel (LocalMaskedMHCA core, csrc/attention.cu) at the MQ
shape (C=1024, H=16, T=1024; W in {9, 17}) and the NLQ shape (C=384, H=4 -> head dim 96, W=9, T=2560).
Algorithmic bytes = 4 T C e per clip (q, k, v read once, out written once; e = 2 bytes, single 16-bit plane)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vilco_b200 import ops  # noqa: E402

ops.set_precision(sys.argv[1] if len(sys.argv) > 1 else "mixed")
peak = 6557.4
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p))["hbm_gbs"]
out = {}
for name, (B, T, C, H, W) in {"mq_w9": (32, 1024, 1024, 16, 9), "mq_w17": (32, 1024, 1024, 16, 17), "nlq_w9": (16, 2560, 384, 4, 9)}.items():
    q, k, v = (ops.split16(torch.randn(B, T, C, device="cuda"), planes=ops.PLANES) for _ in range(3))
    mask = torch.ones(B, T, device="cuda")
    big = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)     # L2 flush between timed launches
    ts = []
    for i in range(8):
        big.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.local_attention(q, k, v, mask, H, W)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1) * 1e3)
    us = sorted(ts)[len(ts) // 2]
    nbytes = 4 * T * C * 2 * B * ops.PLANES
    out[name] = {"B": B, "T": T, "C": C, "H": H, "W": W, "us": us, "GBps": nbytes / us / 1e3, "frac_of_hbm_peak": nbytes / us / 1e3 / peak}
    print(name, out[name], flush=True)
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "r2_local_attn_bench.json"), "w"), indent=1)
