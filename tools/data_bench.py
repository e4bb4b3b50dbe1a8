"""Data-path kernel timing: vilco_resize_feats on B raw clips of T_in x 4096 fp32 -> 1024 rows, CUDA events,
HBM roofline (algorithmic bytes = raw clip read once + outputs written once) next to the reference's CPU permute + F.interpolate.

    python tools/data_bench.py [--clips 32] [--t-in 512]
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vilco_b200 import data, ops, lib as L  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=32)
    ap.add_argument("--t-in", type=int, default=512)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    B, T_in, C, T = a.clips, a.t_in, 4096, 1024
    x = torch.randn(B * T_in, C, device="cuda")
    row_start = torch.arange(0, (B + 1) * T_in, T_in, dtype=torch.int64, device="cuda")
    o16 = ops.empty16(B, T, C)
    o32 = torch.empty(B, T, C, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")       # > 126 MB L2
    out = {}
    for name, a32, a16 in (("planes", None, o16), ("fp32", o32, None)):
        ms = []
        for i in range(a.iters + 3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.check(L.lib().vilco_resize_feats(ops._p(x), ops._p(row_start), B, C, T, ops._p(a32), ops._p(a16),
                                               ops._i64(ops.lo(o16) if a16 is not None else 0), L.stream_ptr()), "resize")
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ms.append(e0.elapsed_time(e1))
        t = sorted(ms)[len(ms) // 2] * 1e-3
        nbytes = x.numel() * 4 + (o16.numel() * 2 if a16 is not None else o32.numel() * 4)
        out[name] = {"ms": t * 1e3, "algorithmic_GB": nbytes / 1e9, "GB_per_s": nbytes / t / 1e9, "clips_per_s": B / t}
    xc = x[:T_in].cpu()
    t0 = time.perf_counter()
    for _ in range(5):
        F.interpolate(xc.permute(1, 0).unsqueeze(0), size=T, mode="linear", align_corners=False)
    out["cpu_interpolate_ms_per_clip"] = (time.perf_counter() - t0) / 5 * 1e3
    out["config"] = {"clips": B, "t_in": T_in, "C": C, "t_out": T, "l2_flush": True}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
