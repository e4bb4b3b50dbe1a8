"""GPU probe: fused attention vs torch fp32 reference (and timing)."""
import os, sys, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vilco_b200 import ops
bad = 0
for prec in ("bf16x3", "bf16"):
    ops.set_precision(prec)
    for (B, H, Tq, Tk, valid) in [(1, 2, 128, 128, [128]), (2, 2, 256, 256, [256, 100]), (2, 16, 1024, 1024, [1024, 611]),
                                  (2, 4, 512, 57, [57, 30]), (1, 2, 64, 64, [40]), (2, 2, 2, 2, [2, 1]), (8, 16, 1024, 1024, [1024] * 8)]:
        torch.manual_seed(0)
        C = H * 64
        q32, k32, v32 = (torch.randn(B, T, C, device="cuda") for T in (Tq, Tk, Tk))
        km = (torch.arange(Tk, device="cuda")[None, :] < torch.tensor(valid, device="cuda")[:, None]).float().contiguous()
        v32 = v32 * km[:, :, None]
        q, k, v = ops.split16(q32), ops.split16(k32), ops.split16(v32)
        o = ops.attention(q, k, v, km, H, 0.125)
        torch.cuda.synchronize()
        qh = ops.merge16(q).view(B, Tq, H, 64).permute(0, 2, 1, 3)
        kh = ops.merge16(k).view(B, Tk, H, 64).permute(0, 2, 1, 3)
        vh = ops.merge16(v).view(B, Tk, H, 64).permute(0, 2, 1, 3)
        att = (qh * 0.125) @ kh.transpose(-1, -2)
        att = att.masked_fill(km[:, None, None, :] == 0, float("-inf")).softmax(-1)
        ref = (att @ vh).permute(0, 2, 1, 3).reshape(B, Tq, C)
        err = ((ops.merge16(o) - ref).abs().max() / ref.abs().max()).item()
        tol = 2e-5 if prec == "bf16x3" else 2e-2
        ok = err < tol and torch.isfinite(ops.merge16(o)).all().item()
        bad += 0 if ok else 1
        msg = ""
        if B == 8:
            for _ in range(3):
                ops.attention(q, k, v, km, H, 0.125)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                ops.attention(q, k, v, km, H, 0.125)
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / 20
            msg = f"  {us:.1f} us  ({4.0 * B * H * Tq * Tk * 64 / us / 1e6:.1f} TFLOP/s algorithmic)"
        print(("OK " if ok else "BAD"), prec, (B, H, Tq, Tk), f"rel err {err:.2e}", msg, flush=True)
sys.exit(1 if bad else 0)
