"""GPU micro-benchmark of the GEMM shapes of one MQ evaluation step (B=8): prints us and algorithmic TFLOP/s."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vilco_b200 import ops  # noqa: E402
from vilco_b200 import lib as L  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
ops.set_precision(prec)
dev = "cuda"
B, T, C, H = 8, 1024, 1024, 16


def timeit(fn, flops, name, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    print(f"{prec:7s} {name:44s} {us:9.1f} us  {flops / us / 1e6:8.1f} TFLOP/s", flush=True)


x = ops.split16(torch.randn(B, T, C, device=dev))
w = ops.split16(torch.randn(C, C, device=dev) * 0.03)
w4 = ops.split16(torch.randn(4 * C, C, device=dev) * 0.03)
w4b = ops.split16(torch.randn(C, 4 * C, device=dev) * 0.03)
h4 = ops.split16(torch.randn(B, T, 4 * C, device=dev))
w3 = ops.split16(torch.randn(3, C, C, device=dev) * 0.03)
bias = torch.randn(C, device=dev)
bias4 = torch.randn(4 * C, device=dev)
rowmul = torch.ones(B * T, device=dev)
resid = torch.randn(B, T, C, device=dev)
cs = torch.randn(C, device=dev)
timeit(lambda: ops.linear(x, w, ops.bf16, bias=bias), 2.0 * B * T * C * C, "linear 8192x1024x1024 -> bf16")
timeit(lambda: ops.linear(x, w, ops.f32, bias=bias, rowmul=rowmul, colscale=cs, resid=resid, resid_masked=True),
       2.0 * B * T * C * C, "linear 8192x1024x1024 -> f32 +resid")
timeit(lambda: ops.linear(x, w4, ops.bf16, bias=bias4, act=ops.ACT_GELU), 2.0 * B * T * C * 4 * C, "mlp1 8192x4096x1024 gelu -> bf16")
timeit(lambda: ops.linear(h4, w4b, ops.f32, bias=bias, rowmul=rowmul, colscale=cs, resid=resid), 2.0 * B * T * C * 4 * C,
       "mlp2 8192x1024x4096 -> f32 +resid")
timeit(lambda: ops.conv3(x, w3, ops.f32, rowmul=rowmul.view(B, T)), 2.0 * B * T * C * C * 3, "conv3 8x1024 1024->1024 -> f32")
q = ops.split16(torch.randn(B, T, C, device=dev))
k = ops.split16(torch.randn(B, T, C, device=dev))
v = ops.split16(torch.randn(B, T, C, device=dev))
timeit(lambda: ops.attn_scores(q, k, H, 0.125), 2.0 * B * H * T * T * 64, "QK^T 128 x (1024x1024x64) -> f32")
S = ops.attn_scores(q, k, H, 0.125)
mask = torch.ones(B, T, device=dev)
timeit(lambda: ops.softmax_rows(S, mask), 1.0, "softmax rows 128x1024x1024")
P = ops.softmax_rows(S, mask)
timeit(lambda: ops.attn_pv(P, v, H, T), 2.0 * B * H * T * T * 64, "P@V 128 x (1024x64x1024) -> bf16")
