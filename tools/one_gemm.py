"""Single GEMM shape for ncu captures: python tools/one_gemm.py {qk|mlp1|lin32|conv3|heads|wgrad} [precision]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vilco_b200 import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "qk"
ops.set_precision(sys.argv[2] if len(sys.argv) > 2 else "mixed")
dev = "cuda"
B, T, C, H = 8, 1024, 1024, 16
x = ops.split16(torch.randn(B, T, C, device=dev))
if which == "qk":
    k = ops.split16(torch.randn(B, T, C, device=dev))
    fn = lambda: ops.attn_scores(x, k, H, 0.125)
elif which == "mlp1":
    w4 = ops.split16(torch.randn(4 * C, C, device=dev) * 0.03)
    b4 = torch.randn(4 * C, device=dev)
    fn = lambda: ops.linear(x, w4, ops.bf16, bias=b4, act=ops.ACT_GELU)
elif which == "lin32":
    w = ops.split16(torch.randn(C, C, device=dev) * 0.03)
    resid = torch.randn(B, T, C, device=dev)
    rm = torch.ones(B * T, device=dev)
    fn = lambda: ops.linear(x, w, ops.f32, rowmul=rm, resid=resid, resid_masked=True)
elif which == "cxc32":   # C x C projection at 32 clips -> operand planes (q / k / v / proj shapes)
    x32 = ops.split16(torch.randn(32, T, C, device=dev), planes=1)
    w = ops.split16(torch.randn(C, C, device=dev) * 0.03, planes=1)
    bb = torch.randn(C, device=dev)
    fn = lambda: ops.linear(x32, w, ops.bf16, bias=bb, planes=1)
elif which == "mlp2_32":  # FFN down projection at 32 clips: K = 4096, fp32 out + residual
    x32 = ops.split16(torch.randn(32, T, 4 * C, device=dev), planes=1)
    w = ops.split16(torch.randn(C, 4 * C, device=dev) * 0.03, planes=1)
    resid = torch.randn(32, T, C, device=dev)
    rm = torch.ones(32 * T, device=dev)
    fn = lambda: ops.linear(x32, w, ops.f32, rowmul=rm, resid=resid, resid_masked=True)
elif which == "heads64":  # the dominant GEMM of bench.py's default step (64 clips): head tower conv over the (64, 2056, 1024) pyramid
    xh = ops.split16(torch.randn(64, 2056, C, device=dev))
    w3 = ops.split16(torch.randn(3, C, C, device=dev) * 0.03)
    rm = torch.ones(64, 2056, device=dev)
    fn = lambda: ops.conv3(xh, w3, ops.f32, rowmul=rm, flat=True)
elif which == "heads":   # the dominant GEMM of bench.py's default step: head tower conv over the (32, 2056, 1024) pyramid
    xh = ops.split16(torch.randn(32, 2056, C, device=dev))
    w3 = ops.split16(torch.randn(3, C, C, device=dev) * 0.03)
    rm = torch.ones(32, 2056, device=dev)
    fn = lambda: ops.conv3(xh, w3, ops.f32, rowmul=rm, flat=True)
elif which == "xl":      # fused XLNet relative attention at 32 clips
    Bx = 32
    mk = lambda *s_: ops.split16(torch.randn(*s_, device=dev), planes=1)
    qw, qr, kk, vv, kr = mk(Bx, T, C), mk(Bx, T, C), mk(Bx, T, C), mk(Bx, T, C), mk(2 * T, C)
    msk = torch.ones(Bx, T, device=dev)
    fn = lambda: ops.xl_attention(qw, qr, kk, vv, kr, msk, H, 0.125)
elif which == "wgrad":   # weight gradient dW = dZ^T X, both operands MN-major, split-K 2 (training step, 16 clips)
    from vilco_b200 import backward as BW
    dz = ops.split16(torch.randn(16 * T, C, device=dev))
    xa = ops.split16(torch.randn(16 * T, C, device=dev))
    fn = lambda: BW.wgrad(dz, xa, C, C, 16 * T)
else:
    w3 = ops.split16(torch.randn(3, C, C, device=dev) * 0.03)
    rm = torch.ones(B, T, device=dev)
    fn = lambda: ops.conv3(x, w3, ops.f32, rowmul=rm)
for _ in range(4):
    fn()
torch.cuda.synchronize()
print("ok")
