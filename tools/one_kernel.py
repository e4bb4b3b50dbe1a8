"""Single non-GEMM kernels for ncu captures: python tools/one_kernel.py {attn|ln|ln_bwd|resid_bwd|softmax|to_planes}"""
import ctypes as CT
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vilco_b200 import backward as BW, lib as L, ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "attn"
dev = "cuda"
R, C = 32768, 1024
if which == "attn":      # fused masked attention, 32 clips x 16 heads x (1024 x 1024), bf16x3
    B, H, T = 32, 16, 1024
    q, k, v = (ops.split16(torch.randn(B, T, C, device=dev)) for _ in range(3))
    km = torch.ones(B, T, device=dev)
    fn = lambda: ops.attention(q, k, v, km, H, 0.125)
elif which == "ln":
    x, w, b = torch.randn(R, C, device=dev), torch.randn(C, device=dev), torch.randn(C, device=dev)
    fn = lambda: ops.layernorm(x, w, b)
elif which == "ln_bwd":
    x, dy, w = torch.randn(R, C, device=dev), torch.randn(R, C, device=dev), torch.randn(C, device=dev)
    dw, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    fn = lambda: BW.layernorm_bwd(dy, x, w, dw_out=dw, db_out=db)
elif which == "resid_bwd":
    g, y = torch.randn(R, C, device=dev), torch.randn(R, C, device=dev)
    rm, sc, bi = torch.ones(R, device=dev), torch.randn(C, device=dev), torch.randn(C, device=dev)
    dres, dz = torch.empty_like(g), ops.empty16(R, C, device=dev)
    dbb, dss = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    fn = lambda: L.check(L.lib().vilco_resid_branch_bwd(ops._p(g), ops._p(rm), ops._p(y), ops._p(bi), ops._p(sc), ops._p(rm),
                                                       ops._p(dres), ops._p(dz), ops._i64(ops.lo(dz)), ops._p(dbb), ops._p(dss),
                                                       R, C, CT.c_float(0.1), CT.c_uint64(5), L.stream_ptr()))
elif which == "softmax":  # XLNet relative-attention softmax, 8 clips
    B, H, T = 8, 16, 1024
    ac, bd = torch.randn(B, H, T, T, device=dev), torch.randn(B, H, T, 2 * T, device=dev)
    km = torch.ones(B, T, device=dev)
    fn = lambda: ops.softmax_rows(ac, km, mode=1, BD=bd, scale=0.125)
else:
    x = torch.randn(R, C, device=dev)
    fn = lambda: BW.to_planes(x)
for _ in range(4):
    fn()
torch.cuda.synchronize()
print("ok")
