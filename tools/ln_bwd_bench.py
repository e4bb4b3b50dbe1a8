"""GPU micro-benchmark of the bandwidth-bound backward kernels: python tools/ln_bwd_bench.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vilco_b200 import backward as BW, ops

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

print("peak HBM (MEASURED_PEAKS.json): see repo root; fp32 token-major (16384, 1024) tensors unless noted")
R, C = 16384, 1024
x, dy = torch.randn(R, C, device="cuda"), torch.randn(R, C, device="cuda")
w = torch.randn(C, device="cuda")
dw, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
us = timeit(lambda: BW.layernorm_bwd(dy, x, w, dw_out=dw, db_out=db))
print(f"layernorm_bwd {R}x{C}: {us:.1f} us  {3 * R * C * 4 / us / 1e6:.2f} TB/s (3 fp32 passes)")
us = timeit(lambda: BW.to_planes(dy))
print(f"to_planes     {R}x{C}: {us:.1f} us  {R * C * 8 / us / 1e6:.2f} TB/s (4 B in, 2x2 B out)")
us = timeit(lambda: BW.colsum(dy, out=dw))
print(f"colsum        {R}x{C}: {us:.1f} us  {R * C * 4 / us / 1e6:.2f} TB/s")
us = timeit(lambda: BW.gelu_bwd(dy, x))
print(f"gelu_bwd      {R}x{C}: {us:.1f} us  {3 * R * C * 4 / us / 1e6:.2f} TB/s")
us = timeit(lambda: ops.ew(1, x, out16=True))
print(f"gelu fwd+planes {R}x{C}: {us:.1f} us  {R * C * 12 / us / 1e6:.2f} TB/s")
us = timeit(lambda: ops.dropout(x, 0.1, 7))
print(f"dropout       {R}x{C}: {us:.1f} us  {R * C * 8 / us / 1e6:.2f} TB/s")
lw, lb = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
us = timeit(lambda: ops.layernorm(x, lw, lb, out32=True, out16=True))
print(f"layernorm fwd (fp32 + planes) {R}x{C}: {us:.1f} us  {R * C * 12 / us / 1e6:.2f} TB/s (4 B in, 4 + 4 B out)")
res, y = torch.randn(R, C, device="cuda"), torch.randn(R, C, device="cuda")
import ctypes as CT
from vilco_b200 import lib as L
out = torch.empty_like(y)
rm = torch.ones(R, device="cuda")
us = timeit(lambda: L.check(L.lib().vilco_resid_branch_fwd(ops._p(res), ops._p(rm), ops._p(y), ops._p(lb), ops._p(lw), ops._p(rm), ops._p(out), ops._i64(R), C, CT.c_float(0.1), CT.c_uint64(5), L.stream_ptr())))
print(f"resid_branch_fwd (dropout) {R}x{C}: {us:.1f} us  {R * C * 12 / us / 1e6:.2f} TB/s (8 B in, 4 B out)")
dz = ops.empty16(R, C, device="cuda")
dbb, dss = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
us = timeit(lambda: L.check(L.lib().vilco_resid_branch_bwd(ops._p(dy), ops._p(rm), ops._p(y), ops._p(lb), ops._p(lw), ops._p(rm), ops._p(out), ops._p(dz), ops._i64(ops.lo(dz)), ops._p(dbb), ops._p(dss), R, C, CT.c_float(0.1), CT.c_uint64(5), L.stream_ptr())))
print(f"resid_branch_bwd (dropout) {R}x{C}: {us:.1f} us  {R * C * 16 / us / 1e6:.2f} TB/s (8 B in, 4 + 4 B out)")
B, H, T = 4, 16, 1024
P = torch.softmax(torch.randn(B, H, T, T, device="cuda"), -1)
P16 = ops.split16(P)
dP = torch.randn(B, H, T, T, device="cuda")
us = timeit(lambda: BW.softmax_bwd(dP, 0.125, P16=P16), 5)
print(f"softmax_bwd {B}x{H}x{T}x{T}: {us:.1f} us  {B * H * T * T * 12 / us / 1e6:.2f} TB/s (4 B P planes + 4 B dP in, 4 B planes out)")
