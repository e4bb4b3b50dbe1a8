"""Why a GEMM micro-benchmark depends on what ran before it: the same launch timed fresh, right after 4 s of dense tensor load
(sw_power_cap engages near 880 W and the launch takes ~25 % longer), after a pause, with new allocations.  python tools/power_state_probe.py"""
import os, sys, time, subprocess
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vilco_b200 import lib as L
dev = "cuda"
F16 = torch.float16
def mk(M, N, K):
    A = torch.randn(1, M, K, device=dev).to(F16); W = (torch.randn(1, 1, N, K, device=dev) * 0.05).to(F16)
    D = torch.empty(1, M, N, device=dev, dtype=F16)
    return A, W, D
def run(A, W, D, n=30):
    M, K = A.shape[1], A.shape[2]; N = W.shape[2]
    fn = lambda: L.gemm(A, W, D, M=M, N=N, K=K, a_rows=M, a_ld=K, b_ld=K, b_s=(N * K, 0), d_ld=N, a_lo=0, b_lo=0)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n
def smi():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_throttle_reasons.active", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
M, N, K = 32768, 1024, 1024
t = mk(M, N, K)
print("fresh               ", run(*t), smi())
print("again               ", run(*t), smi())
big = mk(65792, 1024, 4096)
t0 = time.time()
while time.time() - t0 < 4: run(*big, n=50)
print("after 4 s heavy     ", run(*t), smi())
print("again               ", run(*t), smi())
time.sleep(10)
print("after 10 s sleep    ", run(*t), smi())
t2 = mk(M, N, K)
print("new tensors         ", run(*t2), smi())
del big; torch.cuda.empty_cache()
t3 = mk(M, N, K)
print("new after empty     ", run(*t3), smi())
junk = [torch.empty(100 << 20, device=dev, dtype=torch.uint8) for _ in range(200)]
t4 = mk(M, N, K)
print("after 20 GB of junk ", run(*t4), smi())
print("orig tensors        ", run(*t), smi())
for n in (100, 300, 1000):
    print("n =", n, run(*t, n=n), smi())
