"""GPU probe of the unified tcgen05 GEMM: operand formats (fp16 / bf16, mixed), 1 or 2 planes per operand, one CTA or a CTA
pair (cta_group::2) per tile, k=3 taps, ragged M / N, every output type — against torch float64 matmuls of the same rounded
operands — then CUDA-event timings of the Moment-Query shapes at 32 clips.

    python tools/gemm2_probe.py            # correctness + timing        python tools/gemm2_probe.py check | time
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vilco_b200 import lib as L  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
bad = 0
F16, BF16 = torch.float16, torch.bfloat16


def planes(x, dtype, n):
    hi = x.to(dtype)
    if n == 1:
        return hi.unsqueeze(0).contiguous()
    return torch.stack([hi, (x - hi.float()).to(dtype)]).contiguous()


def lo(t):
    return t.stride(0) if t.shape[0] == 2 else 0


def report(name, got, ref, tol):
    global bad
    got = got.double()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-9
    ok = err <= tol * scale and torch.isfinite(got).all().item()
    bad += 0 if ok else 1
    print(f"{'OK ' if ok else 'BAD'} {name:78s} rel_err {err / scale:.2e} (tol {tol:.0e})", flush=True)


def case(M, N, K, fa, fb, pa, pb, impl, out=torch.float32, taps=1, epi=False, dplanes=1):
    """D = A W^T (taps: k=3 conv over rows) with operands rounded to their planes; reference = float64 on the SAME rounded
    operands, so the tolerance only covers fp32 accumulation (+ the dropped lo*lo term and the output rounding)."""
    A32 = torch.randn(M, K, device=dev)
    W32 = torch.randn(taps, N, K, device=dev) * 0.1
    A, W = planes(A32, fa, pa), planes(W32, fb, pb)
    Ar, Wr = A.double().sum(0), W.double().sum(0)
    if taps == 1:
        ref = Ar @ Wr[0].t()
    else:
        Ap = torch.nn.functional.pad(Ar, (0, 0, 1, 1))
        ref = sum(Ap[t:t + M] @ Wr[t].t() for t in range(3))
    D = torch.zeros(dplanes if out != torch.float32 else 1, M, N, device=dev, dtype=out)
    kw = {}
    if epi:
        bias = torch.randn(N, device=dev)
        rowmul = (torch.rand(M, device=dev) > 0.3).float() * 1.5
        colscale = torch.randn(N, device=dev)
        kw = dict(bias=bias, rowmul=rowmul, colscale=colscale, act=L.ACT_RELU, alpha=0.5)
        ref = torch.relu((ref * 0.5 + bias.double()) * rowmul.double()[:, None]) * colscale.double()
        if out == torch.float32:
            resid = torch.randn(M, N, device=dev)
            kw.update(resid=resid, resid_masked=True)
            ref = ref + resid.double() * rowmul.double()[:, None]
    L.gemm(A, W, D[0] if out == torch.float32 else D, M=M, N=N, K=K, a_rows=M, a_ld=K, b_ld=K, b_s=(N * K, 0), d_ld=N, taps=taps,
           impl=impl, a_lo=lo(A), b_lo=lo(W), d_lo=lo(D) if out != torch.float32 else 0, **kw)
    torch.cuda.synchronize()
    got = D.double().sum(0)
    # dropped lo*lo term: 2^-2p of the product magnitude; output rounding when the result is a single 16-bit plane
    eps = {F16: 2.0 ** -11, BF16: 2.0 ** -8}
    tol = (3e-6 if taps == 1 else 1e-5) + (eps[fa] * eps[fb] * 4 if pa == 2 and pb == 2 else 0)   # fp32 accumulation over K * taps
    if out != torch.float32:
        tol += eps[out] if dplanes == 1 else eps[out] ** 2 * 8
    name = (f"M{M} N{N} K{K} taps{taps} A:{str(fa)[6:]}x{pa} B:{str(fb)[6:]}x{pb} impl{impl} out:{str(out)[6:]}x{dplanes} epi{int(epi)}")
    report(name, got, ref, tol)


def case_major(M, N, K, a_major, b_major, impl, pa=2, pb=1, Z=1):
    """gradient-GEMM operand layouts: A stored (K, M) when a_major, B stored (K, N) when b_major (MN-major operands)"""
    A32 = torch.randn(Z, M, K, device=dev)
    B32 = torch.randn(Z, N, K, device=dev) * 0.1
    A = planes(A32.transpose(1, 2).contiguous() if a_major else A32, F16, pa)      # (P, Z, K, M) or (P, Z, M, K)
    Bt = planes(B32.transpose(1, 2).contiguous() if b_major else B32, F16, pb)
    ref = torch.einsum("zmk,znk->zmn", (A.double().sum(0).transpose(1, 2) if a_major else A.double().sum(0)),
                       (Bt.double().sum(0).transpose(1, 2) if b_major else Bt.double().sum(0)))
    D = torch.zeros(Z, M, N, device=dev)
    L.gemm(A, Bt, D, M=M, N=N, K=K, a_rows=M, a_ld=M if a_major else K, a_major=a_major, a_s=(M * K, 0), Z=(Z, 1),
           b_ld=N if b_major else K, b_s=(N * K, 0), b_batched=True, b_major=b_major, d_ld=N, d_s=(M * N, 0), impl=impl,
           a_lo=lo(A), b_lo=lo(Bt))
    torch.cuda.synchronize()
    report(f"M{M} N{N} K{K} Z{Z} a_major{a_major} b_major{b_major} planes {pa}/{pb} impl{impl}", D.double(), ref, 1e-5 if K < 2048 else 3e-5)


def check():
    for impl in (2, 3):
        case_major(2048, 1024, 1024, 0, 1, impl)              # dgrad: dZ K-major, W as MN-major B
        case_major(1024, 1024, 4096, 1, 1, impl, Z=4)         # wgrad: both MN-major, split-K batches
        case_major(1000, 520, 328, 1, 1, impl)                # ragged
        case_major(4096, 1024, 1024, 0, 1, impl, pa=1, pb=2)
    for impl in (1, 2, 3):   # SIMT, one CTA per tile, CTA pair per tile
        for fa, fb in ((F16, F16), (BF16, BF16)) + (((BF16, F16),) if impl == 1 else ()):   # the tensor core cannot mix formats
            for pa, pb in ((1, 1), (2, 2), (2, 1), (1, 2)):
                case(512, 256, 256, fa, fb, pa, pb, impl)
        case(1000, 520, 328, F16, F16, 1, 1, impl, epi=True)                        # ragged M / N / K tails + epilogue
        case(2056 * 2, 1024, 1024, F16, F16, 1, 1, impl, taps=3, epi=True)           # head-conv shape (flat pyramid rows)
        case(1030, 256, 512, F16, F16, 2, 2, impl, taps=3)
        case(2048, 1024, 1024, F16, F16, 1, 1, impl, out=F16, epi=True)              # 16-bit output, one plane
        case(2048, 1024, 1024, F16, F16, 2, 2, impl, out=F16, dplanes=2, epi=True)   # 16-bit output, two planes
        case(2048, 3072, 1024, BF16, BF16, 2, 2, impl, out=BF16, dplanes=2)
        case(777, 110, 1024, F16, F16, 1, 1, impl, taps=3, epi=True)                 # N % 8 != 0 (K = 110 classes)
        case(4096, 4096, 1024, F16, F16, 1, 1, impl, out=F16, epi=True)
    print("launches", L.launch_count(), "bad", bad)


def timeit(fn, flops, name, n=30):
    """whole-sequence time of n back-to-back launches (what a step sees) and the per-launch durations inside it"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    us = ev[0].elapsed_time(ev[n]) * 1e3 / n
    each = sorted(ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(n))
    print(f"{name:70s} {us:9.1f} us  {flops / us / 1e6:8.1f} TFLOP/s   per launch min {each[0]:.1f} med {each[n // 2]:.1f} max {each[-1]:.1f}",
          flush=True)


def timing():
    Bc = int(os.environ.get("CLIPS", 32))
    shapes = [("heads conv3 flat", Bc * 2056, 1024, 1024, 3, torch.float32),
              ("C x C -> planes", Bc * 1024, 1024, 1024, 1, F16),
              ("C x C -> f32 + resid", Bc * 1024, 1024, 1024, 1, torch.float32),
              ("mlp up (gelu) -> planes", Bc * 1024, 4096, 1024, 1, F16),
              ("mlp down -> f32", Bc * 1024, 1024, 4096, 1, torch.float32),
              ("proj 4096 -> planes", Bc * 1024, 1024, 4096, 1, F16),
              ("branch level 2 C x C", Bc * 256, 1024, 1024, 1, torch.float32),
              ("branch level 4 C x C", Bc * 64, 1024, 1024, 1, torch.float32)]
    for name, M, N, K, taps, out in shapes:
        for (pa, pb, impls) in ((1, 1, (2, 3)), (2, 2, (2, 3))):
            A = planes(torch.randn(M, K, device=dev), F16, pa)
            W = planes(torch.randn(taps, N, K, device=dev) * 0.05, F16, pb)
            D = torch.empty(M, N, device=dev, dtype=out) if out == torch.float32 else torch.empty(1, M, N, device=dev, dtype=out)
            resid = torch.randn(M, N, device=dev) if out == torch.float32 else None
            bias = torch.randn(N, device=dev)
            for impl in impls + (0,):
                fn = lambda: L.gemm(A, W, D, M=M, N=N, K=K, a_rows=M, a_ld=K, b_ld=K, b_s=(N * K, 0), d_ld=N, taps=taps, impl=impl,  # noqa: E731
                                    a_lo=lo(A), b_lo=lo(W), bias=bias, resid=resid, act=L.ACT_GELU if out != torch.float32 else L.ACT_NONE)
                timeit(fn, 2.0 * M * N * K * taps, f"{name} M{M} N{N} K{K} t{taps} planes {pa}/{pb} impl{impl}")
            del A, W, D


def bwd_timing():
    """gradient-GEMM shapes (MN-major operands): one CTA per 128-wide tile (impl 2) vs CTA pairs (impl 3) vs automatic"""
    R = 32768
    def dgrad(M, N, K, impl):
        A = planes(torch.randn(M, K, device=dev), F16, 1)
        Bt = planes(torch.randn(K, N, device=dev) * 0.05, F16, 1)          # W (K rows = out features, N cols): MN-major B
        D = torch.empty(M, N, device=dev)
        return lambda: L.gemm(A, Bt, D, M=M, N=N, K=K, a_rows=M, a_ld=K, b_ld=N, b_major=1, d_ld=N, impl=impl, a_lo=0, b_lo=0)
    def wgradf(Mo, No, Rr, S, impl):
        dz = planes(torch.randn(Rr, Mo, device=dev), F16, 1)
        x = planes(torch.randn(Rr, No, device=dev), F16, 1)
        part = torch.empty(S, Mo, No, device=dev)
        ch = Rr // S
        return lambda: L.gemm(dz, x, part, M=Mo, N=No, K=ch, a_rows=Mo, a_ld=Mo, a_major=1, a_s=(ch * Mo, 0), Z=(S, 1), b_ld=No,
                              b_s=(ch * No, 0), b_batched=True, b_major=1, d_ld=No, d_s=(Mo * No, 0), impl=impl, a_lo=0, b_lo=0)
    for M in (32768, 16384, 8192, 4096, 2048, 1024):
        for impl in (2, 3, 0):
            timeit(dgrad(M, 1024, 1024, impl), 2.0 * M * 1024 * 1024, f"dgrad M{M} N1024 K1024 impl{impl}")
    for impl in (2, 3, 0):
        timeit(dgrad(R, 4096, 1024, impl), 2.0 * R * 4096 * 1024, f"dgrad M{R} N4096 K1024 impl{impl}")
        timeit(dgrad(R, 1024, 4096, impl), 2.0 * R * 4096 * 1024, f"dgrad M{R} N1024 K4096 impl{impl}")
    for (Mo, No, Rr) in ((1024, 1024, R), (4096, 1024, R), (1024, 4096, R), (1024, 1024, 8192), (1024, 1024, 2048)):
        for S in (1, 2, 4, 8, 16):
            if Rr // S < 512:
                continue
            for impl in (2, 3):
                timeit(wgradf(Mo, No, Rr, S, impl), 2.0 * Mo * No * Rr, f"wgrad {Mo}x{No} R{Rr} S{S} impl{impl}")


def sweep():
    """per-tile cost model: time(K) at fixed N and time(N) at fixed K, plain epilogue, one fp16 plane"""
    M = 32768
    for (N, K, out, extra) in [(1024, 256, F16, ""), (1024, 512, F16, ""), (1024, 1024, F16, ""), (1024, 2048, F16, ""), (1024, 4096, F16, ""),
                               (2048, 1024, F16, ""), (4096, 1024, F16, ""), (1024, 1024, torch.float32, ""), (4096, 1024, torch.float32, ""),
                               (1024, 1024, F16, "bias+gelu"), (1024, 1024, torch.float32, "resid")]:
        A = planes(torch.randn(M, K, device=dev), F16, 1)
        W = planes(torch.randn(1, N, K, device=dev) * 0.05, F16, 1)
        D = torch.empty(M, N, device=dev, dtype=out) if out == torch.float32 else torch.empty(1, M, N, device=dev, dtype=out)
        kw = {}
        if extra == "bias+gelu":
            kw = dict(bias=torch.randn(N, device=dev), act=L.ACT_GELU)
        if extra == "resid":
            kw = dict(resid=torch.randn(M, N, device=dev))
        fn = lambda: L.gemm(A, W, D, M=M, N=N, K=K, a_rows=M, a_ld=K, b_ld=K, b_s=(N * K, 0), d_ld=N, a_lo=0, b_lo=0, **kw)  # noqa: E731
        tiles = (M // 256) * (N // 256)
        waves = -(-tiles // 74)
        timeit(fn, 2.0 * M * N * K, f"sweep M{M} N{N} K{K} out {str(out)[6:]} {extra} ({waves} waves)")
        del A, W, D


if __name__ == "__main__":
    what = sys.argv[1:] or ["check", "time"]
    if "check" in what:
        check()
    if "time" in what:
        timing()
    if "bwd" in what:
        bwd_timing()
    if "sweep" in what:
        sweep()
    sys.exit(1 if bad else 0)
