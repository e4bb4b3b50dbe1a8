"""GPU probe: compare vilco_gemm (tcgen05 and SIMT) against a torch fp32 matmul of the same bf16 operands.
Run on the GPU box:  python tools/gemm_probe.py   (prints one line per case; exits non-zero on mismatch)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vilco_b200 import lib as L  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
bad = 0


def report(name, got, ref, tol=2e-2):
    global bad
    got = got.float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-9
    ok = err <= tol * scale and torch.isfinite(got).all().item()
    bad += 0 if ok else 1
    print(f"{'OK ' if ok else 'BAD'} {name:58s} max_abs_err {err:.3e} ref_max {scale:.3e}", flush=True)


def case_linear(M, N, K, impl, out_dtype, epi=False):
    lda = (K + 7) // 8 * 8
    A = torch.zeros(M, lda, device=dev, dtype=torch.bfloat16)
    A[:, :K] = torch.randn(M, K, device=dev)
    W = torch.zeros(N, lda, device=dev, dtype=torch.bfloat16)
    W[:, :K] = torch.randn(N, K, device=dev) * 0.1
    ldd = (N + 7) // 8 * 8
    D = torch.zeros(M, ldd, device=dev, dtype=out_dtype)
    kw = {}
    ref = A[:, :K].float() @ W[:, :K].float().t()
    if epi:
        bias = torch.randn(N, device=dev)
        rowmul = (torch.rand(M, device=dev) > 0.3).float() * 1.5
        colscale = torch.randn(N, device=dev)
        resid = torch.randn(M, ldd, device=dev)
        kw = dict(bias=bias, rowmul=rowmul, colscale=colscale, resid=resid, resid_masked=True, act=L.ACT_GELU, alpha=0.5)
        ref = torch.nn.functional.gelu((ref * 0.5 + bias) * rowmul[:, None]) * colscale + resid[:, :N] * rowmul[:, None]
    L.gemm(A, W, D, M=M, N=N, K=K, a_rows=M, a_ld=lda, b_ld=lda, d_ld=ldd, impl=impl, **kw)
    torch.cuda.synchronize()
    report(f"linear M{M} N{N} K{K} impl{impl} {str(out_dtype)[6:]} epi{int(epi)}", D[:, :N], ref)


def case_conv3(Bsz, T, Cin, Cout, impl):
    x = torch.randn(Bsz, T, Cin, device=dev).bfloat16()
    w = (torch.randn(Cout, Cin, 3, device=dev) * 0.05).bfloat16()
    wt = w.permute(2, 0, 1).contiguous()  # [tap][Cout][Cin]
    D = torch.zeros(Bsz, T, Cout, device=dev, dtype=torch.float32)
    mask = (torch.arange(T, device=dev)[None, :] < torch.tensor([T, T * 2 // 3], device=dev)[:Bsz, None]).float()
    L.gemm(x, wt, D, M=T, N=Cout, K=Cin, a_rows=T, a_ld=Cin, a_s=(0, T * Cin), Z=(1, Bsz), taps=3, b_ld=Cin,
           b_s=(Cout * Cin, 0), d_ld=Cout, d_s=(0, T * Cout), rowmul=mask.contiguous(), rowmul_zs=T, impl=impl)
    torch.cuda.synchronize()
    ref = torch.nn.functional.conv1d(x.float().transpose(1, 2), w.float(), padding=1).transpose(1, 2) * mask[:, :, None]
    report(f"conv3 B{Bsz} T{T} Cin{Cin} Cout{Cout} impl{impl}", D, ref)


def case_attn(Bsz, H, Tq, Tk, impl):
    d = 64
    Cc = H * d
    q = torch.randn(Bsz, Tq, Cc, device=dev).bfloat16()
    k = torch.randn(Bsz, Tk, Cc, device=dev).bfloat16()
    v = torch.randn(Bsz, Tk, Cc, device=dev).bfloat16()
    S = torch.zeros(Bsz, H, Tq, Tk, device=dev, dtype=torch.float32)
    L.gemm(q, k, S, M=Tq, N=Tk, K=d, a_rows=Tq, a_ld=Cc, a_s=(d, Tq * Cc), Z=(H, Bsz), b_ld=Cc, b_s=(d, Tk * Cc),
           b_batched=True, d_ld=Tk, d_s=(Tq * Tk, H * Tq * Tk), alpha=0.125, impl=impl)
    torch.cuda.synchronize()
    qh = q.float().view(Bsz, Tq, H, d).permute(0, 2, 1, 3)
    kh = k.float().view(Bsz, Tk, H, d).permute(0, 2, 1, 3)
    vh = v.float().view(Bsz, Tk, H, d).permute(0, 2, 1, 3)
    ref = qh @ kh.transpose(-1, -2) * 0.125
    report(f"QK^T B{Bsz} H{H} Tq{Tq} Tk{Tk} impl{impl}", S, ref)
    ldp = (Tk + 7) // 8 * 8
    P = torch.zeros(Bsz, H, Tq, ldp, device=dev, dtype=torch.bfloat16)
    P[..., :Tk] = torch.softmax(ref, -1)
    O = torch.zeros(Bsz, Tq, Cc, device=dev, dtype=torch.bfloat16)
    L.gemm(P, v, O, M=Tq, N=d, K=Tk, a_rows=Tq, a_ld=ldp, a_s=(Tq * ldp, H * Tq * ldp), Z=(H, Bsz), b_ld=Cc,
           b_s=(d, Tk * Cc), b_batched=True, b_major=1, d_ld=Cc, d_s=(d, Tq * Cc), impl=impl)
    torch.cuda.synchronize()
    refo = (P[..., :Tk].float() @ vh).permute(0, 2, 1, 3).reshape(Bsz, Tq, Cc)
    report(f"P@V  B{Bsz} H{H} Tq{Tq} Tk{Tk} impl{impl}", O, refo)


def split(x):
    hi = x.bfloat16()
    return torch.stack([hi, (x - hi.float()).bfloat16()]).contiguous()


def case_split(M, N, K, impl):
    """bf16x3: operands as (hi, lo) planes; result must match the fp32 matmul to ~1e-5."""
    A32 = torch.randn(M, K, device=dev)
    W32 = torch.randn(N, K, device=dev) * 0.1
    A, W = split(A32), split(W32)
    D = torch.zeros(2, M, N, device=dev, dtype=torch.bfloat16)
    L.gemm(A, W, D, M=M, N=N, K=K, a_rows=M, a_ld=K, b_ld=K, d_ld=N, impl=impl, a_lo=A.stride(0), b_lo=W.stride(0),
           d_lo=D.stride(0))
    D32 = torch.zeros(M, N, device=dev, dtype=torch.float32)
    L.gemm(A, W, D32, M=M, N=N, K=K, a_rows=M, a_ld=K, b_ld=K, d_ld=N, impl=impl, a_lo=A.stride(0), b_lo=W.stride(0))
    torch.cuda.synchronize()
    ref = (A32.double() @ W32.double().t()).float()
    report(f"split f32-out M{M} N{N} K{K} impl{impl}", D32, ref, tol=2e-5)
    report(f"split hi+lo-out M{M} N{N} K{K} impl{impl}", D.float().sum(0), ref, tol=3e-5)


if __name__ == "__main__":
    impls = [int(a) for a in sys.argv[1:]] or [1, 0]
    for impl in impls:
        case_linear(256, 128, 128, impl, torch.float32)
        case_linear(128, 128, 64, impl, torch.bfloat16)
        case_linear(2048, 1024, 1024, impl, torch.bfloat16)
        case_linear(200, 22, 1024, impl, torch.float32)
        case_linear(57, 2, 104, impl, torch.float32)
        case_linear(300, 200, 100, impl, torch.float32, epi=True)
        case_linear(384, 256, 4096, impl, torch.bfloat16, epi=True)
        case_conv3(2, 256, 128, 128, impl)
        case_conv3(2, 100, 64, 40, impl)
        case_attn(2, 2, 128, 128, impl)
        case_attn(1, 16, 512, 57, impl)
        case_attn(2, 4, 1024, 1024, impl)
        case_split(256, 128, 512, impl)
        case_split(2048, 1024, 1024, impl)
        case_split(100, 22, 200, impl)
    print("launches", L.launch_count())
    sys.exit(1 if bad else 0)
