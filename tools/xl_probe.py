"""GPU probe of the fused XLNet relative attention kernel (csrc/xlattn.cu) against a float64 torch restatement on the same
rounded operands, and against the materialised chain (score GEMMs -> softmax_rows mode 1 -> P V); then timings at 32 clips."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vilco_b200 import ops  # noqa: E402

ops.set_precision(sys.argv[1] if len(sys.argv) > 1 else "mixed")
dev = "cuda"
torch.manual_seed(0)
bad = 0


def ref(qw, qr, k, v, kr, mask, H, scale):
    B, T, C = qw.shape
    d = C // H
    hd = lambda t: t.double().view(t.shape[0], -1, H, d).permute(0, 2, 1, 3)   # noqa: E731
    ac = hd(qw) @ hd(k).transpose(-1, -2)
    bdr = hd(qr) @ hd(kr[None]).transpose(-1, -2)                              # (B,H,T,2T)
    i = torch.arange(T, device=qw.device)
    idx = (T + i[None, :] - i[:, None])
    bd = bdr.gather(3, idx[None, None].expand(B, H, T, T))
    s = (ac + bd) * scale
    pad = (mask == 0)[:, None, None, :].expand(B, 1, T, T).clone()
    pad[:, :, i, i] = False
    s = s - 1e30 * pad
    return (torch.softmax(s, -1) @ hd(v)).permute(0, 2, 1, 3).reshape(B, T, C)


def case(B, T, H, lens, mag=1.0):
    global bad
    C = H * 64
    mk = lambda *s: ops.split16(torch.randn(*s, device=dev) * mag, planes=1)   # noqa: E731
    qw, qr, k, v = mk(B, T, C), mk(B, T, C), mk(B, T, C), mk(B, T, C)
    kr = mk(2 * T, C)
    mask = (torch.arange(T, device=dev)[None, :] < torch.tensor(lens, device=dev)[:, None]).float().contiguous()
    scale = 1.0 / 8
    out = ops.xl_attention(qw, qr, k, v, kr, mask, H, scale)
    torch.cuda.synchronize()
    r = ref(qw[0], qr[0], k[0], v[0], kr[0], mask, H, scale)
    err = float((out[0].double() - r).abs().max() / r.abs().max())
    # materialised chain on the same operands
    krb = kr.unsqueeze(1).expand(-1, B, 2 * T, C).contiguous()
    ac = ops.attn_scores(qw, k, H, 1.0)
    bd = ops.attn_scores(qr, krb, H, 1.0, band=(T, 2 * T))
    P = ops.softmax_rows(ac, mask, mode=1, BD=bd, scale=scale)
    o2 = ops.attn_pv(P, v, H, T)
    torch.cuda.synchronize()
    err2 = float((o2[0].double() - r).abs().max() / r.abs().max())
    ok = err < 2e-3 and torch.isfinite(out.float()).all().item()
    bad += 0 if ok else 1
    print(f"{'OK ' if ok else 'BAD'} B{B} T{T} H{H} lens{lens} mag{mag}: fused rel err {err:.2e}   materialised chain {err2:.2e}", flush=True)


def timeit(fn, name, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:50s} {e0.elapsed_time(e1) * 1e3 / n:9.1f} us", flush=True)


if __name__ == "__main__":
    case(1, 128, 1, [128])
    case(1, 256, 2, [200])
    case(2, 512, 4, [512, 37])
    case(2, 1024, 16, [1024, 700])
    case(2, 1024, 16, [1024, 700], mag=3.0)       # peaky rows: exercises the lazy rescale
    case(1, 2048, 2, [1500])
    if os.environ.get("XL_CHECK_ONLY"):
        sys.exit(1 if bad else 0)
    B, T, H = 32, 1024, 16
    C = H * 64
    mk = lambda *s: ops.split16(torch.randn(*s, device=dev), planes=1)   # noqa: E731
    qw, qr, k, v, kr = mk(B, T, C), mk(B, T, C), mk(B, T, C), mk(B, T, C), mk(2 * T, C)
    mask = torch.ones(B, T, device=dev)
    timeit(lambda: ops.xl_attention(qw, qr, k, v, kr, mask, H, 0.125), "fused xl attention B32 T1024 H16")
    krb = kr.unsqueeze(1).expand(-1, B, 2 * T, C).contiguous()

    def chain():
        ac = ops.attn_scores(qw, k, H, 1.0)
        bd = ops.attn_scores(qr, krb, H, 1.0, band=(T, 2 * T))
        P = ops.softmax_rows(ac, mask, mode=1, BD=bd, scale=0.125)
        return ops.attn_pv(P, v, H, T)
    timeit(chain, "materialised chain B32 T1024 H16")
    timeit(lambda: ops.attention(qw, k, v, mask, H, 0.125), "fused global attention (stem) B32 T1024 H16")
    sys.exit(1 if bad else 0)
