"""Tensor-level wrappers over the C ABI (include/vilco_b200.h).  Token-major layout: (B, T, C), C contiguous.

Every function here launches hand-written sm_100a kernels through ``libvilco_b200.so``; torch only owns the memory.

Operand convention: every 16-bit *operand* tensor carries a leading "plane" dimension of size 1 or 2 (x ~= hi + lo, lo =
round16(x - hi)); a GEMM issues hi*hi, + hi*lo when B has two planes, + lo*hi when A has two (fp32 accumulation).
Precision modes (`set_precision`, env VILCO_PRECISION):

  "mixed"   (default) fp16 planes.  One plane everywhere except the contractions the parity bar is sensitive to — the input
            projection, the embedding convolutions and the channel-attention qkv / core, whose rounding noise the d x d
            channel softmax amplifies (tools/precision_sweep.py) — which run on split operands.
            Measured / emulated error of logits and offsets at the full MQ config: ~4e-4 (bar: 1e-3).
  "fp16x3"  fp16 planes, every operand split: 3 MMAs per k-step, ~1e-6 relative error (exact mode)
  "bf16x3"  bf16 planes, every operand split (round 1's parity mode, ~1e-5)
  "fp16" / "bf16"  one plane everywhere (1.3e-3 / 1e-2 at the full config: above the bar, kept for comparison)

Gradient planes (the 16-bit operands the backward kernels emit) share the activation format — tcgen05 kind::f16 cannot mix
fp16 and bf16 operands in one MMA (measured: illegal instruction).  In the fp16 modes they are stored multiplied by
GRAD_SCALE = 2^10, which centres gradient magnitudes in the fp16 range (normal from 6e-8 / 2^10, saturating at 64), and every
GEMM that consumes one folds 1 / GRAD_SCALE into its alpha (`ginv()`).  Their plane count follows the mode (`grad_planes()`):
one plane in `mixed` / `fp16` / `bf16`, hi + lo in the exact modes; VILCO_BWD_PRECISION = split | single overrides it.
"""
import ctypes as C
import os

import torch

from . import lib as L
from .lib import ACT_GELU, ACT_NONE, ACT_RELU  # noqa: F401

bf16 = torch.bfloat16        # also the `out_dtype` marker meaning "operand planes in the activation format"
f16 = torch.float16
f32 = torch.float32

_MODES = {  # name -> (activation plane dtype, default planes, planes of the sensitive contractions, gradient planes)
    "mixed": (f16, 1, 2, 1), "fp16x3": (f16, 2, 2, 2), "fp16": (f16, 1, 1, 1), "bf16x3": (bf16, 2, 2, 2), "bf16": (bf16, 1, 1, 1),
}
_mode = None
ACT_DTYPE, PLANES, PLANES_HI, GRAD_PLANES = bf16, 2, 2, 2
GRAD_SCALE = 1.0
FUSED_ATTN = os.environ.get("VILCO_FUSED_ATTN", "1") == "1"  # 0: materialised QK^T -> softmax -> PV kernels


def set_precision(name):
    """Select the operand-format policy (see the module docstring).  Weights must be re-packed after a change (the model
    does so: its packed-weight cache is keyed on `precision()`)."""
    global _mode, ACT_DTYPE, PLANES, PLANES_HI, GRAD_PLANES, GRAD_SCALE
    assert name in _MODES, name
    _mode = name
    ACT_DTYPE, PLANES, PLANES_HI, GRAD_PLANES = _MODES[name]
    GRAD_SCALE = float(2 ** int(os.environ.get("VILCO_GRAD_SCALE_LOG2", "10"))) if ACT_DTYPE == f16 else 1.0
    if os.path.exists(L.LIB_PATH):   # the library holds the plane format / gradient scale the non-GEMM kernels use
        L.check(L.lib().vilco_set_plane_format(L.F16 if ACT_DTYPE == f16 else L.BF16), "vilco_set_plane_format")
        L.check(L.lib().vilco_set_grad_scale(C.c_float(GRAD_SCALE)), "vilco_set_grad_scale")


def precision():
    return _mode


class use_precision:
    """context manager: run a block in another operand-format policy (weights are re-packed by the models: their packed-weight
    caches are keyed on `precision()`)"""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        self.prev = _mode
        set_precision(self.name)

    def __exit__(self, *a):
        set_precision(self.prev)


set_precision(os.environ.get("VILCO_PRECISION", "mixed"))


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _i64(v):
    return C.c_int64(int(v))


# Planes of the gradient operands: "mode" (default) = what the operand mode says (grad_planes()), "split" = hi + lo everywhere,
# "single" = one plane everywhere; "bf16" (round 1) additionally makes the backward GEMMs read only the hi plane of every
# operand.  The forward pass, the losses and therefore every parity statement about outputs are unaffected by this knob.
BWD_PRECISION = os.environ.get("VILCO_BWD_PRECISION", "mode")
_single = False


class single_plane:
    """context: operand tensors are treated as single-plane (lo() == 0) for everything launched inside"""

    def __enter__(self):
        global _single
        self.prev, _single = _single, True

    def __exit__(self, *a):
        global _single
        _single = self.prev


def lo(t):
    """element offset of the lo plane of an operand tensor (0 when single-plane)."""
    return t.stride(0) if (t.shape[0] == 2 and not _single) else 0


def grad_planes():
    """planes of a gradient operand: the mode's (one fp16 plane x 2^10 in `mixed`, hi + lo in the exact modes) unless
    VILCO_BWD_PRECISION forces `split` (2) or `single` / `bf16` (1)"""
    if BWD_PRECISION == "split":
        return 2
    if BWD_PRECISION in ("single", "bf16"):
        return 1
    return GRAD_PLANES


def ginv():
    """factor that undoes the scale of a gradient-plane operand: multiply the alpha of every GEMM reading one by it"""
    return 1.0 / GRAD_SCALE


def empty16(*shape, device="cuda", planes=None, grad=False):
    """uninitialised operand tensor (planes, *shape); grad: it will hold a (scaled) gradient"""
    if grad:
        return torch.empty(grad_planes() if planes is None else planes, *shape, device=device, dtype=ACT_DTYPE)
    return torch.empty(PLANES if planes is None else planes, *shape, device=device, dtype=ACT_DTYPE)


def zeros16(*shape, device="cuda", planes=None, grad=False):
    if grad:
        return torch.zeros(grad_planes() if planes is None else planes, *shape, device=device, dtype=ACT_DTYPE)
    return torch.zeros(PLANES if planes is None else planes, *shape, device=device, dtype=ACT_DTYPE)


def split16(x, planes=None, dtype=None):
    """fp32 tensor -> (planes, ...) operand (used for weights / constants at pack time)."""
    dtype = ACT_DTYPE if dtype is None else dtype
    planes = PLANES if planes is None else planes
    xc = x.clamp(-65504.0, 65504.0) if dtype == f16 else x
    hi = xc.to(dtype)
    if planes == 1:
        return hi.unsqueeze(0).contiguous()
    return torch.stack([hi, (xc - hi.float()).to(dtype)]).contiguous()


def merge16(t):
    """(planes, ...) operand -> fp32 value (tests / debugging)."""
    return t.float().sum(0)


def linear(x, w, out_dtype=bf16, bias=None, rowmul=None, act=ACT_NONE, colscale=None, resid=None, resid_masked=False,
           alpha=1.0, out=None, planes=None):
    """y[r, n] = epi(sum_k x[r, k] w[n, k]).  x (NP, ..., K) bf16 operand, w (NP, N, K).  rowmul (rows,) fp32,
    resid (..., N) fp32.  Returns an operand tensor (NP, ..., N) for bf16 output, a plain fp32 tensor otherwise."""
    K = x.shape[-1]
    N = w.shape[1]
    assert w.shape[2] == K and K % 8 == 0
    rows = x[0].numel() // K
    if out is None:
        out = empty16(*x.shape[1:-1], N, device=x.device, planes=planes) if out_dtype == bf16 else \
            torch.empty(*x.shape[1:-1], N, device=x.device, dtype=f32)
    L.gemm(x, w, out, M=rows, N=N, K=K, a_rows=rows, a_ld=K, b_ld=K, d_ld=N, a_lo=lo(x), b_lo=lo(w),
           d_lo=lo(out) if out.dtype != f32 else 0,
           bias=bias, rowmul=rowmul, act=act, colscale=colscale, resid=resid, resid_masked=resid_masked, alpha=alpha)
    return out


def conv3(x, w3, out_dtype=bf16, bias=None, rowmul=None, act=ACT_NONE, flat=False):
    """k=3 stride-1 zero-padded conv over time.  x (NP, B, T, Cin), w3 (NP, 3, Cout, Cin) tap-major,
    rowmul (B, T) fp32 -> (B, T, Cout) fp32 or operand (NP, B, T, Cout)."""
    _, B, T, Cin = x.shape
    Cout = w3.shape[2]
    assert w3.shape[1] == 3 and w3.shape[3] == Cin and Cin % 8 == 0
    out = empty16(B, T, Cout, device=x.device) if out_dtype == bf16 else torch.empty(B, T, Cout, device=x.device, dtype=f32)
    if flat:
        # the caller guarantees that the last row of every clip is zero and masked (the pyramid's gap rows): the batch is then
        # one long sequence, so tiles do not stop at clip boundaries (2056-row clips would waste 1/9 of every 256-row tile)
        L.gemm(x, w3, out, M=B * T, N=Cout, K=Cin, a_rows=B * T, a_ld=Cin, taps=3, b_ld=Cin, b_s=(Cout * Cin, 0), d_ld=Cout,
               a_lo=lo(x), b_lo=lo(w3), d_lo=lo(out) if out.dtype != f32 else 0, bias=bias, rowmul=rowmul, act=act)
        return out
    L.gemm(x, w3, out, M=T, N=Cout, K=Cin, a_rows=T, a_ld=Cin, a_s=(0, T * Cin), Z=(1, B), taps=3, b_ld=Cin,
           b_s=(Cout * Cin, 0), d_ld=Cout, d_s=(0, T * Cout), a_lo=lo(x), b_lo=lo(w3),
           d_lo=lo(out) if out.dtype != f32 else 0, bias=bias, rowmul=rowmul, rowmul_zs=T, act=act)
    return out


def attn_scores(q, k, H, alpha, band=(0, 0)):
    """S[b,h] = alpha * q_h k_h^T.  q (NP,B,Tq,C), k (NP,B,Tk,C) -> (B,H,Tq,Tk) fp32.  band = (lo, hi): only the
    entries with lo <= i + j < hi are computed (the rest of the buffer is left untouched)."""
    _, B, Tq, Cc = q.shape
    Tk = k.shape[2]
    d = Cc // H
    out = torch.empty(B, H, Tq, Tk, device=q.device, dtype=f32)
    L.gemm(q, k, out, M=Tq, N=Tk, K=d, a_rows=Tq, a_ld=Cc, a_s=(d, Tq * Cc), Z=(H, B), b_ld=Cc, b_s=(d, Tk * Cc),
           b_batched=True, d_ld=Tk, d_s=(Tq * Tk, H * Tq * Tk), alpha=alpha, a_lo=lo(q), b_lo=lo(k), band=band)
    return out


def attn_pv(P, v, H, Tk, out32=False, a_trans=False, M=None, alpha=1.0):
    """O[b, :, h*d:(h+1)*d] = P[b,h] v_h.  P (NP,B,H,Tq,ldp), v (NP,B,Tk,C) -> operand (NP,B,Tq,C) (fp32 (B,Tq,C) if out32).
    a_trans: use P[b,h]^T instead — P is (NP,B,H,Tk,ldp) with M <= ldp valid columns, read as an MN-major A operand (no
    transposed copy), result (B,M,C)."""
    _, B, _, R, ldp = P.shape
    Cc = v.shape[3]
    d = Cc // H
    assert d % 8 == 0 and d <= 128, "attn_pv: head dim must be a multiple of 8, <= 128"
    Tq = (M if M is not None else ldp) if a_trans else R
    out = torch.empty(B, Tq, Cc, device=v.device, dtype=f32) if out32 else empty16(B, Tq, Cc, device=v.device)
    L.gemm(P, v, out, M=Tq, N=d, K=Tk, a_rows=Tq, a_ld=ldp, a_s=(R * ldp, H * R * ldp), Z=(H, B), b_ld=Cc,
           b_s=(d, Tk * Cc), b_batched=True, b_major=1, d_ld=Cc, d_s=(d, Tq * Cc), a_lo=lo(P), b_lo=lo(v),
           d_lo=0 if out32 else lo(out), a_major=1 if a_trans else 0, alpha=alpha)
    return out


def attention(q, k, v, kmask, H, scale):
    """Fused masked attention core: q (NP,B,Tq,C), k/v (NP,B,Tk,C) operands, kmask (B,Tk) fp32 or None
    -> operand (NP,B,Tq,C).  Scores never leave TMEM."""
    _, B, Tq, Cc = q.shape
    Tk = k.shape[2]
    assert lo(k) == lo(v)
    out = empty16(B, Tq, Cc, device=q.device)
    L.check(L.lib().vilco_attention(_p(q), _i64(lo(q)), _p(k), _p(v), _i64(lo(k)), _p(kmask), _p(out), _i64(lo(out)), B, H,
                                    Tq, Tk, Cc, C.c_float(scale), L.stream_ptr()), "vilco_attention")
    return out


def self_attention(q, k, v, kmask, H, scale, want_lse=False):
    """Single-pass masked self-attention (csrc/xlattn.cu, REL = false): q / k / v (1,B,T,C) -> operand (1,B,T,C)
    [, lse2 (B,H,T) fp32: the row log-sum-exp (base 2) of the scaled masked scores, for the fused gradient epilogues]."""
    _, B, T, Cc = q.shape
    assert q.shape[0] == 1 and k.shape == q.shape and v.shape == q.shape
    out = empty16(B, T, Cc, device=q.device, planes=1)
    lse = torch.empty(B, H, T, device=q.device, dtype=f32) if want_lse else None
    L.check(L.lib().vilco_self_attention_lse(_p(q), _p(k), _p(v), _p(kmask), _p(out), _p(lse), B, H, T, Cc, C.c_float(scale),
                                             L.stream_ptr()), "vilco_self_attention_lse")
    return (out, lse) if want_lse else out


LOG2E = 1.4426950408889634


def attn_probs_from_lse(q, k, H, scale, lse2, kmask):
    """P[b,h] = exp2(scale * log2(e) * q_h k_h^T - lse2[b,h,:,None]) * kmask[b,None,:] as ONE 16-bit plane (1,B,H,Tq,Tk): the
    QK^T GEMM with the softmax recompute fused into its epilogue (VilcoGemm.rowsub / ACT_EXP2 / colscale)."""
    _, B, Tq, Cc = q.shape
    Tk = k.shape[2]
    d = Cc // H
    out = empty16(B, H, Tq, Tk, device=q.device, planes=1)
    L.gemm(q, k, out, M=Tq, N=Tk, K=d, a_rows=Tq, a_ld=Cc, a_s=(d, Tq * Cc), Z=(H, B), b_ld=Cc, b_s=(d, Tk * Cc), b_batched=True,
           d_ld=Tk, d_s=(Tq * Tk, H * Tq * Tk), alpha=scale * LOG2E, a_lo=lo(q), b_lo=lo(k), act=L.ACT_EXP2,
           rowsub=lse2, rowsub_s=(Tq, H * Tq), colscale=kmask, colscale_zs=Tk if kmask is not None else 0)
    return out


def attn_dscores_fused(dO16, v, H, P16, delta_scaled, alpha):
    """dS[b,h] = P[b,h] * (alpha * dO_h v_h^T - delta_scaled[b,h,:,None]) as one 16-bit gradient plane (1,B,H,Tq,Tk): the dO V^T
    GEMM with the softmax backward fused into its epilogue (VilcoGemm.rowsub / emul)."""
    _, B, Tq, Cc = dO16.shape
    Tk = v.shape[2]
    d = Cc // H
    out = empty16(B, H, Tq, Tk, device=v.device, planes=1)
    L.gemm(dO16, v, out, M=Tq, N=Tk, K=d, a_rows=Tq, a_ld=Cc, a_s=(d, Tq * Cc), Z=(H, B), b_ld=Cc, b_s=(d, Tk * Cc), b_batched=True,
           d_ld=Tk, d_s=(Tq * Tk, H * Tq * Tk), alpha=alpha, a_lo=lo(dO16), b_lo=lo(v), rowsub=delta_scaled, rowsub_s=(Tq, H * Tq),
           emul=P16)
    return out


def xl_attention_ok(qw, T, C, H):
    """the fused XLNet relative-attention kernel takes single-plane operands, head dim 64, T % 128 == 0"""
    return FUSED_ATTN and qw.shape[0] == 1 and C // H == 64 and T % 128 == 0 and 128 <= T <= 2048


def xl_attention(qw, qr, k, v, kr, kmask, H, scale):
    """Fused XLNet relative attention: qw / qr / k / v (1,B,T,C), kr (1,2T,C), kmask (B,T) fp32 -> operand (1,B,T,C).
    Scores, the relative shift and the probabilities never leave the SM (csrc/xlattn.cu)."""
    _, B, T, Cc = qw.shape
    assert kr.shape[-2] == 2 * T and kr.shape[0] == 1 and k.shape[0] == 1 and v.shape[0] == 1 and qr.shape[0] == 1
    out = empty16(B, T, Cc, device=qw.device, planes=1)
    L.check(L.lib().vilco_xl_attention(_p(qw), _p(qr), _p(k), _p(v), _p(kr), _p(kmask), _p(out), B, H, T, Cc, C.c_float(scale),
                                       L.stream_ptr()), "vilco_xl_attention")
    return out


def layernorm(x, w, b, eps=1e-5, add=None, relu=False, pe=None, rowmul=None, zero_rows=None, out32=False, out16=True,
              y16=None, rows_per_batch=None, y_ld=None, y_bs=None, y16_lo=None, planes=None):
    """Channel LN over the last dim of token-major fp32 x.  Returns (y32 or None, y16 operand or None).
    `y16` may be a pre-allocated destination view (then pass y_ld / y_bs / y16_lo / rows_per_batch)."""
    assert x.dtype == f32 and x.is_contiguous()
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    y32 = torch.empty(x.shape, device=x.device, dtype=f32) if out32 else None
    if out16 and y16 is None:
        y16 = empty16(*x.shape, device=x.device, planes=planes)
        y16_lo = lo(y16)
    rpb = rows if rows_per_batch is None else rows_per_batch
    L.check(L.lib().vilco_layernorm(
        _p(x), L.F32, _p(add), _p(w), _p(b), C.c_float(eps), int(relu), _p(pe),
        0 if pe is None else pe.shape[0], _p(rowmul), _p(zero_rows), _p(y32), _p(y16), _i64(y16_lo or 0),
        _i64(Cc if y_ld is None else y_ld), _i64(rpb * Cc if y_bs is None else y_bs), rows, rpb, Cc,
        L.stream_ptr()), "vilco_layernorm")
    return y32, y16


def dwconv_ln(x, mask, wconvs, lnws, lnbs, stride, eps=1e-5, tlen=None):
    """x (B,T,C) fp32, mask (B,T) fp32 -> list of operands (NP,B,T/stride,C), one per (wconv, ln) set."""
    assert x.dtype == f32 and x.is_contiguous()
    B, T, Cc = x.shape
    n = len(wconvs)
    outs = [empty16(B, T // stride, Cc, device=x.device) for _ in range(n)]
    arr = C.c_void_p * n
    L.check(L.lib().vilco_dwconv_ln(
        _p(x), L.F32, _p(mask), _p(tlen), arr(*[w.data_ptr() for w in wconvs]),
        arr(*[w.data_ptr() for w in lnws]), arr(*[w.data_ptr() for w in lnbs]), arr(*[o.data_ptr() for o in outs]),
        _i64(lo(outs[0])), n, B, T, Cc, stride, C.c_float(eps), L.stream_ptr()), "vilco_dwconv_ln")
    return outs


def maxpool3s2(x):
    B, T, Cc = x.shape
    y = torch.empty(B, T // 2, Cc, device=x.device, dtype=f32)
    L.check(L.lib().vilco_maxpool3s2(_p(x), _p(y), B, T, Cc, L.stream_ptr()), "vilco_maxpool3s2")
    return y


def axpby(x, y=None, a=1.0, b=0.0, out32=True, out16=False, planes=None):
    """a*x + b*y on fp32 tensors -> (fp32 or None, operand or None)."""
    o32 = torch.empty_like(x) if out32 else None
    o16 = empty16(*x.shape, device=x.device, planes=planes) if out16 else None
    L.check(L.lib().vilco_axpby(_p(x), _p(y), C.c_float(a), C.c_float(b), _p(o32), _p(o16),
                                _i64(lo(o16) if out16 else 0), _i64(x.numel()), L.stream_ptr()), "vilco_axpby")
    return o32, o16


def scale_add(x, rowmul, y, scale):
    """x * rowmul[row] + scale[c] * y on fp32 (B,T,C) tensors."""
    out = torch.empty_like(x)
    Cc = x.shape[-1]
    L.check(L.lib().vilco_scale_add(_p(x), _p(rowmul), _p(y), _p(scale), _p(out), _i64(x.numel() // Cc), Cc,
                                    L.stream_ptr()), "vilco_scale_add")
    return out


def pack_feats(x, T_out=None, planes=None):
    """(B, C, T) fp32 (reference layout) -> operand (NP, B, T_out, C)."""
    assert x.dtype == f32 and x.is_contiguous()
    B, Cc, T = x.shape
    T_out = T if T_out is None else T_out
    y = empty16(B, T_out, Cc, device=x.device, planes=planes)
    L.check(L.lib().vilco_pack_feats(_p(x), _p(y), _i64(lo(y)), B, Cc, T, T_out, L.stream_ptr()), "vilco_pack_feats")
    return y


def groupnorm(x, w, b, groups, eps=1e-5, relu=False, out32=True, out16=False, planes=None):
    """nn.GroupNorm(groups, C) on token-major fp32 x (B, T, C) -> (y32 or None, y16 operand or None)"""
    assert x.dtype == f32 and x.is_contiguous() and x.dim() == 3
    B, T, Cc = x.shape
    y32 = torch.empty_like(x) if out32 else None
    y16 = empty16(B, T, Cc, device=x.device, planes=planes) if out16 else None
    L.check(L.lib().vilco_groupnorm(_p(x), _p(w), _p(b), _p(y32), _p(y16), _i64(lo(y16) if out16 else 0), B, T, Cc, groups,
                                    C.c_float(eps), int(relu), L.stream_ptr()), "vilco_groupnorm")
    return y32, y16


def upsample2_add(x, y):
    """y (B, 2T, C) += nearest-upsampled x (B, T, C), in place"""
    assert x.dtype == f32 and y.dtype == f32 and x.is_contiguous() and y.is_contiguous()
    B, T2, Cc = y.shape
    assert x.shape == (B, T2 // 2, Cc)
    L.check(L.lib().vilco_upsample2_add(_p(x), _p(y), B, T2, Cc, L.stream_ptr()), "vilco_upsample2_add")
    return y


def unpack(x, out=None):
    """(B, T, C) fp32 -> (B, C, T) fp32 (into `out` when given)."""
    B, T, Cc = x.shape
    y = torch.empty(B, Cc, T, device=x.device, dtype=f32) if out is None else out
    if out is not None:
        assert y.shape == (B, Cc, T) and y.dtype == f32 and y.is_contiguous() and x.is_contiguous()
    L.check(L.lib().vilco_unpack(_p(x), _p(y), B, T, Cc, L.stream_ptr()), "vilco_unpack")
    return y


def softmax_rows(S, kmask, mode=0, BD=None, scale=1.0, want32=False, planes=None):
    """S (B,H,Tq,Tk) fp32 -> P operand (NP,B,H,Tq,ldp) with ldp = roundup(Tk, 8) [, fp32 P (B,H,Tq,Tk) when want32]."""
    B, H, Tq, Tk = S.shape
    ldp = (Tk + 7) // 8 * 8
    P = empty16(B, H, Tq, ldp, device=S.device, planes=planes)
    P32 = torch.empty(B, H, Tq, Tk, device=S.device, dtype=f32) if want32 else None
    L.check(L.lib().vilco_softmax_rows(_p(S), _p(BD), _p(kmask), _p(P), _i64(lo(P)), _p(P32), B, H, Tq, Tk, _i64(ldp),
                                       C.c_float(scale), mode, L.stream_ptr()), "vilco_softmax_rows")
    return (P, P32) if want32 else P


def local_attention(q, k, v, mask, H, W, rel_pe=None):
    NP, B, T, Cc = q.shape
    assert k.shape == q.shape and v.shape == q.shape
    out = empty16(B, T, Cc, device=q.device, planes=NP)
    L.check(L.lib().vilco_local_attention(_p(q), _p(k), _p(v), _p(mask), _p(rel_pe), _p(out), _i64(lo(out)), B, T, Cc, H,
                                          W, L.stream_ptr()), "vilco_local_attention")
    return out


CHAN_TC = os.environ.get("VILCO_CHAN_TC", "1") == "1"   # 0: the SIMT channel-attention kernels


def channel_attention(qkv, H, tlen=None, return_A=False):
    """ChannelAttention core (blocks.py:423-436): per head A = softmax_rows((k / 8)^T v) over ALL T rows, y = (A q^T)^T.
    qkv operand (NP,B,T,3C) -> y operand (NP,B,T,C) [, A operand (NP,B,H,64,64)].
    Default: G = k^T v on the SIMT kernel (exact fp32 products), the row softmax, and y = q A^T on the tensor-core GEMM kernel.
    With per-sequence lengths (`tlen`, batched text evaluation) or VILCO_CHAN_TC=0 the SIMT apply kernel is used too."""
    _, B, T, C3 = qkv.shape
    Cc = C3 // 3
    if tlen is not None or not CHAN_TC or Cc // H != 64:
        G = torch.empty(B, H, 64, 64, device=qkv.device, dtype=f32)
        y = empty16(B, T, Cc, device=qkv.device)
        L.check(L.lib().vilco_channel_attention(_p(qkv), _i64(lo(qkv)), _p(G), _p(y), _i64(lo(y)), _p(tlen), B, T, Cc, H,
                                                L.stream_ptr()), "vilco_channel_attention")
        if not return_A:
            return y
        return y, softmax_rows(G.reshape(B, H, 64, 64), None, mode=0, planes=PLANES_HI)
    q = qkv[..., :Cc]
    G = torch.empty(B, H, 64, 64, device=qkv.device, dtype=f32)
    # G[i, j] = (1/8) sum_t k[t, i] v[t, j]: sums of T products feed a d x d softmax, which turns their ABSOLUTE error into a
    # relative error of A — the one place of the model that amplifies operand rounding.  So qkv arrives
    # in two planes and this stays on the SIMT kernel, which forms the full (hi + lo) x (hi + lo) products in fp32.
    L.check(L.lib().vilco_channel_attention(_p(qkv), _i64(lo(qkv)), _p(G), None, _i64(0), None, B, T, Cc, H, L.stream_ptr()),
            "vilco_channel_attention")
    A16 = softmax_rows(G, None, mode=0, planes=PLANES_HI)                 # (NP,B,H,64,64): the core stays on split operands
    y = empty16(B, T, Cc, device=qkv.device)
    # y[t, i] = sum_j q[t, j] A[i, j]
    L.gemm(q, A16, y, M=T, N=64, K=64, a_rows=T, a_ld=C3, a_s=(64, T * C3), Z=(H, B), b_ld=64, b_s=(64 * 64, H * 64 * 64),
           b_batched=True, d_ld=Cc, d_s=(64, T * Cc), a_lo=lo(qkv), b_lo=lo(A16), d_lo=lo(y))
    return (y, A16) if return_A else y


def channel_attention_bwd(dy, qkv, A16, H):
    """backward of channel_attention: dy (B,T,C) fp32, qkv operand (NP,B,T,3C), A operand (NP,B,H,64,64) -> dqkv (B,T,3C) fp32.
    Six launches of existing kernels: dy -> planes, dA = dy^T q, dq = dy A, softmax backward (dG as planes),
    dk = (1/8) v dG^T, dv = (1/8) k dG."""
    from . import backward as BW
    _, B, T, C3 = qkv.shape
    Cc = C3 // 3
    q, k, v = qkv[..., :Cc], qkv[..., Cc:2 * Cc], qkv[..., 2 * Cc:]
    dy16, _ = BW.to_planes(dy.reshape(-1, Cc))
    dy16 = dy16.reshape(dy16.shape[0], B, T, Cc)
    dqkv = torch.empty(B, T, C3, device=dy.device, dtype=f32)
    dA = torch.empty(B, H, 64, 64, device=dy.device, dtype=f32)
    gi = ginv()     # dy16 / dG16 are gradient planes (stored times GRAD_SCALE)
    # dA[i, j] = sum_t dy[t, i] q[t, j]
    L.gemm(dy16, q, dA, M=64, N=64, K=T, a_rows=64, a_ld=Cc, a_s=(64, T * Cc), a_major=1, Z=(H, B), b_ld=C3, b_s=(64, T * C3),
           b_batched=True, b_major=1, d_ld=64, d_s=(64 * 64, H * 64 * 64), a_lo=lo(dy16), b_lo=lo(qkv), alpha=gi)
    # dq[t, j] = sum_i dy[t, i] A[i, j]
    L.gemm(dy16, A16, dqkv, M=T, N=64, K=64, a_rows=T, a_ld=Cc, a_s=(64, T * Cc), Z=(H, B), b_ld=64, b_s=(64 * 64, H * 64 * 64),
           b_batched=True, b_major=1, d_ld=C3, d_s=(64, T * C3), a_lo=lo(dy16), b_lo=lo(A16), alpha=gi)
    _, dG16 = BW.softmax_bwd(dA, 1.0, P16=A16)                            # dG = A * (dA - rowsum(dA * A)) as planes
    # dk[t, i] = (1/8) sum_j v[t, j] dG[i, j]
    L.gemm(v, dG16, dqkv[..., Cc:2 * Cc], M=T, N=64, K=64, a_rows=T, a_ld=C3, a_s=(64, T * C3), Z=(H, B), b_ld=64,
           b_s=(64 * 64, H * 64 * 64), b_batched=True, d_ld=C3, d_s=(64, T * C3), alpha=0.125 * gi, a_lo=lo(qkv), b_lo=lo(dG16))
    # dv[t, j] = (1/8) sum_i k[t, i] dG[i, j]
    L.gemm(k, dG16, dqkv[..., 2 * Cc:], M=T, N=64, K=64, a_rows=T, a_ld=C3, a_s=(64, T * C3), Z=(H, B), b_ld=64,
           b_s=(64 * 64, H * 64 * 64), b_batched=True, b_major=1, d_ld=C3, d_s=(64, T * C3), alpha=0.125 * gi, a_lo=lo(qkv),
           b_lo=lo(dG16))
    return dqkv


def ew(op, x, y=None, rowmul=None, colmul=None, out32=True, out16=False):
    """fp32 elementwise helper (training path): op 0: x*rowmul*colmul, 1: gelu, 2: relu, 3: x*(y>0).
    Returns the fp32 result, or (fp32 or None, operand planes) when out16 is set."""
    out = torch.empty_like(x) if out32 else None
    o16 = empty16(*x.shape, device=x.device) if out16 else None
    Cc = x.shape[-1]
    L.check(L.lib().vilco_ew(op, _p(x), _p(y), _p(rowmul), _p(colmul), _p(out), _p(o16), _i64(lo(o16) if out16 else 0),
                             _i64(x.numel() // Cc), Cc, L.stream_ptr()), "vilco_ew")
    return (out, o16) if out16 else out


def dropout(x, p, seed, out32=True, out16=False, out=None):
    """inverted dropout with the counter-based mask of call `seed` (fp32 and / or operand-plane output)."""
    o32 = (torch.empty_like(x) if out is None else out) if out32 else None
    o16 = empty16(*x.shape, device=x.device) if out16 else None
    L.check(L.lib().vilco_dropout(_p(x), _p(o32), _p(o16), _i64(lo(o16) if out16 else 0), _i64(x.numel()), C.c_float(p),
                                  C.c_uint64(seed), L.stream_ptr()), "vilco_dropout")
    return (o32, o16) if out16 else o32
