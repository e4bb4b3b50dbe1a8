"""Backward primitives of the Moment-Query operators on the sm_100a kernels (training path, work in progress).

Conventions: gradients are fp32 token-major tensors; GEMM-shaped gradients reuse `vilco_gemm`
(dX = dZ W with W as MN-major B operand, dW = dZ^T X with X as MN-major B operand), everything else is a kernel of
csrc/bwd.cu.  Each function states the forward it differentiates.
"""
import ctypes as C

import torch

from . import lib as L
from . import ops
from .ops import _i64, _p, f32, lo


def to_planes(x, rowmul=None, colmul=None, want=True, want_t=False, batch_dims=0, grad=True, planes=None):
    """x (..., R, C) fp32 -> (y16 (NP,...,R,C) or None, yT16 (NP,...,C,ldT) or None), y = x * rowmul[:,None] * colmul[None,:].
    The leading `batch_dims` dims index independent matrices (transposed separately); otherwise x is flattened to 2-D.
    grad=True (the default here: this module handles gradients): the planes hold x * ops.GRAD_SCALE — every GEMM reading
    them multiplies its alpha by ops.ginv(); grad=False: an activation, stored as is."""
    assert x.dtype == f32 and x.is_contiguous()
    Cc = x.shape[-1]
    if batch_dims:
        lead = tuple(x.shape[:batch_dims])
        R = x.numel() // Cc
        for n in lead:
            R //= n
        Z = x.numel() // (R * Cc)
    else:
        lead, R, Z = (), x.numel() // Cc, 1
    y = ops.empty16(*lead, R, Cc, device=x.device, grad=grad, planes=planes) if want else None
    ldT = (R + 7) // 8 * 8
    # padding columns (ldT > R) must be zero: they are read as K entries of the weight-gradient GEMMs
    yT = (ops.zeros16 if ldT != R else ops.empty16)(*lead, Cc, ldT, device=x.device, grad=grad, planes=planes) if want_t else None
    L.check(L.lib().vilco_to_planes(_p(x), _p(rowmul), _p(colmul), _p(y), _i64(lo(y) if want else 0), _p(yT),
                                    _i64(lo(yT) if want_t else 0), R, Cc, ldT, Z, int(bool(grad)), L.stream_ptr()),
            "vilco_to_planes")
    return y, yT


def colsum(x, y=None, rowmul=None, out=None):
    Cc = x.shape[-1]
    R = x.numel() // Cc
    if out is None:
        out = torch.zeros(Cc, device=x.device, dtype=f32)
    L.check(L.lib().vilco_colsum(_p(x), _p(y), _p(rowmul), _p(out), R, Cc, L.stream_ptr()), "vilco_colsum")
    return out


def wgrad(dz, x2, Mo, No, R, alpha=1.0, out=None):
    """dW (Mo, No) = dz^T @ x with dz (NP, R, Mo) and x2 (NP, R, No) both in their natural token-major layout: dz is read as
    an MN-major A operand and x as an MN-major B operand, so no transposed copy is made.  The reduction runs over all B*T
    rows while the output has few tiles, so K is split over batches (split-K) whenever the plain launch would leave most
    of the 148 SMs idle; partials are summed by vilco_colsum.  out: fp32 (Mo, No) buffer to ACCUMULATE into (the
    parameter's .grad view)."""
    alpha = alpha * ops.ginv()        # dz is a gradient-plane operand
    # split-K so that the 256 x 256 pair tiles (vilco_gemm picks them from 37 tiles on) fill the 74 CTA pairs once: measured
    # (tools/gemm2_probe.py bwd) C x C at 32768 rows 141 -> 81 us with S = 4 pair tiles instead of S = 2 single-CTA tiles
    if Mo > 128 and No >= 256:
        tiles = ((Mo + 255) // 256) * ((No + 255) // 256)
        S = min(16, 74 // tiles) if tiles < 37 else 1
    else:
        tiles = ((Mo + 127) // 128) * ((No + 127) // 128)
        S = min(16, 148 // tiles) if tiles < 100 else 1
    while S > 1 and (R % (8 * S) != 0 or R // S < 512):
        S -= 1
    if S <= 1:
        dw = torch.empty(Mo, No, device=x2.device, dtype=f32) if out is None else out
        L.gemm(dz, x2, dw, M=Mo, N=No, K=R, a_rows=Mo, a_ld=Mo, a_major=1, b_ld=No, d_ld=No, b_major=1, alpha=alpha,
               a_lo=lo(dz), b_lo=lo(x2), resid=out)
        return dw
    chunk = R // S
    part = torch.empty(S, Mo, No, device=x2.device, dtype=f32)
    L.gemm(dz, x2, part, M=Mo, N=No, K=chunk, a_rows=Mo, a_ld=Mo, a_major=1, a_s=(chunk * Mo, 0), Z=(S, 1), b_ld=No,
           b_s=(chunk * No, 0), b_batched=True, b_major=1, d_ld=No, d_s=(Mo * No, 0), alpha=alpha, a_lo=lo(dz), b_lo=lo(x2))
    return colsum(part.reshape(S, Mo * No), out=None if out is None else out.reshape(-1)).reshape(Mo, No)


def linear_bwd(dy, x16, w16, rowmul=None, alpha=1.0, need_dx=True, need_dw=True, need_db=True, dw_out=None, db_out=None):
    """forward: y = (alpha * x w^T + bias) * rowmul[:, None]   (ops.linear without act / colscale / resid).
    dy (..., N) fp32, x16 (NP, ..., K), w16 (NP, N, K) -> (dx (..., K) fp32, dw (N, K) fp32, db (N,) fp32).
    dw_out / db_out: fp32 buffers the parameter gradients are accumulated into instead of fresh tensors."""
    N, K = w16.shape[1], w16.shape[2]
    dy2 = dy.reshape(-1, N)
    R = dy2.shape[0]
    dz, _ = to_planes(dy2, rowmul, None)
    dx = dw = db = None
    if need_dx:
        dx = torch.empty(*dy.shape[:-1], K, device=dy.device, dtype=f32)
        L.gemm(dz, w16, dx, M=R, N=K, K=N, a_rows=R, a_ld=N, b_ld=K, d_ld=K, b_major=1, alpha=alpha * ops.ginv(), a_lo=lo(dz),
               b_lo=lo(w16))
    if need_dw:
        dw = wgrad(dz, x16.reshape(x16.shape[0], -1, K), N, K, R, alpha, out=dw_out)
    if need_db:
        db = colsum(dy2, rowmul=rowmul, out=db_out)
    return dx, dw, db


def linear_bwd16(dz, x16, w16, need_dx=True, dw_out=None):
    """linear_bwd for a gradient that already exists as operand planes dz (NP, R, N) (bias gradient done by the producer)."""
    N, K = w16.shape[1], w16.shape[2]
    R = dz.shape[1]
    dx = None
    if need_dx:
        dx = torch.empty(R, K, device=dz.device, dtype=f32)
        L.gemm(dz, w16, dx, M=R, N=K, K=N, a_rows=R, a_ld=N, b_ld=K, d_ld=K, b_major=1, alpha=ops.ginv(), a_lo=lo(dz), b_lo=lo(w16))
    dw = wgrad(dz, x16.reshape(x16.shape[0], -1, K), N, K, R, out=dw_out)
    return dx, dw


def shift_planes(x16, shift):
    """x16 (NP,B,T,C) -> same shape, rows shifted inside every clip: y[b,t] = x[b,t+shift] (zero outside)."""
    NP, B, T, Cc = x16.shape
    y = torch.empty_like(x16)
    L.check(L.lib().vilco_shift_planes(_p(x16), _p(y), _i64(lo(x16)), B, T, Cc, shift, L.stream_ptr()), "vilco_shift_planes")
    return y


def conv3_bwd(dy, x16, w3, w3_flip, rowmul=None, need_dx=True):
    """forward: y[b,t] = (sum_tap x[b,t+tap-1] w3[tap]^T + bias) * rowmul[b,t]   (ops.conv3).
    dy (B,T,N) fp32, x16 (NP,B,T,K), w3 / w3_flip (NP,3,N,K) (flip = taps reversed) -> (dx (B,T,K), dw3 (3,N,K), db (N,))."""
    B, T, N = dy.shape
    K = w3_flip.shape[3]
    dz, _ = to_planes(dy.reshape(-1, N), rowmul.reshape(-1) if rowmul is not None else None, None)
    dx = None
    if need_dx:
        # dx[b,t] = sum_tap dz[b, t - tap + 1] w3[tap] = sum_tap' dz[b, t + tap' - 1] w3[2 - tap']: the forward conv kernel
        # on dz with the tap-reversed weights as MN-major B operand
        dx = torch.empty(B, T, K, device=dy.device, dtype=f32)
        dz4 = dz.reshape(dz.shape[0], B, T, N)
        L.gemm(dz4, w3_flip, dx, M=T, N=K, K=N, a_rows=T, a_ld=N, a_s=(0, T * N), Z=(1, B), taps=3, b_ld=K, b_s=(N * K, 0),
               b_major=1, d_ld=K, d_s=(0, T * K), alpha=ops.ginv(), a_lo=lo(dz4), b_lo=lo(w3_flip))
    R = B * T
    taps = []
    for tap in range(3):
        xs = x16 if tap == 1 else shift_planes(x16, tap - 1)
        taps.append(wgrad(dz, xs.reshape(xs.shape[0], R, K), N, K, R))
    dw = torch.stack(taps)
    db = colsum(dy.reshape(-1, N), rowmul=rowmul.reshape(-1) if rowmul is not None else None)
    return dx, dw, db


def layernorm_bwd(dy, x, w, eps=1e-5, add=None, y_relu=None, need_dw=True, dw_out=None, db_out=None):
    """forward: y = act(LN(x [+ add]) * w + b)  (ops.layernorm; y_relu = fp32 forward output when relu=True).
    Returns (dx (also the gradient of `add`), dw, db)."""
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    dx = torch.empty_like(x)
    dw = (torch.zeros(Cc, device=x.device, dtype=f32) if dw_out is None else dw_out) if need_dw else None
    db = (torch.zeros(Cc, device=x.device, dtype=f32) if db_out is None else db_out) if need_dw else None
    L.check(L.lib().vilco_layernorm_bwd(_p(x), _p(add), _p(w), _p(dy.contiguous()), _p(y_relu), C.c_float(eps), _p(dx), _p(dw),
                                        _p(db), rows, Cc, L.stream_ptr()), "vilco_layernorm_bwd")
    return dx, dw, db


def gelu_bwd(dy, x):
    dx = torch.empty_like(x)
    L.check(L.lib().vilco_gelu_bwd(_p(x), _p(dy.contiguous()), _p(dx), _i64(x.numel()), L.stream_ptr()), "vilco_gelu_bwd")
    return dx


def maxpool3s2_bwd(dy, x):
    B, T, Cc = x.shape
    dx = torch.zeros_like(x)
    L.check(L.lib().vilco_maxpool3s2_bwd(_p(x), _p(dy.contiguous()), _p(dx), B, T, Cc, L.stream_ptr()), "vilco_maxpool3s2_bwd")
    return dx


def dwconv_ln_bwd(dys, x, mask, wconvs, lnws, stride, eps=1e-5):
    """forward: out_i = LN_i(dwconv_i(x) * mask_out)  (ops.dwconv_ln), i over (q, k, v).
    dys: list of (B,T/stride,C) fp32.  Returns (dx (B,T,C), [dwconv_i (3,C)], [dlnw_i], [dlnb_i]).
    The conv outputs are recomputed from x (cheap, depthwise) instead of being stored."""
    B, T, Cc = x.shape
    To = T // stride
    dx = torch.zeros_like(x)
    dwc, dlw, dlb = [], [], []
    om = mask[:, ::stride].contiguous() if stride > 1 else mask
    for i, dy in enumerate(dys):
        # recompute conv_i(x) * mask (fp32) with torch-free kernels: depthwise conv = 3 shifted axpby ... kept simple:
        conv = _dwconv_fwd32(x, mask, wconvs[i], stride)
        dconv, dw_ln, db_ln = layernorm_bwd(dy, conv, lnws[i], eps)
        dw = torch.zeros(3, Cc, device=x.device, dtype=f32)
        L.check(L.lib().vilco_dwconv_bwd(_p(x), _p(mask), _p(wconvs[i]), _p(dconv), _p(dx), _p(dw), B, T, Cc, stride, 1,
                                         L.stream_ptr()), "vilco_dwconv_bwd")
        dwc.append(dw); dlw.append(dw_ln); dlb.append(db_ln)
    return dx, dwc, dlw, dlb


def _dwconv_fwd32(x, mask, w3c, stride):
    """fp32 depthwise conv * out-mask via the existing forward kernel with an identity LayerNorm is not available, so the
    recomputation uses the dwconv_ln kernel's math through vilco_dwconv_fwd32 (plain fp32 output)."""
    B, T, Cc = x.shape
    out = torch.empty(B, T // stride, Cc, device=x.device, dtype=f32)
    L.check(L.lib().vilco_dwconv_fwd32(_p(x), _p(mask), _p(w3c), _p(out), B, T, Cc, stride, L.stream_ptr()), "vilco_dwconv_fwd32")
    return out


def softmax_bwd(dP, scale, P32=None, P16=None, want32=False, want16=True):
    """dS = scale * P * (dP - rowsum(dP * P)) over the last dim of (B,H,Tq,Tk).  P from fp32 rows or from the operand planes
    of the forward softmax; returns (dS fp32 or None, dS operand planes (NP,B,H,Tq,ldp) or None)."""
    B, H, Tq, Tk = dP.shape
    ldp = (Tk + 7) // 8 * 8
    dS = torch.empty_like(dP) if want32 else None
    dS16 = ops.empty16(B, H, Tq, ldp, device=dP.device, grad=True) if want16 else None
    L.check(L.lib().vilco_softmax_bwd(_p(P32), _p(P16), _i64(lo(P16) if P16 is not None else 0),
                                      _i64(P16.shape[-1] if P16 is not None else 0), _p(dP), _p(dS), _p(dS16),
                                      _i64(lo(dS16) if want16 else 0), _i64(ldp), _i64(B * H * Tq), Tk, C.c_float(scale),
                                      L.stream_ptr()), "vilco_softmax_bwd")
    return dS, dS16


def attention_bwd(dO, q16, k16, v16, kmask, H, scale):
    """forward: O = softmax_j(scale * q k^T | kmask) v per head (ops.attention / the materialised chain).
    dO (B,Tq,C) fp32, q16 (NP,B,Tq,C), k16 / v16 (NP,B,Tk,C) -> (dq, dk, dv) fp32 token-major.
    The probabilities are recomputed (QK^T GEMM + softmax) instead of being stored by the fused forward kernel; P and dS
    only ever exist as operand planes, and their transposes are never formed (MN-major A operands)."""
    _, B, Tq, Cc = q16.shape
    Tk = k16.shape[2]
    S = ops.attn_scores(q16, k16, H, scale)
    P16 = ops.softmax_rows(S, kmask, mode=0)                        # (NP,B,H,Tq,ldp)
    del S
    dO16, _ = to_planes(dO.reshape(-1, Cc))
    dO16 = dO16.reshape(dO16.shape[0], B, Tq, Cc)
    gi = ops.ginv()                                                  # dO16 / dS16 are gradient planes (times GRAD_SCALE)
    dP = ops.attn_scores(dO16, v16, H, gi)                           # dP[b,h] = dO_h v_h^T
    _, dS16 = softmax_bwd(dP, scale, P16=P16)
    del dP
    dq = ops.attn_pv(dS16, k16, H, Tk, out32=True, alpha=gi)                   # dQ_h = dS K_h
    dk = ops.attn_pv(dS16, q16, H, Tq, out32=True, a_trans=True, M=Tk, alpha=gi)   # dK_h = dS^T Q_h
    dv = ops.attn_pv(P16, dO16, H, Tq, out32=True, a_trans=True, M=Tk, alpha=gi)   # dV_h = P^T dO_h
    return dq, dk, dv


def attention_bwd_lse(dO, O16, lse2, q16, k16, v16, kmask, H, scale):
    """attention_bwd for the single-pass forward kernel, which saved the row log-sum-exp: no (T x T) tensor is ever written in
    fp32 and no softmax kernel runs —
        P  = exp2(scale log2e q k^T - lse) * kmask          in the epilogue of the QK^T GEMM      (one fp16 plane)
        dS = P * (dO v^T - delta) * scale,  delta = rowsum(dO * O) per head   in the epilogue of the dO V^T GEMM
    then dQ = dS K, dK = dS^T Q, dV = P^T dO as before.  dO (B,T,C) fp32, O16 / q16 / k16 / v16 single-plane operands."""
    _, B, T, Cc = q16.shape
    d = Cc // H
    P16 = ops.attn_probs_from_lse(q16, k16, H, scale, lse2, kmask)
    dO16, _ = to_planes(dO.reshape(-1, Cc), planes=1)
    dO16 = dO16.reshape(1, B, T, Cc)
    # delta[b,h,i] = sum_c dO[b,i,h,c] O[b,i,h,c]   (tiny glue reduction over the head dim)
    delta = (dO.reshape(B, T, H, d) * O16[0].reshape(B, T, H, d).float()).sum(-1).permute(0, 2, 1).contiguous()
    # dO16 carries GRAD_SCALE: (acc * scale - delta * scale * GS) = GS * scale * (dP - delta) -> dS as a gradient plane
    dS16 = ops.attn_dscores_fused(dO16, v16, H, P16, delta * (scale * ops.GRAD_SCALE), scale)
    gi = ops.ginv()
    dq = ops.attn_pv(dS16, k16, H, T, out32=True, alpha=gi)
    dk = ops.attn_pv(dS16, q16, H, T, out32=True, a_trans=True, M=T, alpha=gi)
    dv = ops.attn_pv(P16, dO16, H, T, out32=True, a_trans=True, M=T, alpha=gi)
    return dq, dk, dv


def local_attention_bwd(dO, q16, k16, v16, mask, H, W, rel_pe=None):
    """forward: ops.local_attention (LocalMaskedMHCA core).  dO (B,T,C) fp32, q16 / k16 / v16 (NP,B,T,C) -> (dq, dk, dv) fp32."""
    _, B, T, Cc = q16.shape
    dq, dk, dv = (torch.empty(B, T, Cc, device=dO.device, dtype=f32) for _ in range(3))
    sp = torch.empty(B, H, T, W, device=dO.device, dtype=f32)
    sds = torch.empty(B, H, T, W, device=dO.device, dtype=f32)
    L.check(L.lib().vilco_local_attention_bwd(_p(dO.contiguous()), _p(q16), _p(k16), _p(v16), _i64(lo(q16)), _p(mask), _p(rel_pe),
                                              _p(sp), _p(sds), _p(dq), _p(dk), _p(dv), B, T, Cc, H, W, L.stream_ptr()),
            "vilco_local_attention_bwd")
    return dq, dk, dv
