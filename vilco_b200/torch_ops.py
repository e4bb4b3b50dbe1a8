"""The thin torch custom-op layer over the C ABI: `torch.ops.vilco.*` (torch.library), one op per kernel family of the
Moment-Query path.  Each op validates device / dtype, allocates its outputs with torch and makes exactly the C-ABI calls of
`vilco_b200.ops` (ctypes -> libvilco_b200.so); there is no CPU implementation — calling one with CPU tensors raises.  Fake
(meta) implementations give shapes / dtypes so the ops can be traced.  Importing this module registers the ops:

    import vilco_b200.torch_ops
    y = torch.ops.vilco.linear(x16, w16, bias, None, 0, True)

| op | C-ABI entry | reference function |
|---|---|---|
| vilco::linear, vilco::conv3 | vilco_gemm | MaskedConv1D / nn.Conv1d k=1, k=3 / nn.Linear (MQ/libs/modeling/blocks.py:106-130) |
| vilco::layernorm | vilco_layernorm | LayerNorm.forward (blocks.py:160-175) |
| vilco::attention | vilco_attention | MaskedMHCA / MaskedMHA core (blocks.py:228-269, 351-410) |
| vilco::self_attention | vilco_self_attention | MaskedMHCA core, single-pass kernel (blocks.py:351-410) |
| vilco::groupnorm | vilco_groupnorm | nn.GroupNorm of DenseAPP (MQ/libs/modeling/utils.py:671-729) |
| vilco::xl_attention | vilco_xl_attention | XLNetRelativeAttention.rel_attn_core (modeling_xlnet_x.py:256-320) |
| vilco::local_attention | vilco_local_attention | LocalMaskedMHCA core (blocks.py:1038-1207) |
| vilco::batched_nms | vilco_batched_nms (+ vilco_seg_voting) | libs.utils.nms.batched_nms (MQ/libs/utils/nms.py:103-190) |
"""
from typing import Optional, Tuple

import torch

from . import ops


def _planes_like(x16, *shape):
    return x16.new_empty((x16.shape[0],) + tuple(shape))


@torch.library.custom_op("vilco::linear", mutates_args=(), device_types="cuda")
def linear(x16: torch.Tensor, w16: torch.Tensor, bias: Optional[torch.Tensor], rowmul: Optional[torch.Tensor], act: int,
           out32: bool) -> torch.Tensor:
    """y = act((x w^T + bias) * rowmul[row]); x16 (P,...,K) / w16 (P,N,K) 16-bit operand planes -> fp32 (...,N) or planes"""
    return ops.linear(x16, w16, ops.f32 if out32 else ops.bf16, bias=bias, rowmul=rowmul, act=act)


@linear.register_fake
def _(x16, w16, bias, rowmul, act, out32):
    shp = tuple(x16.shape[1:-1]) + (w16.shape[1],)
    return x16.new_empty(shp, dtype=torch.float32) if out32 else _planes_like(x16, *shp)


@torch.library.custom_op("vilco::conv3", mutates_args=(), device_types="cuda")
def conv3(x16: torch.Tensor, w3: torch.Tensor, bias: Optional[torch.Tensor], rowmul: Optional[torch.Tensor], act: int) -> torch.Tensor:
    """k=3 stride-1 zero-padded conv over time: x16 (P,B,T,Cin), w3 (P,3,Cout,Cin) -> fp32 (B,T,Cout)"""
    return ops.conv3(x16, w3, ops.f32, bias=bias, rowmul=rowmul, act=act)


@conv3.register_fake
def _(x16, w3, bias, rowmul, act):
    return x16.new_empty((x16.shape[1], x16.shape[2], w3.shape[2]), dtype=torch.float32)


@torch.library.custom_op("vilco::layernorm", mutates_args=(), device_types="cuda")
def layernorm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float, relu: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """channel LayerNorm over the last dim of token-major fp32 x -> (fp32 y, operand planes of y)"""
    y32, y16 = ops.layernorm(x, weight, bias, eps, relu=relu, out32=True, out16=True)
    return y32, y16


@layernorm.register_fake
def _(x, weight, bias, eps, relu):
    return torch.empty_like(x), x.new_empty((ops.PLANES,) + tuple(x.shape), dtype=ops.ACT_DTYPE)


@torch.library.custom_op("vilco::attention", mutates_args=(), device_types="cuda")
def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, kmask: Optional[torch.Tensor], n_head: int, scale: float) -> torch.Tensor:
    """fused masked attention core (scores only in TMEM): q (P,B,Tq,C), k / v (P,B,Tk,C), kmask (B,Tk) -> (P,B,Tq,C)"""
    return ops.attention(q, k, v, kmask, n_head, scale)


@attention.register_fake
def _(q, k, v, kmask, n_head, scale):
    return torch.empty_like(q)


@torch.library.custom_op("vilco::xl_attention", mutates_args=(), device_types="cuda")
def xl_attention(qw: torch.Tensor, qr: torch.Tensor, k: torch.Tensor, v: torch.Tensor, kr: torch.Tensor, kmask: Optional[torch.Tensor],
                 n_head: int, scale: float) -> torch.Tensor:
    """fused XLNet relative attention (content + shifted position scores, softmax, P V in one kernel)"""
    return ops.xl_attention(qw, qr, k, v, kr, kmask, n_head, scale)


@xl_attention.register_fake
def _(qw, qr, k, v, kr, kmask, n_head, scale):
    return torch.empty_like(qw)


@torch.library.custom_op("vilco::local_attention", mutates_args=(), device_types="cuda")
def local_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, mask: torch.Tensor, n_head: int, window: int,
                    rel_pe: Optional[torch.Tensor]) -> torch.Tensor:
    """sliding-window attention core of LocalMaskedMHCA (window staged in shared memory, warp-shuffle softmax)"""
    return ops.local_attention(q, k, v, mask, n_head, window, rel_pe)


@local_attention.register_fake
def _(q, k, v, mask, n_head, window, rel_pe):
    return torch.empty_like(q)


@torch.library.custom_op("vilco::batched_nms", mutates_args=(), device_types="cuda")
def batched_nms(segs: torch.Tensor, scores: torch.Tensor, cls_idxs: torch.Tensor, iou_threshold: float, min_score: float,
                max_seg_num: int, use_soft_nms: bool, multiclass: bool, sigma: float, voting_thresh: float) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """libs.utils.batched_nms on device tensors -> (segs (n,2) f32, scores (n,) f32, labels (n,) i64), sorted by score"""
    from .utils.nms import batched_nms as _bn
    s, sc, lb = _bn(segs, scores, cls_idxs, iou_threshold, min_score, max_seg_num, use_soft_nms=use_soft_nms, multiclass=multiclass,
                    sigma=sigma, voting_thresh=voting_thresh)
    return s.to(segs.device), sc.to(segs.device), lb.to(segs.device)


@batched_nms.register_fake
def _(segs, scores, cls_idxs, iou_threshold, min_score, max_seg_num, use_soft_nms, multiclass, sigma, voting_thresh):
    n = min(int(max_seg_num), segs.shape[0])
    return segs.new_empty((n, 2)), scores.new_empty((n,)), cls_idxs.new_empty((n,), dtype=torch.int64)


@torch.library.custom_op("vilco::self_attention", mutates_args=(), device_types="cuda")
def self_attention(q16: torch.Tensor, k16: torch.Tensor, v16: torch.Tensor, kmask: Optional[torch.Tensor], n_head: int,
                   scale: float) -> torch.Tensor:
    """single-pass masked self-attention: q16 / k16 / v16 (1,B,T,C) single-plane operands, T % 128 == 0, head dim 64"""
    return ops.self_attention(q16, k16, v16, kmask, n_head, scale)


@self_attention.register_fake
def _(q16, k16, v16, kmask, n_head, scale):
    return torch.empty_like(q16)


@torch.library.custom_op("vilco::groupnorm", mutates_args=(), device_types="cuda")
def groupnorm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, groups: int, eps: float, relu: bool) -> torch.Tensor:
    """nn.GroupNorm(groups, C) (+ ReLU) of token-major fp32 x (B,T,C) -> fp32"""
    return ops.groupnorm(x, weight, bias, groups, eps, relu=relu, out32=True, out16=False)[0]


@groupnorm.register_fake
def _(x, weight, bias, groups, eps, relu):
    return torch.empty_like(x)


OPS = ("linear", "conv3", "layernorm", "groupnorm", "attention", "self_attention", "xl_attention", "local_attention", "batched_nms")
