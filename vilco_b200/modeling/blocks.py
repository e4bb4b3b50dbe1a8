"""Host-side mirror of ViLCo's MQ/libs/modeling/blocks.py operator classes.

The classes keep the reference's constructor signatures, parameter names and shapes (so reference checkpoints load
and `make_optimizer`'s isinstance / name rules keep working, MQ/libs/utils/train_utils.py:68-145) but their
`forward` runs the hand-written sm_100a kernels through the C ABI — there is no PyTorch compute fallback.
Standalone module forwards take / return the reference layout (B, C, T) + bool mask (B, 1, T); inside the full
model the engine (vilco_b200/engine.py) drives the same kernels token-major without the layout round trips.
"""
import math

import torch
from torch import nn

from .. import engine as E
from .. import ops


def _pack_module(mod, prefix=""):
    sd = {prefix + k: v for k, v in mod.state_dict().items()}
    return E.pack_weights(sd, next(mod.parameters()).device)


def _to_tokens(x):
    """(B, C, T) fp32 -> (B, T, C) fp32 contiguous (layout plumbing only)."""
    return x.detach().float().transpose(1, 2).contiguous()


def _mask_f(mask):
    return mask.detach().reshape(mask.shape[0], mask.shape[-1]).float().contiguous()


class MaskedConv1D(nn.Module):
    """Masked 1D convolution (reference: blocks.py:57-130).  Dense k in {1,3} stride 1 runs on the tcgen05 GEMM;
    depthwise k=3 (stride 1|2) is only used fused with its LayerNorm inside MaskedMHCA."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 padding_mode="zeros"):
        super().__init__()
        assert (kernel_size % 2 == 1) and (kernel_size // 2 == padding)
        self.stride = stride
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias,
                              padding_mode)
        if bias:
            torch.nn.init.constant_(self.conv.bias, 0.0)

    def augment_classification(self, num_new_classes, device):
        """Widen the classifier by `num_new_classes` outputs (reference: blocks.py:85-104)."""
        old = self.conv
        out_class = old.out_channels
        new = nn.Conv1d(old.in_channels, out_class + num_new_classes, 3, stride=1, padding=1, bias=True).to(device)
        torch.nn.init.constant_(new.bias, -(math.log((1 - 0.01) / 0.01)))
        new.weight.data[:out_class] = old.weight.data[:out_class]
        new.bias.data[:out_class] = old.bias.data[:out_class]
        self.conv = new

    @torch.no_grad()
    def forward(self, x, mask):
        B, C, T = x.size()
        assert T % self.stride == 0
        k = self.conv.kernel_size[0]
        if self.conv.groups != 1 or self.stride != 1 or k not in (1, 3):
            raise NotImplementedError("standalone MaskedConv1D.forward supports dense k in {1,3}, stride 1 "
                                      "(depthwise convs run fused inside MaskedMHCA)")
        W = _pack_module(self)
        x16 = ops.pack_feats(x.detach().float().contiguous())
        m = _mask_f(mask)
        bias = W.get("conv.bias")
        if k == 1:
            y = ops.linear(x16, W["conv.weight"], ops.f32, bias=bias, rowmul=m.reshape(-1))
        else:
            y = ops.conv3(x16, W["conv.weight"], ops.f32, bias=bias, rowmul=m)
        return ops.unpack(y), mask.bool()


class LayerNorm(nn.Module):
    """LayerNorm over the channel dim of (B, C, T) (reference: blocks.py:133-175)."""

    def __init__(self, num_channels, eps=1e-5, affine=True, device=None, dtype=None):
        super().__init__()
        self.num_channels, self.eps, self.affine = num_channels, eps, affine
        assert affine, "only the affine variant is used by the MQ model"
        kw = {"device": device, "dtype": dtype}
        self.weight = nn.Parameter(torch.ones([1, num_channels, 1], **kw))
        self.bias = nn.Parameter(torch.zeros([1, num_channels, 1], **kw))

    @torch.no_grad()
    def forward(self, x):
        assert x.dim() == 3 and x.shape[1] == self.num_channels
        y32, _ = ops.layernorm(_to_tokens(x), self.weight.detach().float().reshape(-1).contiguous(),
                               self.bias.detach().float().reshape(-1).contiguous(), self.eps, out32=True, out16=False)
        return ops.unpack(y32)


class MaskedMHA(nn.Module):
    """Multi-head (cross) attention (reference: blocks.py:194-269)."""

    def __init__(self, n_embd, n_head, attn_pdrop=0.0, proj_pdrop=0.0):
        super().__init__()
        assert n_embd % n_head == 0
        self.n_embd, self.n_head = n_embd, n_head
        self.key = nn.Conv1d(n_embd, n_embd, 1)
        self.query = nn.Conv1d(n_embd, n_embd, 1)
        self.value = nn.Conv1d(n_embd, n_embd, 1)
        self.attn_drop = nn.Dropout(attn_pdrop)
        self.proj_drop = nn.Dropout(proj_pdrop)
        self.proj = nn.Conv1d(n_embd, n_embd, 1)

    @torch.no_grad()
    def forward(self, x, mask, encoder_hidden_states=None, encoder_attention_mask=None):
        assert encoder_hidden_states is not None, "MQ only uses MaskedMHA as cross attention"
        W = _pack_module(self)
        _, x16 = ops.axpby(_to_tokens(x), None, 1.0, 0.0, out32=False, out16=True)
        _, y16 = ops.axpby(_to_tokens(encoder_hidden_states), None, 1.0, 0.0, out32=False, out16=True)
        ym = encoder_attention_mask.detach().float().contiguous()
        o = E.cross_attn_fwd(W, "", x16, y16, ym, self.n_head)
        m = mask.detach().reshape(x.shape[0], -1).float().contiguous()
        out = ops.linear(o, W["proj.weight"], ops.f32, bias=W["proj.bias"], rowmul=m.reshape(-1))
        return ops.unpack(out), mask


class MaskedMHCA(nn.Module):
    """Multi-head conv attention, global (reference: blocks.py:272-410)."""

    window_size = -1

    def __init__(self, n_embd, n_head, n_qx_stride=1, n_kv_stride=1, attn_pdrop=0.0, proj_pdrop=0.0):
        super().__init__()
        assert n_embd % n_head == 0
        self.n_embd, self.n_head = n_embd, n_head
        assert (n_qx_stride == 1) or (n_qx_stride % 2 == 0)
        assert (n_kv_stride == 1) or (n_kv_stride % 2 == 0)
        self.n_qx_stride, self.n_kv_stride = n_qx_stride, n_kv_stride
        ks = n_qx_stride + 1 if n_qx_stride > 1 else 3
        self.query_conv = MaskedConv1D(n_embd, n_embd, ks, stride=n_kv_stride, padding=ks // 2, groups=n_embd, bias=False)
        self.query_norm = LayerNorm(n_embd)
        ks = n_kv_stride + 1 if n_kv_stride > 1 else 3
        self.key_conv = MaskedConv1D(n_embd, n_embd, ks, stride=n_kv_stride, padding=ks // 2, groups=n_embd, bias=False)
        self.key_norm = LayerNorm(n_embd)
        self.value_conv = MaskedConv1D(n_embd, n_embd, ks, stride=n_kv_stride, padding=ks // 2, groups=n_embd, bias=False)
        self.value_norm = LayerNorm(n_embd)
        self.key = nn.Conv1d(n_embd, n_embd, 1)
        self.query = nn.Conv1d(n_embd, n_embd, 1)
        self.value = nn.Conv1d(n_embd, n_embd, 1)
        self.attn_drop = nn.Dropout(attn_pdrop)
        self.proj_drop = nn.Dropout(proj_pdrop)
        self.proj = nn.Conv1d(n_embd, n_embd, 1)

    @torch.no_grad()
    def forward(self, x, mask):
        W = _pack_module(self)
        o, om = E.mhca_fwd(W, "", _to_tokens(x), _mask_f(mask), self.n_head, self.n_kv_stride, self.window_size)
        out = ops.linear(o, W["proj.weight"], ops.f32, bias=W["proj.bias"], rowmul=om.reshape(-1))
        return ops.unpack(out), om.bool().unsqueeze(1)


class LocalMaskedMHCA(MaskedMHCA):
    """Local (windowed) multi-head conv attention (reference: blocks.py:871-1207)."""

    def __init__(self, n_embd, n_head, window_size, n_qx_stride=1, n_kv_stride=1, attn_pdrop=0.0, proj_pdrop=0.0,
                 use_rel_pe=False):
        super().__init__(n_embd, n_head, n_qx_stride, n_kv_stride, attn_pdrop, proj_pdrop)
        assert window_size > 1 and window_size % 2 == 1
        self.window_size = window_size
        self.window_overlap = window_size // 2
        self.use_rel_pe = use_rel_pe
        if use_rel_pe:
            self.rel_pe = nn.Parameter(torch.zeros(1, 1, n_head, window_size))
            nn.init.trunc_normal_(self.rel_pe, std=(2.0 / n_embd) ** 0.5)


class ChannelAttention(nn.Module):
    """reference: blocks.py:412-436"""

    def __init__(self, dim, num_heads=8, qkv_bias=False):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)


class ChannelBlock(nn.Module):
    """reference: blocks.py:438-466 (norm1 is registered but never applied, as in the reference)"""

    def __init__(self, n_embd, num_heads, mlp_ratio=4.0, qkv_bias=False, drop_path=0.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(n_embd)
        self.attn = ChannelAttention(n_embd, num_heads=num_heads, qkv_bias=qkv_bias)
        self.norm2 = nn.LayerNorm(n_embd)
        n_hidden = int(n_embd * mlp_ratio)
        self.mlp = nn.Sequential(nn.Linear(n_embd, n_hidden), nn.GELU(), nn.Linear(n_hidden, n_embd))


class Scale(nn.Module):
    """reference: blocks.py:605-623"""

    def __init__(self, init_value=1.0):
        super().__init__()
        self.scale = nn.Parameter(torch.tensor(init_value, dtype=torch.float32), requires_grad=True)

    def forward(self, x):
        return x * self.scale


class AffineDropPath(nn.Module):
    """Per-channel scale (+ stochastic depth in training) — reference: blocks.py:655-670."""

    def __init__(self, num_dim, drop_prob=0.0, init_scale_value=1e-4):
        super().__init__()
        self.scale = nn.Parameter(init_scale_value * torch.ones((1, num_dim, 1)), requires_grad=True)
        self.drop_prob = drop_prob


class TransformerBlock(nn.Module):
    """reference: blocks.py:468-593"""

    def __init__(self, n_embd, n_head, n_ds_strides=(1, 1), n_out=None, n_hidden=None, act_layer=nn.GELU,
                 attn_pdrop=0.0, proj_pdrop=0.0, path_pdrop=0.0, t_c_alpha=0.8, use_rel_pe=False, use_cross_modal=False,
                 use_adaper=True, mha_win_size=-1):
        super().__init__()
        assert len(n_ds_strides) == 2
        self.t_c_alpha = t_c_alpha
        self.n_head = n_head
        self.ln1 = LayerNorm(n_embd)
        self.ln2 = LayerNorm(n_embd)
        if mha_win_size > 1:
            self.attn = LocalMaskedMHCA(n_embd, n_head, mha_win_size, n_ds_strides[0], n_ds_strides[1], attn_pdrop,
                                        proj_pdrop, use_rel_pe)
        else:
            self.attn = MaskedMHCA(n_embd, n_head, n_ds_strides[0], n_ds_strides[1], attn_pdrop, proj_pdrop)
        self.use_cross_modal = use_cross_modal
        if use_cross_modal:
            self.cross_attn = MaskedMHA(n_embd, n_head, attn_pdrop=attn_pdrop, proj_pdrop=proj_pdrop)
            self.ln3 = LayerNorm(n_embd)
        self.n_ds_strides = n_ds_strides
        n_hidden = 4 * n_embd if n_hidden is None else n_hidden
        n_out = n_embd if n_out is None else n_out
        self.mlp = nn.Sequential(nn.Conv1d(n_embd, n_hidden, 1), act_layer(), nn.Dropout(proj_pdrop, inplace=True),
                                 nn.Conv1d(n_hidden, n_out, 1), nn.Dropout(proj_pdrop, inplace=True))
        self.channel_attn = ChannelBlock(n_embd, n_head, drop_path=path_pdrop)
        if path_pdrop > 0.0:
            self.drop_path_attn = AffineDropPath(n_embd, drop_prob=path_pdrop)
            self.drop_path_mlp = AffineDropPath(n_out, drop_prob=path_pdrop)
        else:
            self.drop_path_attn = nn.Identity()
            self.drop_path_mlp = nn.Identity()

    @torch.no_grad()
    def forward(self, x, mask, cross_y=None, cross_y_mask=None, pos_embd=None):
        assert pos_embd is None
        W = _pack_module(self)
        cross = None
        if self.use_cross_modal and cross_y is not None:
            cross = (_to_tokens(cross_y), cross_y_mask.detach().float().contiguous())
        out, om = E.transformer_block_fwd(W, "", _to_tokens(x), _mask_f(mask), self.n_head, self.n_ds_strides[0], cross,
                                          self.t_c_alpha, self.attn.window_size)
        return ops.unpack(out), om.bool().unsqueeze(1)
