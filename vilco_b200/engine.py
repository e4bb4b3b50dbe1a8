"""Forward engine of the Moment-Query model on hand-written sm_100a kernels (token-major, 16-bit operand planes,
fp32 accumulation / residual stream / statistics).

Operand-format policy (ops.set_precision): in the default "mixed" mode every GEMM reads single fp16
planes except the contractions upstream of the channel-attention softmax that the parity bar is sensitive to — the input
projection, the embedding convolutions, the channel-attention qkv projection and core — whose operands carry two planes
(`ops.PLANES_HI`; `sensitive_key` for the weights).

`pack_weights` turns a reference-named state_dict into the operand formats the kernels want;
the `*_fwd` functions are the B200 counterparts of the reference modules (file:line cited per function).  All
tensors here are token-major: activations (B, T, C), masks (B, T) fp32 1/0.
"""
import math
import os

import numpy as np
import torch

from . import ops
from .ops import ACT_GELU, ACT_NONE, ACT_RELU, bf16, f32


# ----------------------------------------------------------------------------------------------------
# weight packing
# ----------------------------------------------------------------------------------------------------
def _pack_kind(k, v):
    last = k.rsplit(".", 1)[-1]
    if ".rel_attn." in k and last in ("q", "k", "v", "r"):
        return "xl_t"
    if ".rel_attn." in k and last == "o":
        return "gemm"
    if k.endswith("_conv.conv.weight") or (last == "weight" and v.dim() == 3 and v.shape[1] == 1 and v.shape[2] == 3 and v.shape[0] > 1):
        return "dw"        # depthwise k=3 (attention q/k/v convs, FPN1D fpn_convs)
    if last == "weight" and v.dim() == 3 and v.shape[0] == 1 and v.shape[2] == 1:
        return "vec"
    if last == "weight" and v.dim() == 3 and v.shape[2] == 3:
        return "conv3"
    if last == "weight" and (v.dim() == 2 or (v.dim() == 3 and v.shape[2] == 1)):
        return "gemm"
    return "vec"


def sensitive_key(k):
    """weights of the contractions that keep split (two-plane) operands in the mixed mode"""
    return k.startswith(("backbone.proj.", "backbone.embd.", "backbone.vid_embd.", "backbone.txt_embd.")) or \
        k.endswith(".channel_attn.attn.qkv.weight") or ".ac_conv." in k      # FPN1D's DenseAPP stack: five GroupNorm blocks deep


def _planes_for(k):
    return ops.PLANES_HI if sensitive_key(k) else ops.PLANES


def _pack_one(kind, v, planes=None):
    if kind == "xl_t":
        return ops.split16(v.reshape(v.shape[0], -1).t().contiguous(), planes)
    if kind == "gemm":
        return ops.split16(v.reshape(v.shape[0], -1).contiguous(), planes)
    if kind == "dw":                                   # depthwise (C,1,3) -> (3,C) fp32
        return v[:, 0, :].t().contiguous()
    if kind == "conv3":                                # dense k=3 conv -> tap-major (3, Cout, Cin)
        return ops.split16(v.permute(2, 0, 1).contiguous(), planes)
    return v.reshape(-1).contiguous() if v.dim() != 2 else v.contiguous()


def pack_weights(sd, device, flat=None, params=None):
    """state_dict (reference keys, fp32) -> dict of device tensors in kernel operand formats.

    * conv / linear weights feeding a GEMM: bf16, K-major: (Cout, Cin); k=3 convs tap-major (3, Cout, Cin)
    * depthwise k=3 weights: fp32 (3, C) tap-major
    * LayerNorm / AffineDropPath / bias vectors: fp32 flat (C,)
    * XLNet (C, H, d) einsum weights: q,k,v,r transposed to (H*d, C) (K-major B operand), o as (C, H*d)

    flat (a trainer.FlatAdamW) + params (name -> Parameter): parameters owned by the flat optimizer are not copied — vectors
    are live fp32 views, same-layout GEMM weights are views of the optimizer's bf16 planes (kept current by the fused AdamW
    kernel); only the permuted layouts (k=3 convs, depthwise, XLNet q/k/v/r) are re-derived by `refresh_packed`.
    """
    W = {}
    repack = []
    for k, v in sd.items():
        if not torch.is_floating_point(v):
            continue
        if k.startswith("backbone.xlnet.word_embedding") or ".adapters." in k:
            continue  # dead embedding table; adapters are packed once under their pets.* / pets_emas.* names
        v = v.detach().to(device=device, dtype=f32)
        kind = _pack_kind(k, v)
        prm = params.get(k) if (flat is not None and params is not None) else None
        if prm is not None and id(prm) in flat.slots:
            if kind == "vec":
                W[k] = prm.data.reshape(-1) if prm.dim() != 2 else prm.data
                continue
            if kind == "gemm":
                W[k] = flat.plane_view(prm, (prm.shape[0], prm.numel() // prm.shape[0]))[:_planes_for(k)]
                continue
            repack.append((k, kind, prm))
        W[k] = _pack_one(kind, v, _planes_for(k))
    W["_repack"] = repack
    return W


def refresh_packed(W):
    """Re-derive the permuted copies after an in-place parameter update and drop everything cached from the old weights."""
    for k, kind, prm in W.get("_repack", ()):
        W[k] = _pack_one(kind, prm.data, _planes_for(k))
    W["_cache"] = {}


def sinusoid_pe_table(max_len, C, device):
    """get_sinusoid_encoding / sqrt(C) as a token-major (T, C) fp32 table — blocks.py:179-190, backbones.py:60-63."""
    pos = np.arange(max_len)[:, None].astype(np.float64)
    j = np.arange(C)[None, :]
    tab = pos / np.power(10000, 2 * (j // 2) / C)
    tab[:, 0::2] = np.sin(tab[:, 0::2])
    tab[:, 1::2] = np.cos(tab[:, 1::2])
    return (torch.from_numpy(tab.astype(np.float32)) / (C ** 0.5)).to(device).contiguous()


_POS_EMB = {}


SELF_ATTN_SP = os.environ.get("VILCO_SELF_ATTN_SP", "1") == "1"   # 0: the two-pass attn_fused kernel for every shape


def xlnet_pos_emb(T, C, device):
    """relative_positional_encoding (bi, no clamp) — modeling_xlnet_x.py:1029-1066 -> (2T, C) operand.  Constant per
    (T, C, device, operand mode): built and uploaded once (a per-step upload from pageable memory stalls the host until the
    stream has drained)."""
    key = (T, C, str(device), ops.precision())
    if key not in _POS_EMB:
        _POS_EMB[key] = _xlnet_pos_emb(T, C, device)
    return _POS_EMB[key]


def _xlnet_pos_emb(T, C, device):
    freq_seq = torch.arange(0, C, 2.0, dtype=torch.float)
    inv_freq = 1 / torch.pow(10000, (freq_seq / C))
    pos_seq = torch.arange(T, -T, -1.0)
    s = pos_seq[:, None] * inv_freq[None, :]
    return ops.split16(torch.cat([torch.sin(s), torch.cos(s)], dim=-1).to(device).contiguous())


# ----------------------------------------------------------------------------------------------------
# blocks
# ----------------------------------------------------------------------------------------------------
def mhca_fwd(W, pre, xin, mask, H, stride, window=-1, tlen=None):
    """MaskedMHCA.forward up to (not including) the output projection — blocks.py:351-402; LocalMaskedMHCA when
    window > 1 — blocks.py:1140-1200.  xin = ln1(x) (B,T,C) fp32.  Returns (attn_out bf16 (B,T/s,C), out_mask)."""
    B, T, C = xin.shape
    names = ("query", "key", "value")
    qc, kc, vc = ops.dwconv_ln(xin, mask, [W[pre + f"{n}_conv.conv.weight"] for n in names],
                               [W[pre + f"{n}_norm.weight"] for n in names],
                               [W[pre + f"{n}_norm.bias"] for n in names], stride, tlen=tlen)
    omask = mask[:, ::stride].contiguous() if stride > 1 else mask
    q = ops.linear(qc, W[pre + "query.weight"], bf16, bias=W[pre + "query.bias"])
    k = ops.linear(kc, W[pre + "key.weight"], bf16, bias=W[pre + "key.bias"])
    if window > 1:
        v = ops.linear(vc, W[pre + "value.weight"], bf16, bias=W[pre + "value.bias"])
        o = ops.local_attention(q, k, v, omask, H, window, W.get(pre + "rel_pe"))
        return o, omask
    # v * kv_mask (blocks.py:394) fused as the row multiplier of the value projection
    v = ops.linear(vc, W[pre + "value.weight"], bf16, bias=W[pre + "value.bias"], rowmul=omask.reshape(-1))
    if ops.FUSED_ATTN and C // H == 64 and k.shape[2] <= 2048:
        if SELF_ATTN_SP and q.shape == k.shape and v.shape[0] == 1 and ops.xl_attention_ok(q, q.shape[2], C, H):
            return ops.self_attention(q, k, v, omask, H, 1.0 / math.sqrt(C // H)), omask     # single-pass kernel
        return ops.attention(q, k, v, omask, H, 1.0 / math.sqrt(C // H)), omask
    S = ops.attn_scores(q, k, H, 1.0 / math.sqrt(C // H))
    P = ops.softmax_rows(S, omask, mode=0)
    o = ops.attn_pv(P, v, H, k.shape[2])
    return o, omask


def cross_attn_fwd(W, pre, x16, y16, ymask, H):
    """MaskedMHA.forward (cross attention) before the output projection — blocks.py:228-266."""
    C = x16.shape[-1]
    q = ops.linear(x16, W[pre + "query.weight"], bf16, bias=W[pre + "query.bias"])
    k = ops.linear(y16, W[pre + "key.weight"], bf16, bias=W[pre + "key.bias"])
    v = ops.linear(y16, W[pre + "value.weight"], bf16, bias=W[pre + "value.bias"], rowmul=ymask.reshape(-1))
    if ops.FUSED_ATTN and C // H == 64 and k.shape[2] <= 2048:
        return ops.attention(q, k, v, ymask, H, 1.0 / math.sqrt(C // H))
    S = ops.attn_scores(q, k, H, 1.0 / math.sqrt(C // H))
    P = ops.softmax_rows(S, ymask, mode=0)
    return ops.attn_pv(P, v, H, k.shape[2])


def channel_block_fwd(W, pre, ln1_32, ln1_16, H, tlen=None):
    """ChannelBlock.forward — blocks.py:423-466 on x = ln1(x) (no masking, norm1 unused).  Returns fp32 (B,T,C)."""
    qkv = ops.linear(ln1_16, W[pre + "attn.qkv.weight"], bf16, planes=ops.PLANES_HI)
    y = ops.channel_attention(qkv, H, tlen)
    x1 = ops.linear(y, W[pre + "attn.proj.weight"], f32, bias=W[pre + "attn.proj.bias"], resid=ln1_32)
    _, n2 = ops.layernorm(x1, W[pre + "norm2.weight"], W[pre + "norm2.bias"], 1e-5)
    h = ops.linear(n2, W[pre + "mlp.0.weight"], bf16, bias=W[pre + "mlp.0.bias"], act=ACT_GELU)
    return ops.linear(h, W[pre + "mlp.2.weight"], f32, bias=W[pre + "mlp.2.bias"], resid=x1)


def adapter_fwd(W, pre, ln1_32):
    """meta_archs.Adapter.layer: Linear over the TIME axis of (B,C,T): T -> 5T -> T/2 with GELU — meta_archs.py:105-148.
    ln1_32 (B,T,C) fp32 -> (B,T/2,C) fp32.  Two GEMMs over the transposed activations (rows = B*C, K = T)."""
    B, T, C = ln1_32.shape
    xt = ops.unpack(ln1_32)                                         # (B,C,T) fp32
    _, x16 = ops.axpby(xt, None, 1.0, 0.0, out32=False, out16=True)  # operand planes
    h = ops.linear(x16, W[pre + "layer.0.weight"], bf16, bias=W[pre + "layer.0.bias"], act=ACT_GELU)
    o = ops.linear(h, W[pre + "layer.2.weight"], f32, bias=W[pre + "layer.2.bias"])  # (B,C,T/2)
    return ops.unpack(o)                                            # (B,T/2,C): the same transpose kernel


def transformer_block_fwd(W, pre, x32, mask, H, stride, cross=None, t_c_alpha=0.8, window=-1, adapter_pre=None,
                          want16=False, tlen=None):
    """TransformerBlock.forward — blocks.py:561-593 (eval semantics).  x32 (B,T,C) fp32 residual stream.
    cross = (text32 (B,L,C), text_mask (B,L)) or None.  Returns (out32, out_mask[, out16])."""
    B, T, C = x32.shape
    # ln1_16 only feeds the channel-attention qkv projection (stride-1 blocks): split operand
    ln1_32, ln1_16 = ops.layernorm(x32, W[pre + "ln1.weight"], W[pre + "ln1.bias"], out32=True, out16=(stride == 1),
                                   planes=ops.PLANES_HI)
    o, omask = mhca_fwd(W, pre + "attn.", ln1_32, mask, H, stride, window, tlen)
    om = omask.reshape(-1)
    skip = x32 if stride == 1 else ops.maxpool3s2(x32)
    sa, sm = W.get(pre + "drop_path_attn.scale"), W.get(pre + "drop_path_mlp.scale")
    if adapter_pre is not None:
        # out = attn(ln1 x) + adapter(ln1 x) (adapter output is not masked, meta_archs.py:143-147);
        # h = skip*mask + sa*out = [skip*mask + sa*adapter] + sa*((proj(o)+b)*mask)
        ad = adapter_fwd(W, adapter_pre, ln1_32)
        r = ops.scale_add(skip, om, ad, sa)
        h = ops.linear(o, W[pre + "attn.proj.weight"], f32, bias=W[pre + "attn.proj.bias"], rowmul=om, colscale=sa, resid=r)
    else:
        # h = skip*mask + sa * ((proj(o) + b) * mask)      (blocks.py:404-405, 567)
        h = ops.linear(o, W[pre + "attn.proj.weight"], f32, bias=W[pre + "attn.proj.bias"], rowmul=om, colscale=sa,
                       resid=skip, resid_masked=True)
    if cross is not None and (pre + "cross_attn.query.weight") in W:
        text32, tmask = cross
        _, hx = ops.layernorm(h, W[pre + "ln3.weight"], W[pre + "ln3.bias"])
        _, hy = ops.layernorm(text32, W[pre + "ln3.weight"], W[pre + "ln3.bias"])
        c = cross_attn_fwd(W, pre + "cross_attn.", hx, hy, tmask, H)
        h = ops.linear(c, W[pre + "cross_attn.proj.weight"], f32, bias=W[pre + "cross_attn.proj.bias"], rowmul=om,
                       colscale=sa, resid=h, resid_masked=True)
    _, h2 = ops.layernorm(h, W[pre + "ln2.weight"], W[pre + "ln2.bias"])
    m = ops.linear(h2, W[pre + "mlp.0.weight"], bf16, bias=W[pre + "mlp.0.bias"], act=ACT_GELU)
    out = ops.linear(m, W[pre + "mlp.3.weight"], f32, bias=W[pre + "mlp.3.bias"], rowmul=om, colscale=sm, resid=h)
    out16 = None
    if stride == 1:
        out2 = channel_block_fwd(W, pre + "channel_attn.", ln1_32, ln1_16, H, tlen)
        out, out16 = ops.axpby(out, out2, t_c_alpha, 1.0 - t_c_alpha, out32=True, out16=want16)
    elif want16:
        _, out16 = ops.axpby(out, None, 1.0, 0.0, out32=False, out16=True)
    return (out, omask, out16) if want16 else (out, omask)


def xlnet_layer_fwd(W, pre, x32, x16, mask, H, eps=1e-12):
    """One XLNetLayer (inputs_embeds path) — modeling_xlnet_x.py:440-467, 270-332, 482-490, 1121-1283.
    x32 (B,T,C) fp32, x16 operand (NP,B,T,C), mask (B,T).  Returns fp32 (B,T,C)."""
    B, T, C = x32.shape
    d = C // H
    kq, kk, kv, ko, kr = (pre + "rel_attn." + n for n in "qkvor")
    rw = W[pre + "rel_attn.r_w_bias"].reshape(-1)
    rr = W[pre + "rel_attn.r_r_bias"].reshape(-1)
    qw = ops.linear(x16, W[kq], bf16, bias=rw)      # q + r_w_bias
    qr = ops.linear(x16, W[kq], bf16, bias=rr)      # q + r_r_bias
    k = ops.linear(x16, W[kk], bf16)
    v = ops.linear(x16, W[kv], bf16)
    # k_r = pos_emb @ W_r depends only on the weights and (T, B): cached with the packed weights (also keeps host->device
    # copies out of CUDA-graph capture)
    cache = W.setdefault("_cache", {})
    fused = ops.xl_attention_ok(qw, T, C, H)
    krel = cache.get(("krel", T, 1 if fused else B))
    if krel is None:
        pos = xlnet_pos_emb(T, C, x32.device)                               # (2T, C)
        krel = ops.linear(pos, W[kr], bf16)                                 # (NP, 2T, C)
        if not fused:
            krel = krel.unsqueeze(1).expand(-1, B, 2 * T, C).contiguous()
        cache[("krel", T, 1 if fused else B)] = krel
    if fused:
        # scores, relative shift, softmax and P V in one kernel: nothing of size T x T ever reaches HBM
        vec = ops.xl_attention(qw, qr, k, v, krel, mask, H, 1.0 / math.sqrt(d))
    else:
        ac = ops.attn_scores(qw, k, H, 1.0)                                     # (B,H,T,T)
        bd = ops.attn_scores(qr, krel, H, 1.0, band=(T, 2 * T))                 # (B,H,T,2T); only T <= i + p < 2T is read
        P = ops.softmax_rows(ac, mask, mode=1, BD=bd, scale=1.0 / math.sqrt(d))
        vec = ops.attn_pv(P, v, H, T)
    a = ops.linear(vec, W[ko], f32, resid=x32)                              # attn_out + h
    h1_32, h1_16 = ops.layernorm(a, W[pre + "rel_attn.layer_norm.weight"], W[pre + "rel_attn.layer_norm.bias"], eps,
                                 out32=True)
    f = ops.linear(h1_16, W[pre + "ff.layer_1.weight"], bf16, bias=W[pre + "ff.layer_1.bias"], act=ACT_GELU)
    f2 = ops.linear(f, W[pre + "ff.layer_2.weight"], f32, bias=W[pre + "ff.layer_2.bias"], resid=h1_32)
    h2, _ = ops.layernorm(f2, W[pre + "ff.layer_norm.weight"], W[pre + "ff.layer_norm.bias"], eps, out32=True,
                          out16=False)
    return h2


def backbone_fwd(W, cfg, x16, mask, text16=None, tmask=None, pe=None, pets_prefix="pets.", text_lens=None,
                 trunk_only=False):
    """ConvTransformerBackbone.forward — MQ/libs/modeling/backbones.py:181-289.
    x16 (B,T,Cin) bf16, mask (B,T) fp32, text16 (B,L,Ct) bf16, tmask (B,L) fp32.  Returns (feats fp32 list, masks).
    text_lens (B,) int32 or None: when given, every text sequence is treated as if it had been run alone (un-padded),
    which is what the reference's batch-1 evaluation does; None reproduces the reference's padded training batch."""
    pre = "backbone."
    _, B, T, _ = x16.shape
    C, H = cfg.embd_dim, cfg.n_head
    m = mask.reshape(-1)
    # input projection and embedding convolutions: split operands (x16 comes in ops.PLANES_HI planes)
    x = ops.linear(x16, W[pre + "proj.0.conv.weight"], bf16, bias=W[pre + "proj.0.conv.bias"], rowmul=m, planes=ops.PLANES_HI)
    x32 = None
    n_embd = cfg.arch[0]
    for i in range(n_embd):
        c = ops.conv3(x, W[pre + f"embd.{i}.conv.weight"], f32, rowmul=mask)
        last = i == n_embd - 1
        x32, x = ops.layernorm(c, W[pre + f"embd_norm.{i}.weight"], W[pre + f"embd_norm.{i}.bias"], relu=True,
                               pe=pe if last else None, rowmul=m if last else None, out32=last, out16=not last,
                               rows_per_batch=T, planes=ops.PLANES_HI)
    cross = None
    if cfg.use_cross_modal and text16 is not None:
        tm = tmask.reshape(-1)
        t = text16
        t32 = None
        for i in range(n_embd):
            c = ops.linear(t, W[pre + f"txt_embd.{i}.conv.weight"], f32, rowmul=tm)
            last = i == n_embd - 1
            t32, t = ops.layernorm(c, W[pre + f"txt_embd_norm.{i}.weight"], W[pre + f"txt_embd_norm.{i}.bias"],
                                   relu=True, out32=last, out16=not last)
        for i in range(cfg.arch[1]):
            t32, _ = transformer_block_fwd(W, pre + f"txt_stem.{i}.", t32, tmask, H, 1, t_c_alpha=0.8, tlen=text_lens)
        cross = (t32, tmask)
    x16s = None
    for i in range(cfg.arch[1]):
        want16 = cfg.use_xl and i == cfg.arch[1] - 1
        r = transformer_block_fwd(W, pre + f"stem.{i}.", x32, mask, H, 1, t_c_alpha=cfg.t_c_alpha, want16=want16)
        x32 = r[0]
        if want16:
            x16s = r[2]
    feat0 = x32
    if cfg.use_xl and cfg.arch[2] > 0:
        if x16s is None:
            _, x16s = ops.axpby(x32, None, 1.0, 0.0, out32=False, out16=True)
        x32 = xlnet_layer_fwd(W, pre + "xlnet.layer.0.", x32, x16s, mask, H)
    trunk = (feat0, x32, mask, cross)
    if trunk_only:
        return trunk
    return branch_fwd(W, cfg, trunk, pets_prefix)


def branch_fwd(W, cfg, trunk, pets_prefix="pets."):
    """The strided branch of the backbone (backbones.py:266-286) on top of the shared trunk (embedding, text path, stem,
    XLNet).  Adapters (`pets_prefix`) only live here, so the EMA-adapter ensemble of mq_vilco re-runs just this part."""
    pre = "backbone."
    feat0, x32, mask, cross = trunk
    feats, masks = [feat0], [mask]
    for i in range(cfg.arch[2]):
        cr = None if i in (1, 2) else cross
        ad = (pets_prefix + f"{list(cfg.adapt_blocks).index(i)}.") if i in cfg.adapt_blocks else None
        x32, mask = transformer_block_fwd(W, pre + f"branch.{i}.", x32, mask, cfg.n_head, cfg.scale_factor, cross=cr,
                                          t_c_alpha=cfg.t_c_alpha, adapter_pre=ad)
        feats.append(x32)
        masks.append(mask)
    return feats, masks


# ----------------------------------------------------------------------------------------------------
# neck + heads over the concatenated pyramid
# ----------------------------------------------------------------------------------------------------
class Pyramid:
    """Row layout of the concatenated pyramid buffer (B, P, C): level l occupies rows [off[l], off[l] + T_l), followed
    by one all-zero gap row so that the k=3 head convolutions never mix neighbouring levels."""

    def __init__(self, lens, device):
        self.lens = list(lens)
        self.off = []
        p = 0
        for n in self.lens:
            self.off.append(p)
            p += n + 1
        self.P = (p + 7) // 8 * 8
        gap = torch.ones(self.P, dtype=torch.uint8)
        for o, n in zip(self.off, self.lens):
            gap[o:o + n] = 0
        self.gap_rows = gap.to(device)  # 1 = gap (always zero) row

    _cache = {}

    @classmethod
    def cached(cls, lens, device):
        """one instance per (level lengths, device): the training step reuses the gap-row mask, the point table and the
        level one-hot instead of re-uploading them every step"""
        key = (tuple(int(n) for n in lens), str(device))
        if key not in cls._cache:
            cls._cache[key] = cls(lens, device)
        return cls._cache[key]


DENSE_RATES = (3, 6, 12, 18, 24)


def _dilated_conv_relu(W, key, x16, rate):
    """relu(dilated k=3 conv + bias) of the DenseAPP blocks on a token-major operand (NP, B, T, Cin): ONE GEMM over the three
    row-shifted copies concatenated along channels (K = 3 Cin) against the taps concatenated the same way; taps that fall
    entirely outside the sequence (rate >= T — every side tap at the MQ configuration, where the last level has T = 2)
    are dropped."""
    from . import backward as BW
    T = x16.shape[2]
    w3 = W[key + ".weight"]                                   # (NP, 3, Cout, Cin) tap-major
    cache = W.setdefault("_cache", {})
    if rate >= T:
        return ops.linear(x16, w3[:, 1], bf16, bias=W[key + ".bias"], act=ACT_RELU, planes=ops.PLANES_HI)
    ck = ("dil_cat", key)
    if ck not in cache:
        cache[ck] = w3.permute(0, 2, 1, 3).reshape(w3.shape[0], w3.shape[2], -1).contiguous()   # (NP, Cout, 3 Cin)
    xcat = torch.cat([BW.shift_planes(x16, -rate), x16, BW.shift_planes(x16, rate)], dim=-1)
    return ops.linear(xcat, cache[ck], bf16, bias=W[key + ".bias"], act=ACT_RELU, planes=ops.PLANES_HI)


def fpn1d_fwd(W, feats, masks, pre="neck."):
    """FPN1D.forward — MQ/libs/modeling/necks.py:64-106 (evaluation).  feats [(B, T_l, C) fp32], masks [(B, T_l) fp32]
    -> per-level operand tensors (NP, B, T_l, C) = LN(depthwise conv(top-down sum)).  The last level's lateral is ACConv =
    DenseAPP x mask (modeling/utils.py:692-751): 1x1 conv -> GroupNorm(32) -> ReLU -> dilated conv -> ReLU, five times over a
    growing concatenation, then 1x1 conv -> GroupNorm(32)."""
    n = len(feats)
    lat = []
    for i in range(n - 1):
        _, f16 = ops.axpby(feats[i], None, 1.0, 0.0, out32=False, out16=True)
        lat.append(ops.linear(f16, W[pre + f"lateral_convs.{i}.conv.weight"], f32, rowmul=masks[i].reshape(-1)))
    d = pre + "ac_conv.denseapp."
    # the DenseAPP stack keeps split operands in the mixed mode (sensitive_key: ".ac_conv."): it is five GroupNorm blocks deep
    # and runs on the shortest level only (T = 2 at the MQ configuration), so this costs nothing
    _, feature = ops.axpby(feats[-1], None, 1.0, 0.0, out32=False, out16=True, planes=ops.PLANES_HI)
    outs = []
    for r in DENSE_RATES:
        b = f"{d}aspp{r}."
        h32 = ops.linear(feature, W[b + "conv1x1.weight"], f32, bias=W[b + "conv1x1.bias"])
        _, h16 = ops.groupnorm(h32, W[b + "ConvGN.weight"], W[b + "ConvGN.bias"], 32, relu=True, out32=False, out16=True,
                               planes=ops.PLANES_HI)
        o16 = _dilated_conv_relu(W, b + "dilaconv", h16, r)
        outs.append(o16)
        feature = torch.cat([o16, feature], dim=-1)
    y32 = ops.linear(torch.cat(outs, dim=-1), W[d + "conv1x1.weight"], f32, bias=W[d + "conv1x1.bias"])
    y32, _ = ops.groupnorm(y32, W[d + "ConvGN.weight"], W[d + "ConvGN.bias"], 32, out32=True, out16=False)
    lat.append(y32 * masks[-1].unsqueeze(-1))
    for i in range(n - 1, 0, -1):
        ops.upsample2_add(lat[i], lat[i - 1])
    return [ops.dwconv_ln(lat[i].contiguous(), masks[i], [W[pre + f"fpn_convs.{i}.conv.weight"]], [W[pre + f"fpn_norms.{i}.weight"]],
                          [W[pre + f"fpn_norms.{i}.bias"]], 1)[0] for i in range(n)]


def neck_heads_fwd(W, cfg, feats, masks, pyr=None):
    """FPNIdentity.forward (necks.py:173-198) + PtTransformerClsHead / RegHead (meta_archs.py:259-275, 334-349) over
    all levels at once.  Returns (logits (B,P,K) fp32, offsets (B,P,2) fp32, pmask (B,P) fp32, pyr)."""
    B = feats[0].shape[0]
    C = cfg.embd_dim
    dev = feats[0].device
    if pyr is None:
        cache = W.setdefault("_cache", {})
        key = ("pyr",) + tuple(f.shape[1] for f in feats)
        pyr = cache.get(key)
        if pyr is None:
            pyr = cache[key] = Pyramid([f.shape[1] for f in feats], dev)
    P = pyr.P
    fpn = ops.zeros16(B, P, C, device=dev)
    pmask = torch.zeros(B, P, device=dev, dtype=f32)
    rowscale = torch.zeros(B, P, device=dev, dtype=f32)
    lv = fpn1d_fwd(W, feats, masks) if getattr(cfg, "fpn_type", "identity") == "fpn" else None
    for l, (f, mk) in enumerate(zip(feats, masks)):
        o, n = pyr.off[l], pyr.lens[l]
        if lv is not None:
            fpn[:, :, o:o + n] = lv[l]
        else:
            ops.layernorm(f, W[f"neck.fpn_norms.{l}.weight"], W[f"neck.fpn_norms.{l}.bias"], out32=False,
                          y16=fpn[:, :, o:o + n], y16_lo=ops.lo(fpn), rows_per_batch=n, y_ld=C, y_bs=P * C)
        pmask[:, o:o + n] = mk
        rowscale[:, o:o + n] = mk * W[f"reg_head.scale.{l}.scale"]
    zero_rows = pyr.gap_rows.repeat(B)
    outs = []
    for head, final_w, final_b in (("cls_head.", "cls_head.cls_head.conv", None), ("reg_head.", "reg_head.offset_head.conv", None)):
        x = fpn
        for i in range(2):
            # flat=True: the pyramid rows of all clips as ONE sequence — every clip ends in an all-zero, masked gap row, so
            # the k=3 taps never mix clips and no tile is cut at a clip boundary
            c = ops.conv3(x, W[head + f"head.{i}.conv.weight"], f32, rowmul=pmask, flat=True)
            _, x = ops.layernorm(c, W[head + f"norm.{i}.weight"], W[head + f"norm.{i}.bias"], relu=True,
                                 zero_rows=zero_rows)
        if head == "cls_head.":
            outs.append(ops.conv3(x, W[final_w + ".weight"], f32, bias=W[final_w + ".bias"], rowmul=pmask, flat=True))
        else:  # relu(scale_l * (conv + b) * mask)
            outs.append(ops.conv3(x, W[final_w + ".weight"], f32, bias=W[final_w + ".bias"], rowmul=rowscale, act=ACT_RELU,
                                  flat=True))
    return outs[0], outs[1], pmask, pyr
