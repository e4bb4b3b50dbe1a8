"""TEST INFRASTRUCTURE ONLY — CPU (torch fp32) restatement of ViLCo's Moment-Query hot path.

This file is the *oracle* the CUDA path is checked against.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / `--impl reference` legs may import it; the product (vilco_b200/) never does.

It restates, function by function, what the reference computes (file:line cited per function, paths
relative to the ViLCo repo root) in the reference's own tensor layout (B, C, T), eval-mode semantics
(dropout / drop-path are identity; `training=True` only switches the positional-encoding branch and the
loss path).  It is written functionally over a flat ``params`` dict that uses the reference's state_dict
key names, so the same seeded state dict can be loaded into the reference, the
oracle and the CUDA path.

Pinning: oracle/gen_golden.py runs the *reference itself* (imported from /root/reference through
oracle/ref_shim.py) on seeded inputs and commits the outputs under tests/golden/; tests/test_oracle_golden.py
checks this file against those vectors on every box, and tests/test_oracle_vs_reference.py checks it
against the live reference when /root/reference is present.  The reference has no tests or golden vectors
of its own, so these generated vectors are the pin.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------------
# operators
# ----------------------------------------------------------------------------------------------------
def masked_conv1d(x, mask, weight, bias=None, stride=1, groups=1):
    """MaskedConv1D.forward — MQ/libs/modeling/blocks.py:106-130.  x (B,C,T), mask (B,1,T) bool."""
    k = weight.shape[-1]
    T = x.shape[-1]
    assert T % stride == 0
    out = F.conv1d(x, weight, bias, stride=stride, padding=k // 2, groups=groups)
    if stride > 1:
        # nearest interpolation to T//stride == take every stride-th mask entry
        out_mask = mask[:, :, ::stride]
    else:
        out_mask = mask
    return out * out_mask.to(x.dtype), out_mask


def channel_layernorm(x, weight, bias, eps=1e-5):
    """LayerNorm.forward over dim 1 of (B,C,T) — MQ/libs/modeling/blocks.py:160-175 (biased variance)."""
    mu = x.mean(dim=1, keepdim=True)
    r = x - mu
    var = (r * r).mean(dim=1, keepdim=True)
    return r / torch.sqrt(var + eps) * weight + bias


def _heads(x, n_head):
    B, C, T = x.shape
    return x.view(B, n_head, C // n_head, T).transpose(2, 3)  # (B, H, T, d)


def _qkv_conv_norm(P, pre, x, mask, stride):
    """depthwise conv + LN for q, k, v — blocks.py:364-371 (q conv uses the kv stride, blocks.py:313)."""
    C = x.shape[1]
    q, qm = masked_conv1d(x, mask, P[pre + "query_conv.conv.weight"], None, stride, C)
    q = channel_layernorm(q, P[pre + "query_norm.weight"], P[pre + "query_norm.bias"])
    k, km = masked_conv1d(x, mask, P[pre + "key_conv.conv.weight"], None, stride, C)
    k = channel_layernorm(k, P[pre + "key_norm.weight"], P[pre + "key_norm.bias"])
    v, _ = masked_conv1d(x, mask, P[pre + "value_conv.conv.weight"], None, stride, C)
    v = channel_layernorm(v, P[pre + "value_norm.weight"], P[pre + "value_norm.bias"])
    q = F.conv1d(q, P[pre + "query.weight"], P[pre + "query.bias"])
    k = F.conv1d(k, P[pre + "key.weight"], P[pre + "key.bias"])
    v = F.conv1d(v, P[pre + "value.weight"], P[pre + "value.bias"])
    return q, k, v, qm, km


def masked_mhca(P, pre, x, mask, n_head, stride=1):
    """MaskedMHCA.forward (global attention, the window == -1 variant) — blocks.py:351-410."""
    B, C, T = x.shape
    q, k, v, qm, km = _qkv_conv_norm(P, pre, x, mask, stride)
    scale = 1.0 / math.sqrt(C // n_head)
    q, k, v = _heads(q, n_head), _heads(k, n_head), _heads(v, n_head)
    att = (q * scale) @ k.transpose(-2, -1)
    att = att.masked_fill(~km[:, :, None, :], float("-inf"))
    att = torch.softmax(att, dim=-1)
    out = att @ (v * km[:, :, :, None].to(v.dtype))
    out = out.transpose(2, 3).contiguous().view(B, C, -1)
    out = F.conv1d(out, P[pre + "proj.weight"], P[pre + "proj.bias"]) * qm.to(out.dtype)
    return out, qm


def local_masked_mhca(P, pre, x, mask, n_head, window, stride=1, rel_pe=None):
    """LocalMaskedMHCA.forward — blocks.py:1140-1207, restated as a plain banded attention.

    Query i sees keys j in [i-w, i+w] (w = window // 2).  Positions outside the sequence get -inf
    (`_mask_invalid_locations`, blocks.py:994-1006), padded keys get -1e4 added (blocks.py:1175-1187),
    rows of padded queries are zeroed after the softmax (blocks.py:1192-1194).
    """
    B, C, T = x.shape
    w = window // 2
    q, k, v, qm, km = _qkv_conv_norm(P, pre, x, mask, stride)
    d = C // n_head
    q, k, v = _heads(q, n_head) * (1.0 / math.sqrt(d)), _heads(k, n_head), _heads(v, n_head)
    Tq = q.shape[2]
    att = q @ k.transpose(-2, -1)  # (B,H,Tq,Tq) — only the band is used
    idx = torch.arange(Tq)
    rel = idx[None, :] - idx[:, None]  # j - i
    band = rel.abs() <= w
    if rel_pe is not None:  # (1,1,H,W) added per diagonal (blocks.py:1172-1173)
        pe = rel_pe.view(n_head, window)[:, (rel.clamp(-w, w) + w)]  # (H,Tq,Tq)
        att = att + pe[None]
    att = att + (~km[:, :, None, :]).to(att.dtype) * -1e4
    att = att.masked_fill(~band[None, None], float("-inf"))
    att = torch.softmax(att, dim=-1)
    att = att.masked_fill(~km[:, 0, :][:, None, :, None], 0.0)
    out = att @ v
    out = out.transpose(2, 3).contiguous().view(B, C, -1)
    out = F.conv1d(out, P[pre + "proj.weight"], P[pre + "proj.bias"]) * qm.to(out.dtype)
    return out, qm


def masked_mha_cross(P, pre, x, mask_f, y, y_mask, n_head):
    """MaskedMHA.forward, cross-attention branch — blocks.py:228-269.  y_mask (B,L) long/bool."""
    B, C, T = x.shape
    q = F.conv1d(x, P[pre + "query.weight"], P[pre + "query.bias"])
    k = F.conv1d(y, P[pre + "key.weight"], P[pre + "key.bias"])
    v = F.conv1d(y, P[pre + "value.weight"], P[pre + "value.bias"])
    scale = 1.0 / math.sqrt(C // n_head)
    q, k, v = _heads(q, n_head), _heads(k, n_head), _heads(v, n_head)
    am = y_mask.bool()
    att = (q * scale) @ k.transpose(-2, -1)
    att = att.masked_fill(~am[:, None, None, :], float("-inf"))
    att = torch.softmax(att, dim=-1)
    out = att @ (v * am[:, None, :, None].to(v.dtype))
    out = out.transpose(2, 3).contiguous().view(B, C, -1)
    out = F.conv1d(out, P[pre + "proj.weight"], P[pre + "proj.bias"]) * mask_f
    return out


def channel_block(P, pre, x, n_head):
    """ChannelBlock.forward + ChannelAttention.forward — blocks.py:423-466 (no masking; norm1 unused)."""
    x = x.permute(0, 2, 1)  # (B,T,C)
    B, T, C = x.shape
    d = C // n_head
    qkv = F.linear(x, P[pre + "attn.qkv.weight"]).reshape(B, T, 3, n_head, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]  # (B,H,T,d)
    a = torch.softmax((k * d ** -0.5).transpose(-1, -2) @ v, dim=-1)  # (B,H,d,d)
    y = (a @ q.transpose(-1, -2)).transpose(-1, -2)  # (B,H,T,d)
    y = y.transpose(1, 2).reshape(B, T, C)
    y = F.linear(y, P[pre + "attn.proj.weight"], P[pre + "attn.proj.bias"])
    x = x + y
    h = F.layer_norm(x, (C,), P[pre + "norm2.weight"], P[pre + "norm2.bias"], 1e-5)
    h = F.linear(F.gelu(F.linear(h, P[pre + "mlp.0.weight"], P[pre + "mlp.0.bias"])),
                 P[pre + "mlp.2.weight"], P[pre + "mlp.2.bias"])
    x = x + h
    return x.permute(0, 2, 1)


def adapter_time(P, pre, x):
    """meta_archs.Adapter.layer on (B,C,T): Linear over the TIME axis T -> 5T -> T/2 — meta_archs.py:105-148."""
    h = F.gelu(F.linear(x, P[pre + "layer.0.weight"], P[pre + "layer.0.bias"]))
    return F.linear(h, P[pre + "layer.2.weight"], P[pre + "layer.2.bias"])


def transformer_block(P, pre, x, mask, n_head, stride, cross_y=None, cross_y_mask=None, t_c_alpha=0.8,
                      window=-1, adapter_pre=None, channel_mix=True):
    """TransformerBlock.forward — blocks.py:561-593 (eval: AffineDropPath == per-channel scale, blocks.py:655-670)."""
    ln1 = channel_layernorm(x, P[pre + "ln1.weight"], P[pre + "ln1.bias"])
    if window > 1:
        out, out_mask = local_masked_mhca(P, pre + "attn.", ln1, mask, n_head, window, stride,
                                          P.get(pre + "attn.rel_pe"))
    else:
        out, out_mask = masked_mhca(P, pre + "attn.", ln1, mask, n_head, stride)
    if adapter_pre is not None:  # parallel adapter, blocks.py:45-54 + meta_archs.py:139-148
        out = out + adapter_time(P, adapter_pre, ln1)
    mf = out_mask.to(out.dtype)
    skip = x if stride == 1 else F.max_pool1d(x, stride + 1, stride, (stride + 1) // 2)
    sa = P.get(pre + "drop_path_attn.scale")
    sm = P.get(pre + "drop_path_mlp.scale")
    out = skip * mf + (out if sa is None else sa * out)
    if cross_y is not None and (pre + "cross_attn.query.weight") in P:
        ln3x = channel_layernorm(out, P[pre + "ln3.weight"], P[pre + "ln3.bias"])
        ln3y = channel_layernorm(cross_y, P[pre + "ln3.weight"], P[pre + "ln3.bias"])
        c = masked_mha_cross(P, pre + "cross_attn.", ln3x, mf, ln3y, cross_y_mask, n_head)
        out = out * mf + (c if sa is None else sa * c)
    h = channel_layernorm(out, P[pre + "ln2.weight"], P[pre + "ln2.bias"])
    h = F.conv1d(F.gelu(F.conv1d(h, P[pre + "mlp.0.weight"], P[pre + "mlp.0.bias"])),
                 P[pre + "mlp.3.weight"], P[pre + "mlp.3.bias"]) * mf
    out = out + (h if sm is None else sm * h)
    if stride == 1 and channel_mix:   # MQ only; the NLQ block (NLQ/libs/modeling/blocks.py:840-874) has no channel mix
        out2 = channel_block(P, pre + "channel_attn.", ln1, n_head)
        out = t_c_alpha * out + (1 - t_c_alpha) * out2
    return out, out_mask


def xlnet_pos_emb(T, d_model):
    """relative_positional_encoding, attn_type 'bi', bi_data False, clamp_len -1 — modeling_xlnet_x.py:1029-1066."""
    freq_seq = torch.arange(0, d_model, 2.0, dtype=torch.float)
    inv_freq = 1 / torch.pow(10000, (freq_seq / d_model))
    pos_seq = torch.arange(T, -T, -1.0)
    s = pos_seq[:, None] * inv_freq[None, :]
    return torch.cat([torch.sin(s), torch.cos(s)], dim=-1)  # (2T, d_model)


def xlnet_layer(P, pre, x_btc, attn_mask_bt, eps=1e-12):
    """One XLNetLayer as MQ runs it (inputs_embeds path, bi-directional, no mems/segments) —
    modeling_xlnet_x.py:1121-1283 (mask prep), :440-467 + :270-332 (rel attention), :482-490 (FF).
    x_btc (B,T,C); attn_mask_bt (B,T) with 1 = valid.  Returns (B,T,C)."""
    B, T, C = x_btc.shape
    Wq, Wk, Wv, Wo, Wr = (P[pre + "rel_attn." + n] for n in "qkvor")
    H, d = Wq.shape[1], Wq.shape[2]
    h = x_btc.transpose(0, 1)  # (T,B,C)
    q = torch.einsum("ibh,hnd->ibnd", h, Wq)
    k = torch.einsum("ibh,hnd->ibnd", h, Wk)
    v = torch.einsum("ibh,hnd->ibnd", h, Wv)
    kr = torch.einsum("ph,hnd->pnd", xlnet_pos_emb(T, C), Wr)  # (2T,H,d)
    ac = torch.einsum("ibnd,jbnd->bnij", q + P[pre + "rel_attn.r_w_bias"], k)
    bd_raw = torch.einsum("ibnd,pnd->bnip", q + P[pre + "rel_attn.r_r_bias"], kr)  # (B,H,T,2T)
    i = torch.arange(T)
    gather = (T + i[None, :] - i[:, None])  # rel_shift_bnij (:256-268): bd[i,j] = bd_raw[i, T + j - i]
    bd = bd_raw.gather(3, gather[None, None].expand(B, H, T, T))
    score = (ac + bd) * (1.0 / math.sqrt(d))
    pad = (1.0 - attn_mask_bt.to(score.dtype))  # (B,T) 1 = padding key
    m = pad[:, None, None, :].expand(B, 1, T, T).clone()
    m[:, :, i, i] = 0.0  # a token can always attend to itself (:1184-1188)
    score = score - 1e30 * (m > 0).to(score.dtype)
    prob = torch.softmax(score, dim=3)
    vec = torch.einsum("bnij,jbnd->ibnd", prob, v)
    attn_out = torch.einsum("ibnd,hnd->ibh", vec, Wo)
    h1 = F.layer_norm(attn_out + h, (C,), P[pre + "rel_attn.layer_norm.weight"], P[pre + "rel_attn.layer_norm.bias"], eps)
    f = F.linear(F.gelu(F.linear(h1, P[pre + "ff.layer_1.weight"], P[pre + "ff.layer_1.bias"])),
                 P[pre + "ff.layer_2.weight"], P[pre + "ff.layer_2.bias"])
    h2 = F.layer_norm(f + h1, (C,), P[pre + "ff.layer_norm.weight"], P[pre + "ff.layer_norm.bias"], eps)
    return h2.transpose(0, 1).contiguous()


def sinusoid_pe(n_position, d_hid):
    """get_sinusoid_encoding — blocks.py:179-190, returned as (1, C, T); caller scales by 1/sqrt(C) (backbones.py:62)."""
    pos = np.arange(n_position)[:, None].astype(np.float64)
    j = np.arange(d_hid)[None, :]
    tab = pos / np.power(10000, 2 * (j // 2) / d_hid)
    tab[:, 0::2] = np.sin(tab[:, 0::2])
    tab[:, 1::2] = np.cos(tab[:, 1::2])
    return torch.FloatTensor(tab).unsqueeze(0).transpose(1, 2)


# ----------------------------------------------------------------------------------------------------
# model
# ----------------------------------------------------------------------------------------------------
class ModelCfg:
    """The handful of MQ/libs/core/config.py fields the path depends on (defaults = mq_no_cl.yaml)."""

    def __init__(self, **kw):
        self.input_dim = 4096
        self.embd_dim = 1024
        self.n_head = 16
        self.max_seq_len = 1024
        self.arch = (2, 2, 9)
        self.scale_factor = 2
        self.num_classes = 22
        self.n_txt_in = 768
        self.use_cross_modal = True
        self.use_xl = True
        self.t_c_alpha = 0.8
        self.fpn_type = "identity"      # "fpn": FPN1D + ACConv / DenseAPP (necks.py:13-106; no config selects it)
        self.regression_range = [[0, 4], [2, 8], [4, 16], [8, 32], [16, 64], [32, 128], [64, 256], [128, 512],
                                 [256, 1024], [512, 10000]]
        self.center_sample_radius = 1.5
        self.init_loss_norm = 100.0
        self.loss_weight = 1.0
        self.al_loss_weight = 0.2
        self.pre_nms_thresh = 0.001
        self.pre_nms_topk = 5000
        self.iou_threshold = 0.1
        self.min_score = 0.0001
        self.max_seg_num = 200
        self.nms_sigma = 0.99
        self.duration_thresh = 0.01
        self.adapt_blocks = ()  # vilco: (0,1,2,3,4)
        self.prompt_pool = None  # vilco: dict(pool_size=10, top_k=4, length=20)
        self.n_emas = 0          # vilco: 1 (EMA copy of the adapters, ensembled at inference)
        self.narration_dim = 0   # vilco training: 512 (narration SSL branch; only the parameter spec uses it here)
        for k, v in kw.items():
            assert hasattr(self, k), k
            setattr(self, k, v)

    @property
    def n_levels(self):
        return self.arch[2] + 1

    @property
    def strides(self):
        return [self.scale_factor ** i for i in range(self.n_levels)]


def backbone(P, cfg, x, mask, text=None, text_mask=None, training=False, pets_prefix="pets."):
    """ConvTransformerBackbone.forward — MQ/libs/modeling/backbones.py:181-289."""
    pre = "backbone."
    x, mask = masked_conv1d(x, mask, P[pre + "proj.0.conv.weight"], P[pre + "proj.0.conv.bias"])  # :185-190
    for i in range(cfg.arch[0]):  # :217-219
        x, mask = masked_conv1d(x, mask, P[pre + f"embd.{i}.conv.weight"], None)
        x = F.relu(channel_layernorm(x, P[pre + f"embd_norm.{i}.weight"], P[pre + f"embd_norm.{i}.bias"]))
    T = x.shape[-1]
    pe = sinusoid_pe(cfg.max_seq_len, cfg.embd_dim) / (cfg.embd_dim ** 0.5)
    if (not training) and T >= cfg.max_seq_len:  # :229-236
        pe = F.interpolate(pe, T, mode="linear", align_corners=False)
    x = x + pe[:, :, :T] * mask.to(x.dtype)
    q, qm = None, None
    if cfg.use_cross_modal and text is not None:  # :242-252
        tm = text_mask
        for i in range(cfg.arch[0]):
            text, tm = masked_conv1d(text, tm, P[pre + f"txt_embd.{i}.conv.weight"], None)
            text = F.relu(channel_layernorm(text, P[pre + f"txt_embd_norm.{i}.weight"], P[pre + f"txt_embd_norm.{i}.bias"]))
        q, qm = text, tm
        for i in range(cfg.arch[1]):
            q, qm = transformer_block(P, pre + f"txt_stem.{i}.", q, qm, cfg.n_head, 1, t_c_alpha=0.8)
        qm = qm.squeeze(1).long()
    for i in range(cfg.arch[1]):  # :255-256 (no cross attention in the stem)
        x, mask = transformer_block(P, pre + f"stem.{i}.", x, mask, cfg.n_head, 1, t_c_alpha=cfg.t_c_alpha)
    feats, masks = [x], [mask]
    for i in range(cfg.arch[2]):  # :266-286
        if cfg.use_xl and i == 0:
            x = xlnet_layer(P, pre + "xlnet.layer.0.", x.permute(0, 2, 1), mask.squeeze(1).long()).permute(0, 2, 1)
        cy, cm = (None, None) if i in (1, 2) else (q, qm)
        ad = (pets_prefix + f"{cfg.adapt_blocks.index(i)}.") if i in cfg.adapt_blocks else None
        x, mask = transformer_block(P, pre + f"branch.{i}.", x, mask, cfg.n_head, cfg.scale_factor, cy, cm,
                                    t_c_alpha=cfg.t_c_alpha, adapter_pre=ad)
        feats.append(x)
        masks.append(mask)
    return feats, masks


def fpn_identity(P, feats, masks):
    """FPNIdentity.forward — MQ/libs/modeling/necks.py:173-198."""
    return [channel_layernorm(f, P[f"neck.fpn_norms.{i}.weight"], P[f"neck.fpn_norms.{i}.bias"])
            for i, f in enumerate(feats)], masks


DENSE_RATES = (3, 6, 12, 18, 24)


def dense_app(P, pre, x):
    """DenseAPP.forward — MQ/libs/modeling/utils.py:692-729 (evaluation: dropout off).  Five DenseBlocks (:671-689: 1x1 conv ->
    GroupNorm(32) -> ReLU -> dilated k=3 conv, rate r -> ReLU), each fed the concatenation [newest, ..., oldest, input]; the five
    outputs are concatenated, 1x1 conv, GroupNorm(32).  Plain (unmasked) convolutions, as in the reference."""
    feature, outs = x, []
    for r in DENSE_RATES:
        b = f"{pre}aspp{r}."
        h = F.conv1d(feature, P[b + "conv1x1.weight"], P[b + "conv1x1.bias"])
        h = F.relu(F.group_norm(h, 32, P[b + "ConvGN.weight"], P[b + "ConvGN.bias"]))
        h = F.relu(F.conv1d(h, P[b + "dilaconv.weight"], P[b + "dilaconv.bias"], padding=r, dilation=r))
        outs.append(h)
        feature = torch.cat([h, feature], dim=1)
    y = F.conv1d(torch.cat(outs, dim=1), P[pre + "conv1x1.weight"], P[pre + "conv1x1.bias"])
    return F.group_norm(y, 32, P[pre + "ConvGN.weight"], P[pre + "ConvGN.bias"])


def fpn1d(P, feats, masks, pre="neck."):
    """FPN1D.forward — MQ/libs/modeling/necks.py:64-106: 1x1 lateral convs (the LAST level goes through ACConv = DenseAPP x mask
    instead, utils.py:732-751; its CxAM / CnAM branches are commented out in the reference), top-down nearest x2 upsample-add,
    depthwise k=3 conv + channel LayerNorm per level.  feats [(B, C, T_l)], masks [(B, 1, T_l) bool]."""
    n = len(feats)
    lat = []
    for i in range(n):
        if i == n - 1:
            lat.append(dense_app(P, pre + "ac_conv.denseapp.", feats[-1]) * masks[i].to(feats[-1].dtype))
        else:
            lat.append(masked_conv1d(feats[i], masks[i], P[pre + f"lateral_convs.{i}.conv.weight"], None)[0])
    for i in range(n - 1, 0, -1):
        lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], scale_factor=2.0, mode="nearest")
    out = []
    for i in range(n):
        C = lat[i].shape[1]
        x, _ = masked_conv1d(lat[i], masks[i], P[pre + f"fpn_convs.{i}.conv.weight"], None, groups=C)
        out.append(channel_layernorm(x, P[pre + f"fpn_norms.{i}.weight"], P[pre + f"fpn_norms.{i}.bias"]))
    return out, masks


def _head_tower(P, pre, x, mask):
    for i in range(2):
        x, _ = masked_conv1d(x, mask, P[pre + f"head.{i}.conv.weight"], None)
        x = F.relu(channel_layernorm(x, P[pre + f"norm.{i}.weight"], P[pre + f"norm.{i}.bias"]))
    return x


def cls_head(P, feats, masks):
    """PtTransformerClsHead.forward — MQ/libs/modeling/meta_archs.py:259-275 -> list of (B,K,T_l)."""
    return [masked_conv1d(_head_tower(P, "cls_head.", f, m), m, P["cls_head.cls_head.conv.weight"],
                          P["cls_head.cls_head.conv.bias"])[0] for f, m in zip(feats, masks)]


def reg_head(P, feats, masks):
    """PtTransformerRegHead.forward — meta_archs.py:334-349 -> list of (B,2,T_l) = relu(scale_l * conv)."""
    out = []
    for l, (f, m) in enumerate(zip(feats, masks)):
        o, _ = masked_conv1d(_head_tower(P, "reg_head.", f, m), m, P["reg_head.offset_head.conv.weight"],
                             P["reg_head.offset_head.conv.bias"])
        out.append(F.relu(o * P[f"reg_head.scale.{l}.scale"]))
    return out


def points(cfg, lens):
    """PointGenerator — MQ/libs/modeling/loc_generators.py:59-92: rows [t, reg_lo, reg_hi, stride] per level."""
    out = []
    for l, (stride, n) in enumerate(zip(cfg.strides, lens)):
        t = torch.arange(0, cfg.max_seq_len * 64, stride, dtype=torch.float)[:n, None]
        rr = torch.tensor(cfg.regression_range[l], dtype=torch.float)[None].repeat(n, 1)
        st = torch.full((n, 1), float(stride))
        out.append(torch.cat((t, rr, st), dim=1))
    return out


def prompt_prepend(P, cfg, text_bcl):
    """Prompt.forward with prompt_mask None (evaluation) — MQ/libs/cl_methods/prompt.py:46-137 (embedding_key 'mean',
    batchwise_prompt).  text (B, Ct, L) -> (B, Ct, top_k*length + L)."""
    pp = cfg.prompt_pool
    x = text_bcl.permute(0, 2, 1)
    nrm = lambda v: v * torch.rsqrt(torch.maximum((v ** 2).sum(1, keepdim=True), torch.tensor(1e-12)))  # noqa: E731
    sim = nrm(x.mean(dim=1)) @ nrm(P["prompt.prompt_key"]).t()
    _, idx = torch.topk(sim, k=pp["top_k"], dim=1)
    pid, cnt = torch.unique(idx, return_counts=True, sorted=True)
    if pid.shape[0] < pp["pool_size"]:
        pad = pp["pool_size"] - pid.shape[0]
        pid = torch.cat([pid, torch.full((pad,), int(idx.min()))])
        cnt = torch.cat([cnt, torch.zeros(pad, dtype=cnt.dtype)])
    _, major = torch.topk(cnt, k=pp["top_k"])
    idx = pid[major].expand(x.shape[0], -1)
    bp = P["prompt.prompt"][idx].reshape(x.shape[0], -1, x.shape[2])
    return torch.cat([bp, x], dim=1).permute(0, 2, 1)


def forward_heads(P, cfg, feats_bct, mask_b1t, text=None, text_mask=None, training=False):
    """backbone -> neck -> heads (+ the EMA-adapter ensemble of mq_vilco at inference, meta_archs.py:854-881)."""
    out = _forward_heads_once(P, cfg, feats_bct, mask_b1t, text, text_mask, training, "pets.")
    if not training and cfg.adapt_blocks:
        for e in range(cfg.n_emas):
            o2 = _forward_heads_once(P, cfg, feats_bct, mask_b1t, text, text_mask, training, f"pets_emas.{e}.module.")
            out = ([(a + b) / 2 for a, b in zip(out[0], o2[0])], [(a + b) / 2 for a, b in zip(out[1], o2[1])], out[2], out[3])
    return out


def _forward_heads_once(P, cfg, feats_bct, mask_b1t, text=None, text_mask=None, training=False, pets_prefix="pets."):
    """backbone -> neck -> heads; returns lists permuted like meta_archs.py:848-852:
    logits (B,T_l,K), offsets (B,T_l,2), masks (B,T_l)."""
    feats, masks = backbone(P, cfg, feats_bct, mask_b1t, text, text_mask, training, pets_prefix)
    fpn, masks = fpn1d(P, feats, masks) if getattr(cfg, "fpn_type", "identity") == "fpn" else fpn_identity(P, feats, masks)
    offs = reg_head(P, fpn, masks)
    logits = cls_head(P, fpn, masks)
    return ([x.permute(0, 2, 1) for x in logits], [x.permute(0, 2, 1) for x in offs], [m.squeeze(1) for m in masks],
            fpn)


# ----------------------------------------------------------------------------------------------------
# targets and losses
# ----------------------------------------------------------------------------------------------------
def _normal(x, mu, sigma):
    return (-(x - mu) ** 2 / (2 * sigma ** 2)).exp()  # meta_archs.py:20-21


def label_points_single_video(P, cfg, concat_points, gt_segment, gt_label):
    """label_points_single_video — meta_archs.py:1253-1344 (center_sample == 'radius')."""
    num_pts, num_gts = concat_points.shape[0], gt_segment.shape[0]
    K = P["mu"].shape[0]
    lens = (gt_segment[:, 1] - gt_segment[:, 0])[None, :].repeat(num_pts, 1)
    gt_segs = gt_segment[None].expand(num_pts, num_gts, 2)
    t = concat_points[:, 0, None]
    stride = concat_points[:, 3, None]
    left = t - gt_segs[:, :, 0]
    right = gt_segs[:, :, 1] - t
    xrel = ((right - left) / 2.0) / (stride * lens)
    g = lambda m, s: _normal(xrel, P[m][gt_label].permute(1, 0), P[s][gt_label].permute(1, 0))  # noqa: E731
    npc, npl, npr = g("mu", "sigma"), g("mu_reg_left", "sigma_reg_left"), g("mu_reg_right", "sigma_reg_right")
    reg_targets = torch.stack((left, right), dim=-1)
    center = 0.5 * (gt_segs[:, :, 0] + gt_segs[:, :, 1])
    t_mins = center - stride * cfg.center_sample_radius
    t_maxs = center + stride * cfg.center_sample_radius
    cb_l = t - torch.maximum(t_mins, gt_segs[:, :, 0])
    cb_r = torch.minimum(t_maxs, gt_segs[:, :, 1]) - t
    inside = torch.stack((cb_l, cb_r), -1).min(-1)[0] > 0
    maxreg = reg_targets.max(-1)[0]
    in_range = (maxreg >= concat_points[:, 1, None]) & (maxreg <= concat_points[:, 2, None])
    lens = lens.masked_fill(~inside, float("inf")).masked_fill(~in_range, float("inf"))
    min_len, min_inds = lens.min(dim=1)
    min_len_mask = ((lens <= (min_len[:, None] + 1e-3)) & (lens < float("inf"))).to(reg_targets.dtype)
    onehot = F.one_hot(gt_label, K).to(reg_targets.dtype)
    cls_targets = (min_len_mask @ onehot).clamp(min=0.0, max=1.0)
    r = torch.arange(num_pts)
    reg_t = reg_targets[r, min_inds] / stride
    return cls_targets, reg_t, (npc[r, min_inds], npl[r, min_inds], npr[r, min_inds])


def sigmoid_focal_loss(inputs, targets, alpha=0.25, gamma=2.0):
    """sigmoid_focal_loss, reduction none — MQ/libs/modeling/losses.py:5-51."""
    p = torch.sigmoid(inputs)
    ce = F.binary_cross_entropy_with_logits(inputs, targets, reduction="none")
    p_t = p * targets + (1 - p) * (1 - targets)
    loss = ce * ((1 - p_t) ** gamma)
    return (alpha * targets + (1 - alpha) * (1 - targets)) * loss


def ctr_diou_loss_1d(inp, tgt, eps=1e-8):
    """ctr_diou_loss_1d, reduction none — losses.py:109-168."""
    lp, rp, lg, rg = inp[:, 0], inp[:, 1], tgt[:, 0], tgt[:, 1]
    inter = torch.min(lp, lg) + torch.min(rp, rg)
    union = (lp + rp) + (lg + rg) - inter
    iou = inter / union.clamp(min=eps)
    len_c = torch.max(lp, lg) + torch.max(rp, rg)
    rho = 0.5 * (rp - lp - rg + lg)
    return 1.0 - iou + torch.square(rho / len_c.clamp(min=eps))


def losses(P, cfg, fpn_masks, out_cls_logits, out_offsets, gt_segments, gt_labels, loss_normalizer):
    """label_points + PtTransformer.losses — meta_archs.py:1224-1251, 1374-1480 (no CL distillation terms).
    Returns (dict of losses, updated loss_normalizer)."""
    pts = torch.cat(points(cfg, [m.shape[1] for m in fpn_masks]), dim=0)
    gt_cls, gt_off, npc, npl, npr = [], [], [], [], []
    for seg, lab in zip(gt_segments, gt_labels):
        c, r, (a, b, d) = label_points_single_video(P, cfg, pts, seg, lab)
        gt_cls.append(c.detach()); gt_off.append(r.detach()); npc.append(a); npl.append(b); npr.append(d)
    valid = torch.cat(fpn_masks, dim=1)
    gt_cls = torch.stack(gt_cls)
    npc, npl, npr = torch.stack(npc).clone(), torch.stack(npl).clone(), torch.stack(npr).clone()
    pos = (gt_cls.sum(-1) > 0) & valid
    pred = torch.cat(out_offsets, dim=1)[pos]
    gto = torch.stack(gt_off)[pos]
    num_pos = int(pos.sum().item())
    loss_normalizer = 0.9 * loss_normalizer + 0.1 * max(num_pos, 1)
    logits = torch.cat(out_cls_logits, dim=1)
    cl = sigmoid_focal_loss(logits[valid], gt_cls[valid])
    npc = torch.where(pos, npc, torch.ones_like(npc))
    cls_loss = (cl.sum(-1) * npc[valid]).sum() / loss_normalizer
    # label-involved ("al") loss, meta_archs.py:1436-1446
    s = logits.masked_fill(~valid.unsqueeze(-1), -1e7).softmax(-1).max(dim=1)[0]
    inv = torch.zeros_like(s)
    for i, lab in enumerate(gt_labels):
        inv[i, lab] = 1
    al_loss = (-inv * s.log() - (1 - inv) * (1 - s).log()).sum() / loss_normalizer
    if num_pos == 0:
        reg_loss = 0 * pred.sum()
    else:
        rl = ctr_diou_loss_1d(pred, gto)
        rl = rl * (npl[pos] + npr[pos]) / 2.0 * npc[pos]
        reg_loss = rl.sum() / loss_normalizer
    final = cls_loss + reg_loss * cfg.loss_weight + al_loss * cfg.al_loss_weight
    return {"cls_loss": cls_loss, "reg_loss": reg_loss, "al_loss": al_loss, "final_loss": final}, loss_normalizer


# ----------------------------------------------------------------------------------------------------
# decode + NMS
# ----------------------------------------------------------------------------------------------------
def decode_single_video(cfg, pts_list, masks, logits, offsets):
    """inference_single_video — meta_archs.py:1594-1692 (cls_preds_per_vid is None).  Inputs per level:
    pts (T_l,4), mask (T_l,), logits (T_l,K), offsets (T_l,2)."""
    K = logits[0].shape[-1]
    segs_all, scores_all, cls_all = [], [], []
    for cls_i, off_i, pts_i, mask_i in zip(logits, offsets, pts_list, masks):
        prob = (cls_i.sigmoid() * mask_i.unsqueeze(-1)).flatten()
        keep = prob > cfg.pre_nms_thresh
        prob = prob[keep]
        topk_idxs = keep.nonzero(as_tuple=True)[0]
        num_topk = min(cfg.pre_nms_topk, topk_idxs.size(0))
        prob, idxs = prob.sort(descending=True)
        prob = prob[:num_topk].clone()
        topk_idxs = topk_idxs[idxs[:num_topk]].clone()
        pt_idxs = torch.div(topk_idxs, K, rounding_mode="floor")
        cls_idxs = torch.fmod(topk_idxs, K)
        offs = off_i[pt_idxs]
        pts = pts_i[pt_idxs]
        left = pts[:, 0] - offs[:, 0] * pts[:, 3]
        right = pts[:, 0] + offs[:, 1] * pts[:, 3]
        keep2 = (right - left) > cfg.duration_thresh
        segs_all.append(torch.stack((left, right), -1)[keep2])
        scores_all.append(prob[keep2])
        cls_all.append(cls_idxs[keep2])
    return torch.cat(segs_all), torch.cat(scores_all), torch.cat(cls_all)


def softnms_1d(segs, scores, iou_threshold, sigma, min_score, method):
    """softnms_1d_cpu — MQ/libs/utils/csrc/nms_cpu.cpp:67-160, float32 arithmetic step by step (numpy scalars).
    Returns (dets (n,3) float32 in pick order, inds (n,) int64)."""
    f = np.float32
    n = int(segs.shape[0])
    if n == 0:
        return np.zeros((0, 3), np.float32), np.zeros((0,), np.int64)
    x1 = np.ascontiguousarray(segs[:, 0], dtype=np.float32).copy()
    x2 = np.ascontiguousarray(segs[:, 1], dtype=np.float32).copy()
    sc = np.ascontiguousarray(scores, dtype=np.float32).copy()
    ar = (x2 - x1 + f(1e-6)).astype(np.float32)
    inds = np.arange(n, dtype=np.int64)
    dets = np.zeros((n, 3), np.float32)
    sig = f(sigma); thr = f(iou_threshold); ms = f(min_score)
    i = 0
    while i < n:
        mp = i + int(np.argmax(sc[i:n]))  # first maximum wins (strict '<' in the reference scan)
        for a in (x1, x2, sc, ar, inds):
            a[i], a[mp] = a[mp], a[i]
        ix1, ix2, iar = x1[i], x2[i], ar[i]
        dets[i] = (ix1, ix2, sc[i])
        pos = i + 1
        while pos < n:
            inter = max(f(0.0), f(min(ix2, x2[pos]) - max(ix1, x1[pos])))
            ovr = f(inter / f(f(iar + ar[pos]) - inter))
            w = f(1.0)
            if method == 0:
                if ovr >= thr:
                    w = f(0.0)
            elif method == 1:
                if ovr >= thr:
                    w = f(f(1.0) - ovr)
            else:
                w = f(math_expf(f(-(f(ovr * ovr)) / sig)))
            sc[pos] = f(sc[pos] * w)
            if sc[pos] < ms:
                n -= 1
                for a in (x1, x2, sc, ar, inds):
                    a[pos] = a[n]
                pos -= 1
            pos += 1
        i += 1
    return dets[:n].copy(), inds[:n].copy()


_libm = None


def math_expf(x):
    """float32 exp exactly as the C library computes it: `std::exp(float)` in nms_cpu.cpp:136 is glibc's expf
    (numpy's vectorised float32 exp differs from it in the last bit, so it cannot be used here)."""
    global _libm
    if _libm is None:
        import ctypes
        import ctypes.util
        _libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
        _libm.expf.restype = ctypes.c_float
        _libm.expf.argtypes = [ctypes.c_float]
    return np.float32(_libm.expf(float(x)))


def nms_1d(segs, scores, iou_threshold):
    """nms_1d_cpu — nms_cpu.cpp:19-57 (greedy hard NMS; returns kept original indices, score-descending)."""
    n = int(segs.shape[0])
    if n == 0:
        return np.zeros((0,), np.int64)
    x1 = np.asarray(segs[:, 0], dtype=np.float32)
    x2 = np.asarray(segs[:, 1], dtype=np.float32)
    ar = (x2 - x1 + np.float32(1e-6)).astype(np.float32)
    order = torch.sort(torch.as_tensor(np.asarray(scores, dtype=np.float32)), 0, descending=True)[1].numpy()
    sel = np.ones(n, bool)
    for _i in range(n):
        if not sel[_i]:
            continue
        i = order[_i]
        rest = order[_i + 1:]
        inter = np.maximum(np.float32(0), np.minimum(x2[i], x2[rest]) - np.maximum(x1[i], x1[rest])).astype(np.float32)
        ovr = inter / ((ar[i] + ar[rest]).astype(np.float32) - inter)
        sel[_i + 1:] &= ~(ovr >= np.float32(iou_threshold))
    return order[sel]


def batched_nms(segs, scores, cls_idxs, iou_threshold, min_score, max_seg_num, use_soft_nms=True, multiclass=True,
                sigma=0.5, voting_thresh=0.75, softnms_fn=None):
    """batched_nms — MQ/libs/utils/nms.py:103-190 (+ SoftNMSop :38-64, NMSop :8-35).  torch CPU tensors in/out.
    `softnms_fn(segs, scores, iou, sigma, min_score, method) -> (dets, inds)` defaults to the numpy restatement;
    tests also plug in oracle/softnms.c and the reference extension here."""
    softnms_fn = softnms_fn or softnms_1d
    if segs.shape[0] == 0:
        return torch.zeros([0, 2]), torch.zeros([0]), torch.zeros([0], dtype=cls_idxs.dtype)

    def one(s, sc, ci):
        if use_soft_nms:
            dets, inds = softnms_fn(s.numpy(), sc.numpy(), float(iou_threshold), float(sigma), float(min_score), 2)
            n = min(len(inds), max_seg_num) if max_seg_num > 0 else len(inds)
            dets = torch.from_numpy(np.asarray(dets))
            return dets[:n, :2].clone(), dets[:n, 2].clone(), ci[torch.from_numpy(np.asarray(inds))][:n].clone()
        if min_score > 0:
            vm = sc > min_score
            s, sc, ci = s[vm], sc[vm], ci[vm]
        inds = torch.from_numpy(nms_1d(s.numpy(), sc.numpy(), float(iou_threshold)))
        if max_seg_num > 0:
            inds = inds[:min(max_seg_num, len(inds))]
        return s[inds].clone(), sc[inds].clone(), ci[inds].clone()

    if multiclass:
        ns, nsc, nc = [], [], []
        for c in torch.unique(cls_idxs):
            idx = torch.where(cls_idxs == c)[0]
            a, b, d = one(segs[idx], scores[idx], cls_idxs[idx])
            ns.append(a); nsc.append(b); nc.append(d)
        ns, nsc, nc = torch.cat(ns), torch.cat(nsc), torch.cat(nc)
    else:
        ns, nsc, nc = one(segs, scores, cls_idxs)
        if voting_thresh > 0:
            ns = seg_voting(ns, segs, scores, voting_thresh)
    _, idxs = nsc.sort(descending=True)
    m = min(max_seg_num, ns.shape[0])
    return ns[idxs[:m]], nsc[idxs[:m]], nc[idxs[:m]]


def seg_voting(nms_segs, all_segs, all_scores, iou_threshold):
    """seg_voting — nms.py:67-101."""
    a = nms_segs[:, None].expand(nms_segs.shape[0], all_segs.shape[0], 2)
    b = all_segs[None, :].expand(nms_segs.shape[0], all_segs.shape[0], 2)
    inter = (torch.minimum(a[:, :, 1], b[:, :, 1]) - torch.maximum(a[:, :, 0], b[:, :, 0])).clamp(min=0)
    iou = inter / ((a[:, :, 1] - a[:, :, 0]) + (b[:, :, 1] - b[:, :, 0]) - inter)
    w = (iou >= iou_threshold).to(all_scores.dtype) * all_scores[None, :] * iou
    w = w / torch.sum(w, dim=1, keepdim=True)
    return w @ all_segs


def postprocess(cfg, segs, scores, labels, fps, duration, feat_stride, feat_num_frames, softnms_fn=None):
    """postprocessing — meta_archs.py:1695-1736."""
    segs, scores, labels = batched_nms(segs, scores, labels, cfg.iou_threshold, cfg.min_score, cfg.max_seg_num,
                                       use_soft_nms=True, multiclass=True, sigma=cfg.nms_sigma, softnms_fn=softnms_fn)
    if segs.shape[0] > 0:
        segs = (segs * feat_stride + 0.5 * feat_num_frames) / fps
        segs[segs <= 0.0] *= 0.0
        segs[segs >= duration] = segs[segs >= duration] * 0.0 + duration
    return segs, scores, labels


# ----------------------------------------------------------------------------------------------------
# whole-model entry points on reference-style video_list dicts
# ----------------------------------------------------------------------------------------------------
def preprocess(cfg, video_list, training):
    """preprocessing + query_preprocessing — meta_archs.py:1134-1221 (batch > 1 allowed in eval here)."""
    vl = [v for v in video_list if len(v["labels"]) > 0] if training else video_list
    lens = [v["feats"].shape[-1] for v in vl]
    max_len = max(lens)
    if training or max_len <= cfg.max_seq_len:
        max_len = cfg.max_seq_len
    else:
        stride = cfg.strides[-1]
        max_len = (max_len + stride - 1) // stride * stride
    x = torch.zeros(len(vl), vl[0]["feats"].shape[0], max_len)
    for i, v in enumerate(vl):
        x[i, :, :lens[i]] = v["feats"]
    mask = (torch.arange(max_len)[None, :] < torch.tensor(lens)[:, None]).unsqueeze(1)
    text, tmask = None, None
    if cfg.use_cross_modal:
        tl = [v["prompt_feature"].shape[-1] for v in video_list]
        text = torch.zeros(len(video_list), video_list[0]["prompt_feature"].shape[0], max(tl))
        for i, v in enumerate(video_list):
            text[i, :, :tl[i]] = v["prompt_feature"]
        tmask = (torch.arange(max(tl))[None, :] < torch.tensor(tl)[:, None]).unsqueeze(1)
        if cfg.prompt_pool is not None and not training:  # meta_archs.py:759-780 (mask from the PRE-prompt lengths)
            text = prompt_prepend(P_for_prompt[0], cfg, text)
            tmask = (torch.arange(text.shape[-1])[None, :] < torch.tensor(tl)[:, None]).unsqueeze(1)
    return x, mask, text, tmask


P_for_prompt = [None]  # set by model_infer (keeps preprocess' signature)


def model_train_losses(P, cfg, video_list, loss_normalizer=None):
    """PtTransformer.forward(is_training=True) — meta_archs.py:753-945 for cl_cfg.name None (no prompts/SSL)."""
    x, mask, text, tmask = preprocess(cfg, video_list, True)
    logits, offs, masks, _ = forward_heads(P, cfg, x, mask, text, tmask, training=True)
    segs = [v["segments"] for v in video_list if len(v["labels"]) > 0]
    labs = [v["labels"] for v in video_list if len(v["labels"]) > 0]
    ln = cfg.init_loss_norm if loss_normalizer is None else loss_normalizer
    return losses(P, cfg, masks, logits, offs, segs, labs, ln)


def model_infer(P, cfg, video_list, softnms_fn=None, return_raw=False):
    """PtTransformer.forward(is_training=False) — meta_archs.py:753-969, 1527-1736, one result dict per video."""
    results = []
    raw = []
    P_for_prompt[0] = P
    for v in video_list:
        x, mask, text, tmask = preprocess(cfg, [v], False)
        logits, offs, masks, _ = forward_heads(P, cfg, x, mask, text, tmask, training=False)
        pts = points(cfg, [m.shape[1] for m in masks])
        segs, scores, labels = decode_single_video(cfg, pts, [m[0] for m in masks], [l[0] for l in logits],
                                                   [o[0] for o in offs])
        raw.append((logits, offs, masks, segs, scores, labels))
        s, sc, lb = postprocess(cfg, segs, scores, labels, v["fps"], v["duration"], v["feat_stride"],
                                v["feat_num_frames"], softnms_fn)
        results.append({"video_id": v["video_id"], "segments": s, "scores": sc, "labels": lb})
    return (results, raw) if return_raw else results
