"""Parameter container for the single XLNet layer the MQ backbone runs between stem and branch
(reference: MQ/libs/modeling/modeling_xlnet_x.py — XLNetRelativeAttention :210, XLNetFeedForward :470, XLNetLayer :493,
XLNetModel :938).  Names / shapes follow the reference state_dict, including the parameters that
never receive a gradient (word_embedding, mask_emb, r_s_bias, seg_embed).  The math runs in vilco_b200.engine.xlnet_layer_fwd.
"""
import json
import os

import torch
from torch import nn


class XLNetConfig:
    def __init__(self, **kw):
        self.vocab_size = 32000
        self.d_model, self.n_head, self.d_head, self.d_inner, self.n_layer = 1024, 16, 64, 2048, 1
        self.layer_norm_eps, self.dropout, self.initializer_range = 1e-12, 0.1, 0.02
        self.ff_activation, self.attn_type, self.bi_data, self.clamp_len = "gelu", "bi", False, -1
        for k, v in kw.items():
            setattr(self, k, v)

    @classmethod
    def from_dict(cls, d):
        return cls(**d)

    @classmethod
    def for_width(cls, n_embd, search_dirs=("configs",)):
        """Reads configs/xlnet_config_<n_embd>.json like the reference (cwd-relative, backbones.py:130-134); falls back
        to the shipped values of that file when it is not present."""
        for d in search_dirs:
            p = os.path.join(d, f"xlnet_config_{n_embd}.json")
            if os.path.exists(p):
                with open(p) as f:
                    return cls.from_dict(json.load(f))
        inner = {256: 1024, 512: 1024, 1024: 2048, 1536: 3072}.get(n_embd, 2 * n_embd)
        assert n_embd % 64 == 0
        return cls(d_model=n_embd, n_head=n_embd // 64, d_head=64, d_inner=inner)


class XLNetRelativeAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        shape = (config.d_model, config.n_head, config.d_head)
        for n in "qkvor":
            setattr(self, n, nn.Parameter(torch.empty(shape).normal_(0.0, config.initializer_range)))
        for n in ("r_r_bias", "r_s_bias", "r_w_bias"):
            setattr(self, n, nn.Parameter(torch.empty(config.n_head, config.d_head).normal_(0.0, config.initializer_range)))
        self.seg_embed = nn.Parameter(torch.empty(2, config.n_head, config.d_head).normal_(0.0, config.initializer_range))
        self.layer_norm = nn.LayerNorm(config.d_model, eps=config.layer_norm_eps)


class XLNetFeedForward(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.layer_norm = nn.LayerNorm(config.d_model, eps=config.layer_norm_eps)
        self.layer_1 = nn.Linear(config.d_model, config.d_inner)
        self.layer_2 = nn.Linear(config.d_inner, config.d_model)
        for lin in (self.layer_1, self.layer_2):
            lin.weight.data.normal_(0.0, config.initializer_range)
            lin.bias.data.zero_()


class XLNetLayer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.rel_attn = XLNetRelativeAttention(config)
        self.ff = XLNetFeedForward(config)


class XLNetModel(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.word_embedding = nn.Embedding(config.vocab_size, config.d_model)
        self.word_embedding.weight.data.normal_(0.0, config.initializer_range)
        self.mask_emb = nn.Parameter(torch.empty(1, 1, config.d_model).normal_(0.0, config.initializer_range))
        self.layer = nn.ModuleList([XLNetLayer(config) for _ in range(config.n_layer)])


class XLNetLMHeadModel(XLNetModel):  # imported by the reference's train_utils.py:20 for isinstance checks only
    pass
