"""Parameter tree of the convTransformer backbone (reference: MQ/libs/modeling/backbones.py:12-289); the forward
runs in vilco_b200.engine.backbone_fwd."""
import torch
from torch import nn

from .blocks import LayerNorm, MaskedConv1D, TransformerBlock
from .modeling_xlnet_x import XLNetConfig, XLNetModel
from .models import register_backbone


@register_backbone("convTransformer")
class ConvTransformerBackbone(nn.Module):
    def __init__(self, n_in, n_embd, n_head, n_embd_ks, max_len, use_xl, arch=(2, 2, 5), t_c_alpha=0.8, scale_factor=2,
                 with_ln=False, attn_pdrop=0.0, proj_pdrop=0.0, path_pdrop=0.0, use_abs_pe=False, use_rel_pe=False,
                 use_dcn=False, dcn_start_layer=0, use_cross_modal=False, n_txt_in=768, mha_win_size=-1):
        super().__init__()
        assert len(arch) == 3
        assert not use_dcn, "deformable convs are not on the MQ path (use_dcn is False in every MQ config)"
        self.t_c_alpha, self.arch, self.max_len, self.scale_factor = t_c_alpha, arch, max_len, scale_factor
        self.use_abs_pe, self.use_rel_pe, self.use_xl, self.use_cross_modal = use_abs_pe, use_rel_pe, use_xl, use_cross_modal
        self.n_in = n_in
        assert isinstance(n_in, (list, tuple)) and len(n_in) == 1, "MQ configs use input_dim: [4096] (one projection)"
        assert isinstance(n_embd, (list, tuple)) and len(n_in) == len(n_embd)
        self.proj = nn.ModuleList([MaskedConv1D(c0, c1, 1) for c0, c1 in zip(n_in, n_embd)])
        n_in = n_embd = sum(n_embd)
        self.n_embd, self.n_head = n_embd, n_head

        def blocks(n, strides, cross):
            return nn.ModuleList([TransformerBlock(n_embd, n_head, n_ds_strides=strides, attn_pdrop=attn_pdrop,
                                                   proj_pdrop=proj_pdrop, path_pdrop=path_pdrop, t_c_alpha=t_c_alpha,
                                                   use_rel_pe=use_rel_pe, use_cross_modal=cross,
                                                   mha_win_size=mha_win_size) for _ in range(n)])

        def embd(n, first_in, k):
            convs, norms = nn.ModuleList(), nn.ModuleList()
            for idx in range(n):
                convs.append(MaskedConv1D(first_in if idx == 0 else n_embd, n_embd, k, stride=1, padding=k // 2,
                                          bias=(not with_ln)))
                norms.append(LayerNorm(n_embd) if with_ln else nn.Identity())
            return convs, norms

        assert with_ln, "MQ configs use embd_with_ln: True"
        self.embd, self.embd_norm = embd(arch[0], n_in, n_embd_ks)
        self.stem = blocks(arch[1], (1, 1), use_cross_modal)
        self.branch = blocks(arch[2], (scale_factor, scale_factor), use_cross_modal)
        if use_xl:
            self.xlnet = XLNetModel(XLNetConfig.for_width(n_embd))
        if use_cross_modal:
            self.txt_embd, self.txt_embd_norm = embd(arch[0], n_txt_in, 1)
            self.txt_stem = blocks(arch[1], (1, 1), False)
        self.apply(self.__init_weights__)

    def __init_weights__(self, module):
        if isinstance(module, (nn.Linear, nn.Conv1d)) and module.bias is not None:
            torch.nn.init.constant_(module.bias, 0.0)
