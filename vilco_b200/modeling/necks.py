"""FPNIdentity / FPN1D parameter trees (reference: MQ/libs/modeling/necks.py:118-198 and :13-106); forwards in
engine.neck_heads_fwd / engine.fpn1d_fwd."""
import torch
from torch import nn

from .blocks import LayerNorm
from .models import register_neck


@register_neck("identity")
class FPNIdentity(nn.Module):
    def __init__(self, in_channels, out_channel, scale_factor=2.0, start_level=0, end_level=-1, with_ln=True,
                 use_us_fpn=False):
        super().__init__()
        assert not use_us_fpn and start_level == 0 and with_ln
        self.in_channels, self.out_channel = in_channels, out_channel
        self.end_level = len(in_channels) if end_level == -1 else end_level
        self.fpn_norms = nn.ModuleList()
        for i in range(start_level, self.end_level):
            assert in_channels[i] == out_channel
            self.fpn_norms.append(LayerNorm(out_channel))


class _DenseBlock(nn.Module):           # MQ/libs/modeling/utils.py:671-689
    def __init__(self, input_num, num1, num2, rate):
        super().__init__()
        self.conv1x1 = nn.Conv1d(input_num, num1, 1)
        self.ConvGN = nn.GroupNorm(32, num1)
        self.dilaconv = nn.Conv1d(num1, num2, 3, padding=rate, dilation=rate)


class _DenseAPP(nn.Module):             # utils.py:692-714
    def __init__(self, num_channels):
        super().__init__()
        for k, rate in enumerate((3, 6, 12, 18, 24)):
            setattr(self, f"aspp{rate}", _DenseBlock(num_channels + 256 * k, 512, 256, rate))
        self.conv1x1 = nn.Conv1d(5 * 256, num_channels, 1)
        self.ConvGN = nn.GroupNorm(32, num_channels)


class _Attn(nn.Module):                 # CxAM / CnAM (utils.py:619-668): registered by the reference, never called (necks.py / utils.py:741-745)
    def __init__(self, c, order):
        super().__init__()
        for n in order:
            setattr(self, n, nn.Conv1d(c, c if n == "value_conv" else c // 8, 1))


class _ACConv(nn.Module):               # utils.py:732-751
    def __init__(self, d):
        super().__init__()
        self.denseapp = _DenseAPP(d)
        self.CxAM = _Attn(d, ("key_conv", "query_conv", "value_conv"))
        self.CnAM = _Attn(d, ("query_conv", "key_conv", "value_conv"))


class _Conv(nn.Module):                 # MaskedConv1D parameter holder
    def __init__(self, cin, cout, k, groups=1):
        super().__init__()
        self.conv = nn.Conv1d(cin, cout, k, padding=k // 2, groups=groups, bias=False)


@register_neck("fpn")
class FPN1D(nn.Module):
    """Parameter tree of the reference's FPN1D (same names / shapes / order); evaluation forward = engine.fpn1d_fwd."""

    def __init__(self, in_channels, out_channel, scale_factor=2.0, start_level=0, end_level=-1, with_ln=True, use_us_fpn=False):
        super().__init__()
        assert start_level == 0 and with_ln and scale_factor == 2.0 and end_level in (-1, len(in_channels))
        assert all(c == out_channel for c in in_channels) and out_channel % 32 == 0, "FPN1D assumes equal widths (necks.py:42)"
        self.in_channels, self.out_channel = in_channels, out_channel
        n = len(in_channels)
        self.lateral_convs = nn.ModuleList(_Conv(out_channel, out_channel, 1) for _ in range(n))
        self.ac_conv = _ACConv(out_channel)
        self.fpn_convs = nn.ModuleList(_Conv(out_channel, out_channel, 3, groups=out_channel) for _ in range(n))
        self.fpn_norms = nn.ModuleList(LayerNorm(out_channel) for _ in range(n))
