"""FPNIdentity parameter tree (reference: MQ/libs/modeling/necks.py:118-198); forward in engine.neck_heads_fwd."""
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
