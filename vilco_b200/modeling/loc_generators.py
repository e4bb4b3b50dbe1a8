"""Point generator (reference: MQ/libs/modeling/loc_generators.py:28-92): per level rows [t, reg_lo, reg_hi, stride]."""
import torch
from torch import nn

from .models import register_generator


class BufferList(nn.Module):
    def __init__(self, buffers):
        super().__init__()
        for i, buffer in enumerate(buffers):
            self.register_buffer(str(i), buffer, persistent=False)

    def __len__(self):
        return len(self._buffers)

    def __iter__(self):
        return iter(self._buffers.values())


@register_generator("point")
class PointGenerator(nn.Module):
    def __init__(self, max_seq_len, fpn_strides, regression_range, use_offset=False, use_us_fpn=False):
        super().__init__()
        assert len(regression_range) == len(fpn_strides) and not use_us_fpn
        self.max_seq_len, self.fpn_levels = max_seq_len, len(fpn_strides)
        self.fpn_strides, self.regression_range, self.use_offset = fpn_strides, regression_range, use_offset
        pts = []
        for l, stride in enumerate(fpn_strides):
            reg_range = torch.as_tensor(regression_range[l], dtype=torch.float)
            points = torch.arange(0, max_seq_len, stride)[:, None]
            if use_offset:
                points = points + 0.5 * stride
            n = points.shape[0]
            pts.append(torch.cat((points, reg_range[None].repeat(n, 1),
                                  torch.as_tensor(stride, dtype=torch.float)[None].repeat(n, 1)), dim=1))
        self.buffer_points = BufferList(pts)

    def forward(self, feat_lens):
        out = []
        for n, buf in zip(feat_lens, self.buffer_points):
            assert n <= buf.shape[0], "Reached max buffer length for point generator"
            out.append(buf[:n, :])
        return out
