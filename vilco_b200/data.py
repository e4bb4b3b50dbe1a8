"""Data path in front of the model: the per-clip work `Ego4dCLDataset.__getitem__` does on the CPU before
the hot loop sees a clip (MQ/libs/datasets/ego4d.py:610-651) —

    feats = torch.load(filename)                       # (T_in, 4096) fp32, token-major on disk
    feats = feats.permute(1, 0)                        # (4096, T_in)
    feats = F.interpolate(feats[None], size=max_seq_len, mode='linear', align_corners=False)[0]    # force_upsampling

— as one sm_100a kernel (`vilco_resize_feats`) on the features *as stored*: the upload is the raw `(T_in, C)` block
(not the resized 16 MiB per clip), the resize reads it token-major and writes token-major, and the result can be taken

* in the reference layout `(C, max_seq_len)` on the device (`resize_feats`; a drop-in for the three lines above — the
  model and `EvalGraph.load_inputs` accept device tensors as `video['feats']`), or
* directly as the bf16 operand planes the backbone's first GEMM reads (`resize_pack`), skipping the transpose of
  `PtTransformer.preprocessing` + `vilco_pack_feats` altogether.

No CPU fallback: without the CUDA library these functions raise.
"""
import torch

from . import lib as L
from . import ops


def _concat_clips(clips):
    """list of (T_i, C) fp32 tensors (host — pinned or not — or device) -> (sum T_i, C) device tensor + (B+1,) row offsets.
    Host clips are uploaded as they are stored; nothing is resized or transposed on the CPU."""
    if isinstance(clips, torch.Tensor):
        clips = [clips]
    if not clips:
        raise ValueError("resize_feats: no clips")
    C = clips[0].shape[1]
    starts = [0]
    for f in clips:
        if f.dim() != 2 or f.shape[1] != C or f.shape[0] < 1:
            raise ValueError(f"resize_feats: every clip must be (T >= 1, {C}), got {tuple(f.shape)}")
        if f.dtype != torch.float32:
            raise TypeError(f"resize_feats: fp32 features expected, got {f.dtype}")
        starts.append(starts[-1] + f.shape[0])
    if C % 4:
        raise ValueError(f"resize_feats: feature dim {C} must be a multiple of 4")
    dev = torch.device("cuda", torch.cuda.current_device())
    if len(clips) == 1 and clips[0].is_cuda and clips[0].is_contiguous():
        x = clips[0]
    else:
        x = torch.empty(starts[-1], C, device=dev, dtype=torch.float32)
        for f, r0, r1 in zip(clips, starts[:-1], starts[1:]):
            x[r0:r1].copy_(f, non_blocking=True)
    row_start = torch.tensor(starts, dtype=torch.int64).to(dev, non_blocking=True)
    return x, row_start, len(clips), C


def _launch(clips, max_seq_len, want32, want16):
    x, row_start, B, C = _concat_clips(clips)
    o32 = torch.empty(B, max_seq_len, C, device=x.device, dtype=torch.float32) if want32 else None
    o16 = ops.empty16(B, max_seq_len, C, device=x.device, planes=ops.PLANES_HI) if want16 else None   # feeds the input projection
    L.check(L.lib().vilco_resize_feats(ops._p(x), ops._p(row_start), B, C, int(max_seq_len), ops._p(o32), ops._p(o16),
                                       ops._i64(ops.lo(o16) if want16 else 0), L.stream_ptr()), "vilco_resize_feats")
    return o32, o16


def resize_feats(clips, max_seq_len):
    """(T_in, C) clip (or a list of them) -> (C, max_seq_len) device tensor(s) in the reference layout, equal to the
    `force_upsampling` branch of ego4d.py:644-651 (fp32; |Δ| ≤ 1 ulp of the two products, tests/test_gpu_data.py)."""
    single = isinstance(clips, torch.Tensor)
    o32, _ = _launch(clips, max_seq_len, True, False)
    out = ops.unpack(o32)                               # (B, T, C) -> (B, C, T)
    return out[0] if single else list(out.unbind(0))


def resize_feats_into(clips, out):
    """list of B raw (T_in, C) clips -> `out` (B, C, max_seq_len), a preallocated device buffer in the reference layout (the
    static input buffer of an `EvalGraph`): upload of the raw blocks, one resize launch, one transpose launch."""
    o32, _ = _launch(clips, out.shape[2], True, False)
    ops.unpack(o32, out=out)
    return out


def resize_pack(clips, max_seq_len):
    """list of (T_in, C) clips -> operand planes (NP, B, max_seq_len, C): what `ops.pack_feats` returns for the resized,
    transposed batch, without either intermediate."""
    _, o16 = _launch(clips, max_seq_len, False, True)
    return o16


def feat_stride_after_resize(duration, fps, max_seq_len):
    """`feat_stride` / `num_frames` the dataset reports for a force-upsampled fixed-length clip (ego4d.py:631-640)."""
    stride = duration * fps / max_seq_len
    return stride, stride


def truncate_feats(data_dict, max_seq_len, trunc_thresh, crop_ratio=None, max_num_trials=200, has_action=True, no_trunc=False):
    """Training-time random crop of a clip — same contract, same consumption of Python's `random` stream and same result as
    `truncate_feats` of MQ/libs/datasets/data_utils.py:24-112 (so a seeded run draws the identical windows), with two
    differences in how it is done: `feats` may live on the device (the crop is a device slice, e.g. of `resize_feats`
    output), and only the entries that change are copied — the reference deep-copies the whole dict, 16 MiB of features
    included, before it knows the window.

    Window search: up to `max_num_trials` uniform windows of the target length; accepted when (has_action) at least one
    segment keeps >= trunc_thresh of its length inside, or (no_trunc) additionally no segment is cut; the last trial is
    used if none qualifies.  Segments are clipped to the window, filtered by the threshold and shifted to its origin."""
    import random
    feat_len = data_dict['feats'].shape[1]
    if feat_len <= max_seq_len:
        if crop_ratio is None:
            return data_dict
        max_seq_len = random.randint(max(round(crop_ratio[0] * feat_len), 1), min(round(crop_ratio[1] * feat_len), feat_len))
        if feat_len == max_seq_len:
            return data_dict
    segs = data_dict['segments']
    length = torch.abs(segs[:, 1] - segs[:, 0])
    for _ in range(max_num_trials):
        st = random.randint(0, feat_len - max_seq_len)
        ed = st + max_seq_len
        left = segs[:, 0].clamp(min=float(st))
        right = segs[:, 1].clamp(max=float(ed))
        ratio = (right - left).clamp(min=0) / length
        keep = ratio >= trunc_thresh
        if no_trunc:
            if bool(keep.any()) and not bool(((ratio > 0.0) & (ratio < 1.0)).any()):
                break
        elif not has_action or bool(keep.any()):
            break
    out = dict(data_dict)                       # shallow: untouched entries (ids, text features, ...) are shared
    out['feats'] = data_dict['feats'][:, st:ed].clone()
    out['segmentation_labels'] = data_dict['segmentation_labels'][st:ed, :].clone()
    out['segments'] = torch.stack((left[keep], right[keep]), dim=1) - st
    out['labels'] = data_dict['labels'][keep].clone()
    return out
