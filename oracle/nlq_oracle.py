"""TEST INFRASTRUCTURE ONLY — CPU restatement (torch fp32) of the NLQ model's forward pass, SURVEY.md §8f-1, composed from the
operator restatements of oracle/mq_oracle.py (the NLQ blocks are the MQ blocks without the channel mix and with the windowed
attention live).  Pinned against the reference's own NLQ model: tests/golden/nlq_small.npz (oracle/gen_golden_nlq.py),
tests/test_nlq_pin.py.  The product does not import this.

Follows NLQ/libs/modeling/backbones.py:546-615 (ConvTransformerBackbone.forward), blocks.py:840-874 (TransformerBlock.forward),
necks.py (FPNIdentity), meta_archs.py:182-338 (heads), :614-748 (PtTransformer.forward, evaluation branch).
"""
import torch
import torch.nn.functional as F

from . import mq_oracle as O


class NlqCfg:
    """ego4d_nlq_v2_egovlp_1e-4.yaml"""

    def __init__(self, **kw):
        self.input_vid_dim, self.input_txt_dim, self.embd_dim, self.n_head = 256, 512, 384, 4
        self.max_seq_len, self.arch, self.window, self.scale_factor, self.num_classes = 2560, (2, 4, 4, 0, 6), 9, 2, 1
        for k, v in kw.items():
            assert hasattr(self, k), k
            setattr(self, k, v)


def backbone(P, cfg, vid, vid_mask, txt, txt_mask):
    """vid (B, Cv, T), vid_mask (B, 1, T) bool, txt (B, Ct, L), txt_mask (B, 1, L) bool -> per-level features and masks."""
    pre = "backbone."
    x, m = vid, vid_mask
    for i in range(cfg.arch[0]):                                                    # backbones.py:552-554
        x, m = O.masked_conv1d(x, m, P[pre + f"vid_embd.{i}.conv.weight"], None)
        x = F.relu(O.channel_layernorm(x, P[pre + f"vid_embd_norm.{i}.weight"], P[pre + f"vid_embd_norm.{i}.bias"]))
    T = x.shape[-1]
    pe = O.sinusoid_pe(cfg.max_seq_len, cfg.embd_dim) / (cfg.embd_dim ** 0.5)       # :446-448, evaluation branch :564-572
    if T >= cfg.max_seq_len:
        pe = F.interpolate(pe, T, mode="linear", align_corners=False)
    x = x + pe[:, :, :T] * m.to(x.dtype)
    q, qm = txt, txt_mask
    for i in range(cfg.arch[0]):                                                    # :577-579 (k = 1 convolutions)
        q, qm = O.masked_conv1d(q, qm, P[pre + f"txt_embd.{i}.conv.weight"], None)
        q = F.relu(O.channel_layernorm(q, P[pre + f"txt_embd_norm.{i}.weight"], P[pre + f"txt_embd_norm.{i}.bias"]))
    for i in range(cfg.arch[1]):                                                    # :585-586 global self-attention on the text
        q, qm = O.transformer_block(P, pre + f"txt_stem.{i}.", q, qm, cfg.n_head, 1, channel_mix=False)
    qml = qm.squeeze(1).long()
    for i in range(cfg.arch[2]):                                                    # :589-590 window attention + cross attention
        x, m = O.transformer_block(P, pre + f"vid_stem.{i}.", x, m, cfg.n_head, 1, q, qml, window=cfg.window,
                                   channel_mix=False)
    feats, masks = [x], [m]
    for i in range(cfg.arch[3] + cfg.arch[4]):                                      # :601-604; the first arch[3] blocks cross-attend
        cy, cm = (q, qml) if i < cfg.arch[3] else (None, None)
        x, m = O.transformer_block(P, pre + f"branch.{i}.", x, m, cfg.n_head, cfg.scale_factor, cy, cm, window=cfg.window,
                                   channel_mix=False)
        feats.append(x)
        masks.append(m)
    return feats, masks


def forward_heads(P, cfg, vid, vid_mask, txt, txt_mask):
    """-> (list of (B, T_l, K) logits, list of (B, T_l, 2) offsets, list of (B, T_l) masks), the `get_emb=True` return of
    PtTransformer.forward (meta_archs.py:744-745)."""
    feats, masks = backbone(P, cfg, vid, vid_mask, txt, txt_mask)
    fpn, masks = O.fpn_identity(P, feats, masks)
    logits = [t.permute(0, 2, 1) for t in O.cls_head(P, fpn, masks)]
    offsets = [t.permute(0, 2, 1) for t in O.reg_head(P, fpn, masks)]
    return logits, offsets, [m.squeeze(1) for m in masks]


def preprocess_eval(cfg, clip):
    """PtTransformer.preprocessing / query_preprocessing for one evaluation clip (meta_archs.py:918-957): pad to max_seq_len."""
    f = clip["feats"]
    T = cfg.max_seq_len
    assert f.shape[-1] <= T
    vid = F.pad(f, [0, T - f.shape[-1]]).unsqueeze(0)
    vmask = (torch.arange(T)[None, :] < f.shape[-1]).unsqueeze(1)
    txt = clip["query_feats"].unsqueeze(0)
    tmask = torch.ones(1, 1, txt.shape[-1], dtype=torch.bool)
    return vid, vmask, txt, tmask


class NlqTestCfg:
    """test_cfg of ego4d_nlq_v2_egovlp_1e-4.yaml in the field names oracle/mq_oracle.py's decode / NMS restatements read."""

    def __init__(self, cfg):
        self.pre_nms_thresh, self.pre_nms_topk, self.duration_thresh = 0.001, 2000, 0.001
        self.iou_threshold, self.min_score, self.max_seg_num, self.nms_sigma = 0.1, 0.001, 5, 0.75
        self.max_seq_len = cfg.max_seq_len
        n_levels = 1 + cfg.arch[3] + cfg.arch[4]
        self.strides = [cfg.scale_factor ** l for l in range(n_levels)]
        self.regression_range = [[0, 4], [2, 8], [4, 16], [8, 32], [16, 64], [32, 128], [64, 10000]]


def infer(P, cfg, clip, softnms_fn=None):
    """PtTransformer.forward(is_training=False) for one clip -> (segments (n, 2) seconds, scores, labels): the decode and
    soft-NMS restatements of the MQ oracle apply unchanged (NLQ/libs/modeling/meta_archs.py inference / postprocessing)."""
    tc = NlqTestCfg(cfg)
    logits, offsets, masks = forward_heads(P, cfg, *preprocess_eval(cfg, clip))
    lens = [t.shape[1] for t in logits]
    segs, scores, labels = O.decode_single_video(tc, O.points(tc, lens), [m[0] for m in masks], [t[0] for t in logits],
                                                 [t[0] for t in offsets])
    return O.postprocess(tc, segs, scores, labels, clip["fps"], clip["duration"], clip["feat_stride"],
                         clip["feat_num_frames"], softnms_fn=softnms_fn)


def label_points(points, gt_segments, gt_onehot, radius=1.5):
    """Classification / regression targets of one clip, the classic ActionFormer assignment the NLQ model keeps
    (NLQ/libs/modeling/meta_archs.py:980-1072, `center_sample: radius`): a point is positive for a moment when it lies within
    `radius` strides of the moment's centre (clipped to the moment) and the farther boundary falls into the level's regression
    range; among several candidates the shortest moment wins.  points (P, 4) = [t, lo, hi, stride]; returns (P, K), (P, 2)."""
    n_pts, n_gt = points.shape[0], gt_segments.shape[0]
    if n_gt == 0:
        return gt_segments.new_zeros((n_pts, gt_onehot.shape[1])), gt_segments.new_zeros((n_pts, 2))
    t, lo, hi, stride = (points[:, j, None] for j in range(4))
    s0, s1 = gt_segments[None, :, 0], gt_segments[None, :, 1]
    reg = torch.stack((t - s0, s1 - t), dim=-1)                                     # (P, N, 2)
    centre = 0.5 * (s0 + s1)
    inside = torch.minimum(t - torch.maximum(centre - stride * radius, s0), torch.minimum(centre + stride * radius, s1) - t) > 0
    far = reg.max(-1)[0]
    in_range = (far >= lo) & (far <= hi)
    length = (s1 - s0).repeat(n_pts, 1).masked_fill(~(inside & in_range), float("inf"))
    shortest, idx = length.min(dim=1)
    pick = ((length <= shortest[:, None] + 1e-3) & (length < float("inf"))).to(reg.dtype)
    cls_t = (pick @ gt_onehot.to(reg.dtype)).clamp(0.0, 1.0)
    reg_t = reg[torch.arange(n_pts), idx] / stride
    return cls_t, reg_t


def train_losses(P, cfg, clips, loss_normalizer=200.0, momentum=0.9, label_smoothing=0.1, loss_weight=1.0):
    """PtTransformer.forward in training mode with every dropout / drop-path probability at 0 (meta_archs.py:746-776) and
    `losses` (:1094-1155): focal loss over valid points (label smoothing s: y(1-s) + s/(K+1)), DIoU on positives, both divided
    by the EMA of the number of positives.  Returns the dict of the reference and the updated normaliser."""
    T = cfg.max_seq_len
    B = len(clips)
    vid = torch.zeros(B, cfg.input_vid_dim, T)
    L = max(c["query_feats"].shape[-1] for c in clips)
    txt = torch.zeros(B, cfg.input_txt_dim, L)
    for i, c in enumerate(clips):
        vid[i, :, :c["feats"].shape[-1]] = c["feats"]
        txt[i, :, :c["query_feats"].shape[-1]] = c["query_feats"]
    vmask = (torch.arange(T)[None, :] < torch.tensor([c["feats"].shape[-1] for c in clips])[:, None]).unsqueeze(1)
    tmask = (torch.arange(L)[None, :] < torch.tensor([c["query_feats"].shape[-1] for c in clips])[:, None]).unsqueeze(1)
    logits, offsets, masks = forward_heads(P, cfg, vid, vmask, txt, tmask)
    tc = NlqTestCfg(cfg)
    pts = torch.cat(O.points(tc, [t.shape[1] for t in logits]), dim=0)
    targets = [label_points(pts, c["segments"], c["one_hot_labels"]) for c in clips]
    gt_cls, gt_off = torch.stack([t[0] for t in targets]), torch.stack([t[1] for t in targets])
    valid = torch.cat(masks, dim=1)
    pos = (gt_cls.sum(-1) > 0) & valid
    num_pos = int(pos.sum())
    norm = momentum * loss_normalizer + (1 - momentum) * max(num_pos, 1)
    K = gt_cls.shape[-1]
    target = gt_cls[valid] * (1 - label_smoothing) + label_smoothing / (K + 1)
    cls_loss = O.sigmoid_focal_loss(torch.cat(logits, dim=1)[valid], target).sum() / norm
    pred = torch.cat(offsets, dim=1)[pos]
    reg_loss = O.ctr_diou_loss_1d(pred, gt_off[pos]).sum() / norm if num_pos else 0 * pred.sum()
    return {"cls_loss": cls_loss, "reg_loss": reg_loss, "final_loss": cls_loss + reg_loss * loss_weight}, norm
