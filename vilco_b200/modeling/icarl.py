"""iCaRL nearest-mean-of-exemplars re-scoring — the glue of `PtTransformer.classify` (MQ/libs/modeling/meta_archs.py:1061-1131)
and of the `cls_preds_per_vid` branch of `inference_single_video` (:1625-1643).  SURVEY.md §8a row 14 keeps this glue in
Python ("semantics fixed"): the functions below are device-agnostic torch expressions on the pyramid features / head outputs
that the CUDA path produces; the network passes themselves (backbone, neck, heads, NMS) are the kernels.

Reference behaviour kept as it is, including its oddities:
* a level's feature map is normalised by the Frobenius norm of the WHOLE (1, C, T_l) tensor, padding included (:1080, :1120);
* the distance table (T_l, n_classes) is indexed with the flattened (T_l, num_classes) logit index, so the exemplar memory
  must hold exactly `num_classes` classes, in class order;
* candidates are the entries whose distance is below the table's mean; the following "top-k by distance" step sorts the
  UNFILTERED table and applies those indices to the FILTERED arrays unless any of them is too large (:1636-1643).
"""
import torch

FPN_LEVELS = 10     # hard-coded in the reference (meta_archs.py:1065): classify only works for the 10-level architecture


def normalize_level(feat):
    """feat / ||feat||_F  (the `feat / feat.norm()` of :1080 and :1120)."""
    return feat / feat.norm()


def exemplar_mean(level_feats):
    """normalised mean of the normalised exemplar features of one class at one level (:1085-1089).
    level_feats: list of (1, C, T_l) tensors (one per exemplar clip, already normalised) -> (C, T_l)."""
    mu = torch.stack(level_feats, dim=0).mean(0).squeeze()
    return mu / mu.norm()


def nme_dists(level_feat, class_means):
    """squared distance of every time step's normalised feature column to every class mean (:1098-1127).
    level_feat (1, C, T_l) (un-normalised FPN output), class_means: list over classes of (C, T_l) -> (1, T_l, n_classes)."""
    f = normalize_level(level_feat)[0]                                  # (C, T_l)
    cols = [(f - mu).pow(2).sum(0) for mu in class_means]               # one class at a time: no (C, T_l, n_classes) temporary
    return torch.stack(cols, dim=1).unsqueeze(0)


def select_candidates(cls_i, offsets_i, pts_i, mask_i, dists_i, num_classes, pre_nms_topk, duration_thresh):
    """One pyramid level of `inference_single_video` when `cls_preds_per_vid` is given (:1625-1643, 1660-1682).
    cls_i (T_l, K) logits, offsets_i (T_l, 2), pts_i (T_l, 4), mask_i (T_l,) bool/float, dists_i (1, T_l, K).
    Returns (segments (n, 2), scores (n,), labels (n,) int64) in the reference's order."""
    pred_prob = (cls_i.sigmoid() * mask_i.unsqueeze(-1)).flatten()
    d = dists_i.flatten()
    keep1 = d < d.mean()
    pred_prob = pred_prob[keep1]
    topk_idxs = keep1.nonzero(as_tuple=True)[0]
    num_topk = min(pre_nms_topk, topk_idxs.size(0))
    _, idxs = d.sort(descending=False)
    if num_topk > 0 and not bool(idxs[:num_topk].max() > pred_prob.shape[0]):
        pred_prob = pred_prob[idxs[:num_topk]]
        topk_idxs = topk_idxs[idxs[:num_topk]]
    pt_idxs = torch.div(topk_idxs, num_classes, rounding_mode='floor')
    cls_idxs = torch.fmod(topk_idxs, num_classes)
    offsets = offsets_i[pt_idxs]
    pts = pts_i[pt_idxs]
    seg_left = pts[:, 0] - offsets[:, 0] * pts[:, 3]
    seg_right = pts[:, 0] + offsets[:, 1] * pts[:, 3]
    keep2 = (seg_right - seg_left) > duration_thresh
    return torch.stack((seg_left, seg_right), -1)[keep2], pred_prob[keep2], cls_idxs[keep2]
