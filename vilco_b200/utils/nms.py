"""1-D NMS on the GPU behind the reference's `batched_nms` signature (MQ/libs/utils/nms.py:103-190)."""
import ctypes as C

import torch

from .. import lib as L


def _run(segs, scores, labels, region_count, B, n_regions, region_cap, num_classes, multiclass, method, iou_threshold,
         sigma, min_score, max_seg_num):
    """Device-level call; all tensors CUDA.  Returns (out_segs (B,M,2), out_scores (B,M), out_labels (B,M) i64, out_count (B,) i32)."""
    dev = segs.device
    ncls = num_classes if multiclass else 1
    wsb = L.lib().vilco_nms_workspace_bytes(B, n_regions, region_cap, ncls, max_seg_num)
    ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
    out_segs = torch.zeros(B, max_seg_num, 2, device=dev)
    out_scores = torch.zeros(B, max_seg_num, device=dev)
    out_labels = torch.zeros(B, max_seg_num, device=dev, dtype=torch.int64)
    out_count = torch.zeros(B, device=dev, dtype=torch.int32)
    L.check(L.lib().vilco_batched_nms(
        C.c_void_p(segs.data_ptr()), C.c_void_p(scores.data_ptr()), C.c_void_p(labels.data_ptr()),
        C.c_void_p(region_count.data_ptr()), B, n_regions, region_cap, num_classes, int(multiclass), int(method),
        C.c_float(iou_threshold), C.c_float(sigma), C.c_float(min_score), int(max_seg_num), C.c_void_p(ws.data_ptr()),
        C.c_size_t(wsb), C.c_void_p(out_segs.data_ptr()), C.c_void_p(out_scores.data_ptr()),
        C.c_void_p(out_labels.data_ptr()), C.c_void_p(out_count.data_ptr()), L.stream_ptr()), "vilco_batched_nms")
    return out_segs, out_scores, out_labels, out_count


def batched_nms(segs, scores, cls_idxs, iou_threshold, min_score, max_seg_num, use_soft_nms=True, multiclass=True,
                sigma=0.5, voting_thresh=0.75):
    """Drop-in for libs.utils.batched_nms: (N,2) f32, (N,) f32, (N,) i64 -> sorted (<=max_seg_num) CPU tensors.

    Accepts CPU or CUDA inputs (the reference receives CPU tensors, meta_archs.py:1707-1709); the work happens on
    cuda:0 / the inputs' device.  Soft-NMS is the gaussian variant (method 2) like the reference wrapper (nms.py:139)."""
    n = segs.shape[0]
    if n == 0:  # nms.py:118-121
        return torch.zeros([0, 2]), torch.zeros([0]), torch.zeros([0], dtype=cls_idxs.dtype)
    L.lib().vilco_nms_workspace_bytes.restype = C.c_size_t
    dev = segs.device if segs.is_cuda else torch.device("cuda", torch.cuda.current_device())
    s = segs.to(dev, torch.float32).contiguous()
    sc = scores.to(dev, torch.float32).contiguous()
    lb = cls_idxs.to(dev, torch.int32).contiguous()
    num_classes = int(cls_idxs.max().item()) + 1 if multiclass else 1
    cnt = torch.tensor([n], device=dev, dtype=torch.int32)
    os_, osc, ol, oc = _run(s, sc, lb, cnt, 1, 1, n, num_classes, multiclass, 2 if use_soft_nms else 3, iou_threshold,
                            sigma, min_score, max_seg_num)
    if not multiclass and voting_thresh > 0:      # nms.py:174-181
        seg_voting(os_, oc, s, sc, cnt, 1, 1, n, max_seg_num, voting_thresh)
    k = int(oc.item())
    return os_[0, :k].cpu(), osc[0, :k].cpu(), ol[0, :k].cpu().to(cls_idxs.dtype)


def seg_voting(out_segs, out_count, segs, scores, region_count, B, n_regions, region_cap, max_seg_num, voting_thresh):
    """in-place segment voting of the kept segments against all candidates (device tensors, see vilco_seg_voting)"""
    L.check(L.lib().vilco_seg_voting(
        C.c_void_p(out_segs.data_ptr()), C.c_void_p(out_count.data_ptr()), C.c_void_p(segs.data_ptr()),
        C.c_void_p(scores.data_ptr()), C.c_void_p(region_count.data_ptr()), B, n_regions, region_cap, int(max_seg_num),
        C.c_float(voting_thresh), L.stream_ptr()), "vilco_seg_voting")
