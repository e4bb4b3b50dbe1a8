import contextlib

import numpy as np
import torch

TOL = 1e-3  # north-star tolerance: logits / offsets / losses within 1e-3 relative (bf16 operands, fp32 accumulate)


def rel_max(a, b):
    """max |a - b| / max |b| — the relative error used throughout the parity tests."""
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def build_pair(cfg, seed=0):
    """(vilco_b200 model on cuda with seeded weights, the same weights as an oracle params dict)."""
    from oracle import params as PR
    from vilco_b200.config import mq_model_kwargs
    from vilco_b200.modeling import make_meta_arch
    P = PR.random_state(PR.param_spec(cfg), seed)
    model = make_meta_arch("LocPointTransformer", **mq_model_kwargs(
        cfg.input_dim, cfg.embd_dim, cfg.n_head, cfg.max_seq_len, cfg.arch, cfg.num_classes, cfg.n_txt_in,
        cfg.regression_range))
    missing, unexpected = model.load_state_dict(P, strict=False)
    assert not unexpected
    return model.cuda().eval(), P


def build_vilco_pair(cfg, seed=1):
    """mq_vilco.yaml-like model (L2P prompts, temporal adapters + EMA copy) with seeded weights."""
    from oracle import params as PR
    from vilco_b200.config import mq_model_kwargs
    from vilco_b200.modeling import make_meta_arch
    P = PR.random_state(PR.param_spec(cfg), seed)
    kw = mq_model_kwargs(cfg.input_dim, cfg.embd_dim, cfg.n_head, cfg.max_seq_len, cfg.arch, cfg.num_classes, cfg.n_txt_in,
                         cfg.regression_range)
    kw["cl_cfg"].update(name="l2p", memory_size=1010, prompt_pool=True, pool_size=cfg.prompt_pool["pool_size"],
                        topk=cfg.prompt_pool["top_k"], length=cfg.prompt_pool["length"], embed_dim=cfg.n_txt_in,
                        narration_ssl=True, use_adapt=True, adapt_blocks=list(cfg.adapt_blocks))
    model = make_meta_arch("LocPointTransformer", **kw)
    missing, unexpected = model.load_state_dict(P, strict=False)
    assert not unexpected
    return model.cuda().eval(), P


def build_vilco_train_pair(cfg, seed=2):
    """mq_vilco.yaml training model (prompts, adapters, narration SSL) with every dropout off — the configuration
    tests/golden/train_vilco.npz was generated with (oracle/gen_golden.py gen_vilco_train_golden)."""
    from oracle import params as PR
    from vilco_b200.config import mq_model_kwargs
    from vilco_b200.modeling import make_meta_arch
    P = PR.random_state(PR.param_spec(cfg), seed)
    kw = mq_model_kwargs(cfg.input_dim, cfg.embd_dim, cfg.n_head, cfg.max_seq_len, cfg.arch, cfg.num_classes, cfg.n_txt_in,
                         cfg.regression_range)
    kw["cl_cfg"].update(name="l2p", prompt_pool=True, pool_size=cfg.prompt_pool["pool_size"], topk=cfg.prompt_pool["top_k"],
                        length=cfg.prompt_pool["length"], embed_dim=cfg.n_txt_in, narration_ssl=True,
                        narration_dim=cfg.narration_dim, memory_size=48, ssl_factor=0.01, use_adapt=True,
                        adapt_blocks=list(cfg.adapt_blocks))
    kw["train_cfg"].update(dropout=0.0, droppath=1e-12)
    model = make_meta_arch("LocPointTransformer", **kw)
    missing, unexpected = model.load_state_dict(P, strict=False)
    assert not unexpected
    model.xl_dropout = 0.0
    return model.cuda(), P


@contextlib.contextmanager
def precision(name):
    """run a block in another operand-format policy (vilco_b200.ops.set_precision)"""
    from vilco_b200 import ops
    prev = ops.precision()
    ops.set_precision(name)
    try:
        yield
    finally:
        ops.set_precision(prev)


def oracle_detections(cfg, video, cls_l, off_l, msk_l):
    """The REFERENCE ALGORITHM's decode + soft-NMS + post-processing (oracle restatement, pinned bit-exactly to the
    reference's own extension) applied to GIVEN head outputs (per-level lists, batch 1).  Comparing it with the CUDA path's
    detections isolates the decode / NMS kernels from the rounding of the network in front of them."""
    from oracle import mq_oracle as O
    from oracle import nms_c
    pts = O.points(cfg, [m.shape[1] for m in msk_l])
    segs, scores, labels = O.decode_single_video(cfg, pts, [m[0].cpu().bool() for m in msk_l], [l[0].cpu() for l in cls_l],
                                                 [o[0].cpu() for o in off_l])
    return O.postprocess(cfg, segs, scores, labels, video["fps"], video["duration"], video["feat_stride"],
                         video["feat_num_frames"], nms_c.softnms_1d)


def match_detections(res, g_segs, g_scores, g_labels, seg_tol=2e-3):
    """-> (max score diff by rank, number of ranks whose label differs, max |segment diff| over same-label ranks,
    number of our detections with no (label, segment) partner anywhere in the reference list)."""
    segs, scores, labels = (np.asarray(res[k]) for k in ("segments", "scores", "labels"))
    g_segs, g_scores, g_labels = np.asarray(g_segs), np.asarray(g_scores), np.asarray(g_labels)
    assert segs.shape == g_segs.shape, (segs.shape, g_segs.shape)
    ds = float(np.abs(scores - g_scores).max()) if len(scores) else 0.0
    same = labels == g_labels
    dseg = float(np.abs(segs[same] - g_segs[same]).max()) if same.any() else 0.0
    orphans = 0
    for s, lb in zip(segs, labels):
        cand = g_segs[g_labels == lb]
        if cand.size == 0 or np.abs(cand - s[None]).max(1).min() > seg_tol:
            orphans += 1
    return ds, int((~same).sum()), dseg, orphans
