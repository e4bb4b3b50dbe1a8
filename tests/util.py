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


def kernel_parity_on_own_outputs(cfg, model, video):
    """Pins the decode + soft-NMS kernels to the reference ALGORITHM on the CUDA path's own head outputs of one clip, so that
    the check is independent of the operand rounding of the network in front of them:
      (a) decode kernel vs the oracle's decode of the same logits / offsets: same candidates (counts within 0.2 %: a
          probability within 1 ulp of the 1e-3 threshold may fall on either side; sorted scores within 1e-6);
      (b) oracle batched soft-NMS + post-processing (the C restatement pinned bit-exactly to the reference's extension) on the
          KERNEL's candidates vs the detections the public `model.inference` returns for the same head outputs.
    Returns dict(cand_count, cand_count_oracle, cand_score, score, rank_swaps, seg, orphans)."""
    import torch
    from oracle import mq_oracle as O
    from oracle import nms_c
    cls_l, off_l, msk_l = model([video], is_training=False, get_emb=True)
    msk_l = [m.reshape(m.shape[0], -1) for m in msk_l]
    pyr, pmask, logits, offsets = model._lists_to_pyramid(msk_l, cls_l, off_l)
    res = model.inference([video], pyr, pmask, logits, offsets)[0]
    cs, csc, cl, cc = (t.cpu() for t in model._decode_device(pyr, pmask, logits, offsets))
    topk = int(model.test_pre_nms_topk)
    idx = torch.cat([torch.arange(l * topk, l * topk + int(cc[0, l])) for l in range(cc.shape[1])])
    segs, scores, labels = cs[0, idx], csc[0, idx], cl[0, idx].long()
    pts = O.points(cfg, [m.shape[1] for m in msk_l])
    o_segs, o_scores, o_labels = O.decode_single_video(cfg, pts, [m[0].cpu().bool() for m in msk_l], [l[0].cpu() for l in cls_l],
                                                        [o[0].cpu() for o in off_l])
    n = min(len(scores), len(o_scores))
    a, b = scores.sort(descending=True)[0][:n], o_scores.sort(descending=True)[0][:n]
    s2, sc2, lb2 = O.postprocess(cfg, segs, scores, labels, video["fps"], video["duration"], video["feat_stride"],
                                 video["feat_num_frames"], nms_c.softnms_1d)
    ds, swaps, dseg, orphans = match_detections(res, s2.numpy(), sc2.numpy(), lb2.numpy())
    return dict(cand_count=len(scores), cand_count_oracle=len(o_scores), cand_score=float((a - b).abs().max()) if n else 0.0,
                score=ds, rank_swaps=swaps, seg=dseg, orphans=orphans)


def assert_kernel_parity(r):
    assert abs(r["cand_count"] - r["cand_count_oracle"]) <= max(2, 0.002 * r["cand_count_oracle"]), r
    assert r["cand_score"] < 1e-6, r
    assert r["score"] < 1e-6 and r["rank_swaps"] == 0 and r["orphans"] == 0 and r["seg"] < 1e-4, r


def match_detections(res, g_segs, g_scores, g_labels, seg_tol=2e-3):
    """-> (max score diff by rank,
           number of ranks that hold a different detection (other label, or same label but another segment: a near-tie swap),
           max |segment diff| over the ranks that hold the same detection,
           number of our detections with no (label, segment) partner ANYWHERE in the reference list)."""
    segs, scores, labels = (np.asarray(res[k]) for k in ("segments", "scores", "labels"))
    g_segs, g_scores, g_labels = np.asarray(g_segs), np.asarray(g_scores), np.asarray(g_labels)
    assert segs.shape == g_segs.shape, (segs.shape, g_segs.shape)
    ds = float(np.abs(scores - g_scores).max()) if len(scores) else 0.0
    d = np.abs(segs - g_segs).max(1) if len(scores) else np.zeros(0)
    same = (labels == g_labels) & (d <= seg_tol)
    dseg = float(d[same].max()) if same.any() else 0.0
    orphans = 0
    for s, lb in zip(segs, labels):
        cand = g_segs[g_labels == lb]
        if cand.size == 0 or np.abs(cand - s[None]).max(1).min() > seg_tol:
            orphans += 1
    return ds, int((~same).sum()), dseg, orphans
