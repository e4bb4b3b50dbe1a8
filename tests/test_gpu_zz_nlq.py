"""NLQ model mirror (evaluation path) on the B200 against the goldens produced by the reference's own NLQ
model (tests/golden/nlq_small.npz, oracle/gen_golden_nlq.py): real widths and depth (C = 384, 4 heads of 96, window-9 local
attention, text cross-attention in the video stem, 7 levels), T = 512, two clips."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from util import precision, rel_max

pytestmark = pytest.mark.gpu

NLQ_TEST_CFG = dict(voting_thresh=0.9, pre_nms_topk=2000, max_seg_num=5, min_score=0.001, nms_sigma=0.75, duration_thresh=0.001)
REG_RANGE = [[0, 4], [2, 8], [4, 16], [8, 32], [16, 64], [32, 128], [64, 10000]]


def _build():
    from oracle.gen_golden_nlq import nlq_random_state, synth_clips
    from vilco_b200.modeling import make_meta_arch
    model = make_meta_arch("NlqLocPointTransformer", max_seq_len=512, regression_range=REG_RANGE, test_cfg=NLQ_TEST_CFG).cuda().eval()
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    missing, unexpected = model.load_state_dict(nlq_random_state(shapes, 0), strict=True)
    clips = synth_clips({"dataset": {"max_seq_len": 512, "input_vid_dim": 256, "input_txt_dim": 512}}, 2, 0)
    return model, clips


@pytest.mark.parametrize("mode,tol", [("fp16x3", 1e-4), ("mixed", 5e-2)])
def test_nlq_logits_offsets_vs_reference_golden(mode, tol):
    """the model's own operand policy (exact split fp16) to 1e-4 — measured 1e-5 / 2e-5; the MQ-tuned single-plane policy only as a
    documented bound (measured 1.4e-3 / 1.7e-2: not shipped for this model)"""
    g = np.load(os.path.join(GOLDEN, "nlq_small.npz"))
    model, clips = _build()
    model.operand_mode = mode
    with precision(mode):
        for i, clip in enumerate(clips):
            logits, offsets, masks = model([clip], is_training=False, get_emb=True)
            torch.cuda.synchronize()
            assert len(logits) == 7
            lg = torch.cat(logits, 1)[0].cpu().numpy()
            of = torch.cat(offsets, 1)[0].cpu().numpy()
            rl = np.concatenate([g[f"logits_{i}_{l}"].reshape(-1, 1) for l in range(7)])
            ro = np.concatenate([g[f"offsets_{i}_{l}"].reshape(-1, 2) for l in range(7)])
            for l in range(7):
                assert (masks[l][0].cpu().numpy().reshape(-1) == g[f"mask_{i}_{l}"].reshape(-1).astype(bool)).all()
            e1, e2 = rel_max(lg, rl), rel_max(of, ro)
            print(f"NLQ {mode} clip {i}: logits {e1:.2e} offsets {e2:.2e}")
            assert e1 < tol and e2 < tol


def test_nlq_detections_vs_reference_golden():
    g = np.load(os.path.join(GOLDEN, "nlq_small.npz"))
    model, clips = _build()
    res = model(clips, is_training=False)              # the model's own operand policy; ragged query lengths: clip by clip inside
    for i, r in enumerate(res):
        seg, sc = r["segments"].numpy(), r["scores"].numpy()
        assert seg.shape == g[f"det_segments_{i}"].shape == (5, 2)
        order = np.argsort(-sc, kind="stable")
        ref_order = np.argsort(-g[f"det_scores_{i}"], kind="stable")
        assert np.abs(sc[order] - g[f"det_scores_{i}"][ref_order]).max() < 1e-4
        assert np.abs(seg[order] - g[f"det_segments_{i}"][ref_order]).max() < 2e-3
        assert (r["labels"].numpy() == 0).all()


def test_nlq_eval_graph_equals_eager():
    """the captured NLQ evaluation step (NlqEvalGraph) returns exactly what the eager call returns, also after the weights change"""
    from vilco_b200.modeling.nlq import NlqEvalGraph
    model, clips = _build()
    batch = [dict(clips[0]), dict(clips[0])]
    batch[1]["feats"] = clips[1]["feats"]                      # two clips of different length, one query length
    batch[1]["duration"], batch[1]["video_id"] = clips[1]["duration"], "second"
    eager = model(batch, is_training=False)
    g = NlqEvalGraph(model, batch_size=2, text_len=batch[0]["query_feats"].shape[-1])
    for _ in range(2):
        got = g.run(batch)
        for a, b in zip(got, eager):
            assert torch.equal(a["segments"], b["segments"]) and torch.equal(a["scores"], b["scores"])
            assert torch.equal(a["labels"], b["labels"]) and a["video_id"] == b["video_id"]
    with torch.no_grad():
        model.cls_head.cls_head.conv.bias.add_(0.5)
    eager2 = model(batch, is_training=False)
    got2 = g.run(batch)
    assert g.launches > 100
    for a, b in zip(got2, eager2):
        assert torch.equal(a["scores"], b["scores"]) and not torch.equal(a["scores"], eager[0]["scores"])
