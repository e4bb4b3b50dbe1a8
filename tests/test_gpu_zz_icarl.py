"""iCaRL nearest-mean re-scoring through the public model API (classify / forward(..., val_qilDatasetList=...)) on the CUDA
path against goldens produced by the reference's own classify + inference (tests/golden/icarl_small.npz,
oracle/gen_golden_icarl.py): distance tables within 1e-3 relative, detections equal up to the few candidates whose distance
sits within rounding of the selection threshold (distance < mean of the table)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from util import build_pair, rel_max

pytestmark = pytest.mark.gpu
PRECISION = "fp16x3"   # the tolerances below state the exact (split-operand) arithmetic; see conftest._precision_mode


def test_icarl_rescoring_vs_reference_golden():
    from oracle.gen_golden_icarl import FakeTask, icarl_cfg, memory_and_clips
    g = np.load(os.path.join(GOLDEN, "icarl_small.npz"))
    cfg = icarl_cfg()
    model, _ = build_pair(cfg)
    memory, clips = memory_and_clips(cfg)
    for i, clip in enumerate(clips):
        model.memory = memory
        model.compute_means = True
        dists = model.classify(clip, FakeTask())
        assert model.compute_means is False and len(dists) == 10
        for l, d in enumerate(dists):
            assert d.shape == (1,) + g[f"dists_{i}_{l}"].shape
            assert rel_max(d[0], g[f"dists_{i}_{l}"]) < 1e-3, (i, l)
        if i == 0:
            for l in range(10):
                m = torch.stack(model.exemplar_means[l], 0)
                assert rel_max(m.flatten(1).norm(dim=1), g[f"means_norm_{l}"]) < 1e-5
                assert rel_max(m[:, :4, :2], g[f"means_head_{l}"]) < 1e-3
        model.compute_means = True
        with torch.no_grad():
            res = model([clip], is_training=False, val_qilDatasetList=FakeTask())[0]
            plain = model([clip], is_training=False, val_qilDatasetList=FakeTask())[0]    # flag cleared: the kernel path
        assert model.compute_means is False
        assert res["segments"].device.type == "cpu" and res["labels"].dtype == torch.int64
        assert np.abs(plain["scores"].numpy() - g[f"plain_scores_{i}"]).max() < 1e-5
        gs, gsc, gl = g[f"det_segments_{i}"], g[f"det_scores_{i}"], g[f"det_labels_{i}"]
        s, sc, lb = res["segments"].numpy(), res["scores"].numpy(), res["labels"].numpy()
        assert abs(len(sc) - len(gsc)) <= 2
        hit = 0
        for j in range(len(gsc)):
            ok = (lb == gl[j]) & (np.abs(sc - gsc[j]) < 1e-5) & (np.abs(s - gs[j]).max(1) < 5e-2)
            hit += bool(ok.any())
        assert hit / len(gsc) > 0.75, (i, hit, len(gsc))


def test_reference_style_ensemble_inference_call():
    """infer_one_epoch_ensemble's calling form (train_utils.py:945-961): forward(ensemble=True) returns the per-level lists,
    the caller averages them over models and hands them to model.inference(video_list, points, masks, cls, offs, None, None).
    Averaging two copies of the same outputs is the identity, so the detections must equal the plain evaluation call."""
    from oracle import params as PR
    from oracle.gen_golden import small_cfg
    cfg = small_cfg()
    model, _ = build_pair(cfg)
    videos = PR.synth_video_list(cfg, 2, seed=0, lens=[128, 100], text_lens=[40, 57], n_gt=[3, 2])
    with torch.no_grad():
        want = model(videos, is_training=False)
        vl, points, masks, cls_l, off_l = model(videos, ensemble=True, is_training=False)
        assert len(points) == len(cls_l) == len(off_l) == len(masks) and cls_l[0].shape[:2] == (2, 128)
        cls_avg = [(a + a) / 2.0 for a in cls_l]
        off_avg = [(a + a) / 2.0 for a in off_l]
        got = model.inference(vl, points, masks, cls_avg, off_avg, None, None)
    assert len(got) == len(want) == 2
    for a, b in zip(got, want):
        assert a["video_id"] == b["video_id"]
        assert torch.equal(a["scores"], b["scores"]) and torch.equal(a["labels"], b["labels"])
        assert torch.equal(a["segments"], b["segments"])
