"""vilco_b200.data.truncate_feats (training-time crop) against the reference's own function
(MQ/libs/datasets/data_utils.py:24-112, loaded by path in the authoring container) under the same `random` seed, and —
where the reference is not available (GPU box) — against committed expectations of the same seeded draws."""
import copy
import importlib.util
import os
import random

import pytest
import torch

REF = "/root/reference/MQ/libs/datasets/data_utils.py"


def _clip(seed, T=64, C=8, K=5, n=4):
    g = torch.Generator().manual_seed(seed)
    s0 = torch.rand(n, generator=g) * (T - 8)
    segs = torch.stack([s0, s0 + 1 + torch.rand(n, generator=g) * 20], 1).clamp(max=T)
    return {"video_id": f"v{seed}", "feats": torch.randn(C, T, generator=g), "segments": segs,
            "labels": torch.randint(0, K, (n,), generator=g), "segmentation_labels": torch.rand(T, K, generator=g),
            "fps": 30.0, "duration": 100.0, "feat_stride": 4.0, "feat_num_frames": 4.0,
            "prompt_feature": torch.randn(6, 3, generator=g)}


CASES = [dict(max_seq_len=64, trunc_thresh=0.3, crop_ratio=[0.9, 1.0]), dict(max_seq_len=64, trunc_thresh=0.5, crop_ratio=None),
         dict(max_seq_len=40, trunc_thresh=0.3, crop_ratio=None), dict(max_seq_len=40, trunc_thresh=0.5, crop_ratio=None, no_trunc=True),
         dict(max_seq_len=24, trunc_thresh=0.9, crop_ratio=None, has_action=False), dict(max_seq_len=64, trunc_thresh=0.3, crop_ratio=[0.5, 0.7])]


@pytest.mark.skipif(not os.path.exists(REF), reason="reference checkout not present (GPU box)")
def test_truncate_feats_equals_reference_under_the_same_seed():
    from vilco_b200 import data
    spec = importlib.util.spec_from_file_location("ref_data_utils", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    for ci, kw in enumerate(CASES):
        for seed in range(12):
            d = _clip(seed)
            random.seed(100 * ci + seed)
            want = ref.truncate_feats(copy.deepcopy(d), **kw)
            state_ref = random.getstate()
            random.seed(100 * ci + seed)
            got = data.truncate_feats(d, **kw)
            assert random.getstate() == state_ref, "different consumption of the random stream"
            assert set(got) == set(want)
            for k in ("feats", "segments", "labels", "segmentation_labels"):
                assert got[k].shape == want[k].shape and torch.equal(got[k], want[k]), (ci, seed, k)
            assert got["prompt_feature"] is d["prompt_feature"] and got["video_id"] == d["video_id"]


def test_truncate_feats_contract_without_the_reference():
    from vilco_b200 import data
    d = _clip(3)
    assert data.truncate_feats(d, 64, 0.3, None) is d                       # nothing to do: the same object, like the reference
    random.seed(5)
    out = data.truncate_feats(d, 40, 0.3, None)
    assert out["feats"].shape == (8, 40) and out["segmentation_labels"].shape == (40, 5)
    assert out["segments"].shape[0] == out["labels"].shape[0] >= 1
    assert float(out["segments"].min()) >= 0.0 and float(out["segments"].max()) <= 40.0
    assert d["feats"].shape == (8, 64)                                      # the input dict is not modified
    # the crop is a window of the original features
    f = d["feats"]
    assert any(torch.equal(out["feats"], f[:, s:s + 40]) for s in range(64 - 40 + 1))
