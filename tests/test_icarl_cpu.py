"""iCaRL nearest-mean re-scoring glue (vilco_b200/modeling/icarl.py) on CPU: the candidate selection fed with the reference's
own logits / offsets / masks / distance tables (tests/golden/icarl_small.npz, oracle/gen_golden_icarl.py) followed by the
oracle's soft-NMS must reproduce the reference's detections exactly; the distance / mean expressions are checked against a
literal restatement of meta_archs.py:1098-1127 on small tensors."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import mq_oracle as O
from oracle import nms_c
from oracle.gen_golden_icarl import icarl_cfg
from vilco_b200.modeling import icarl


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "icarl_small.npz"))


def test_candidate_selection_reproduces_reference_detections(g):
    cfg = icarl_cfg()
    T = cfg.max_seq_len
    lens = [T >> l for l in range(icarl.FPN_LEVELS)]
    pts = O.points(cfg, lens)
    for i in range(2):
        segs, scores, labels = [], [], []
        for l in range(icarl.FPN_LEVELS):
            s, sc, lb = icarl.select_candidates(
                torch.from_numpy(g[f"logits_{i}_{l}"]), torch.from_numpy(g[f"offsets_{i}_{l}"]), pts[l],
                torch.from_numpy(g[f"mask_{i}_{l}"]).reshape(-1), torch.from_numpy(g[f"dists_{i}_{l}"])[None],
                cfg.num_classes, cfg.pre_nms_topk, cfg.duration_thresh)
            segs.append(s), scores.append(sc), labels.append(lb)
        stride = 480.0 * 30.0 / T
        s, sc, lb = O.postprocess(cfg, torch.cat(segs), torch.cat(scores), torch.cat(labels), 30.0, 480.0, stride, stride,
                                  softnms_fn=nms_c.softnms_1d)
        assert np.array_equal(sc.numpy(), g[f"det_scores_{i}"]), i
        assert np.array_equal(lb.numpy(), g[f"det_labels_{i}"]) and np.array_equal(s.numpy(), g[f"det_segments_{i}"])
        assert not np.array_equal(g[f"det_scores_{i}"], g[f"plain_scores_{i}"])      # the re-scoring changes the result


def test_second_branch_of_the_selection_indexes_the_filtered_arrays():
    """when the `num_topk` smallest distances all sit at small flat indices the reference re-orders the FILTERED arrays by
    indices of the UNFILTERED sort (meta_archs.py:1641-1643)."""
    T, K = 8, 2
    d = torch.full((1, T, K), 5.0)
    d[0, 0, 0], d[0, 0, 1], d[0, 1, 0] = 1.0, 0.5, 0.8            # below the mean: flat indices 0, 1, 2
    logits = torch.arange(T * K, dtype=torch.float32).reshape(T, K) / 10 - 0.5
    offs = torch.ones(T, 2)
    pts = torch.stack([torch.arange(T, dtype=torch.float32), torch.zeros(T), torch.full((T,), 9.0), torch.ones(T)], 1)
    s, sc, lb = icarl.select_candidates(logits, offs, pts, torch.ones(T), d, K, 5000, 0.01)
    prob = logits.sigmoid().flatten()
    assert torch.equal(sc, prob[[1, 2, 0]])                           # ascending distance: flat 1 (0.5), 2 (0.8), 0 (1.0)
    assert lb.tolist() == [1, 0, 0] and s[:, 0].tolist() == [-1.0, 0.0, -1.0]


def test_distance_and_mean_expressions_match_the_reference_form():
    torch.manual_seed(0)
    C, T, n_cls, n_ex = 16, 12, 5, 3
    feats = [[torch.randn(1, C, T) for _ in range(n_ex)] for _ in range(n_cls)]
    means = [icarl.exemplar_mean([icarl.normalize_level(f) for f in fs]) for fs in feats]
    # literal form of meta_archs.py:1085-1089
    for fs, mu in zip(feats, means):
        ref = torch.stack([f / f.norm() for f in fs], 0).mean(0).squeeze()
        ref = ref / ref.norm()
        assert torch.allclose(mu, ref, rtol=0, atol=1e-7)
    x = torch.randn(1, C, T)
    got = icarl.nme_dists(x, means)
    # literal form of :1098-1127
    m = torch.stack(means, 0)                                        # (n_classes, C, T)
    m = torch.stack([m] * 1).permute(0, 2, 3, 1)                     # (1, C, T, n_classes)
    f = (x / x.norm()).unsqueeze(3).expand_as(m)
    want = (f - m).pow(2).sum(1).squeeze().unsqueeze(0)
    assert got.shape == want.shape == (1, T, n_cls)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-9)
