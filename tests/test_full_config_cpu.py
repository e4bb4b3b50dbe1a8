"""The oracle at the north-star configuration itself (mq_no_cl.yaml full size: C = 1024, T = 1024, input 4096, 10 levels,
K = 22 and K = 110) against tests/golden/model_full.npz, which the REFERENCE produced (oracle/gen_golden.py full).  CPU."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import mq_oracle as O
from oracle import nms_c
from oracle import params as PR
from oracle.gen_golden import FULL_VIDEOS, full_cfg


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / (np.abs(b).max() + 1e-12)


@pytest.mark.parametrize("K", [22, 110])
def test_oracle_matches_reference_at_full_size(K):
    g = np.load(os.path.join(GOLDEN, "model_full.npz"))
    c = full_cfg(K)
    P = PR.random_state(PR.param_spec(c), 4)
    videos = PR.synth_video_list(c, 2, **FULL_VIDEOS)
    clips = [0, 1] if K == 22 else [1]          # keep the CPU suite short: the ragged clip only at K = 110
    with torch.no_grad():
        res, raw = O.model_infer(P, c, [videos[i] for i in clips], softnms_fn=nms_c.softnms_1d, return_raw=True)
    for j, i in enumerate(clips):
        logits, offs, masks, *_ = raw[j]
        assert (torch.cat(masks, 1)[0].numpy() == g[f"k{K}_masks_{i}"]).all()
        assert _rel(torch.cat(logits, 1)[0].numpy(), g[f"k{K}_logits_{i}"]) < 5e-5
        assert _rel(torch.cat(offs, 1)[0].numpy(), g[f"k{K}_offsets_{i}"]) < 5e-5
        assert res[j]["segments"].shape == g[f"k{K}_det_segments_{i}"].shape
        assert np.abs(res[j]["scores"].numpy() - g[f"k{K}_det_scores_{i}"]).max() < 1e-5
