"""The oracle (oracle/mq_oracle.py) against the golden vectors produced by the REFERENCE itself
(oracle/gen_golden.py, run where /root/reference exists).  CPU only."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import mq_oracle as O
from oracle import params as PR
from oracle.gen_golden import small_cfg


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / (np.abs(b).max() + 1e-12)


@pytest.fixture(scope="module")
def small():
    c = small_cfg()
    P = PR.random_state(PR.param_spec(c), 0)
    videos = PR.synth_video_list(c, 2, seed=0, lens=[128, 100], text_lens=[40, 57], n_gt=[3, 2])
    g = np.load(os.path.join(GOLDEN, "model_small.npz"))
    return c, P, videos, g


def test_model_logits_offsets(small):
    c, P, videos, g = small
    with torch.no_grad():
        res, raw = O.model_infer(P, c, videos, return_raw=True)
    for i in range(2):
        logits, offs, masks, *_ = raw[i]
        assert (torch.cat(masks, 1)[0].numpy() == g[f"masks_{i}"]).all()
        assert _rel(torch.cat(logits, 1)[0].numpy(), g[f"logits_{i}"]) < 2e-5
        assert _rel(torch.cat(offs, 1)[0].numpy(), g[f"offsets_{i}"]) < 2e-5


def test_model_detections(small):
    c, P, videos, g = small
    with torch.no_grad():
        res = O.model_infer(P, c, videos)
    for i in range(2):
        # the oracle's logits differ from the reference's in the last bits (different op order), so the
        # candidate sets can differ at threshold edges: compare the detections as sets with tolerance
        assert res[i]["segments"].shape == g[f"det_segments_{i}"].shape
        assert np.abs(res[i]["scores"].numpy() - g[f"det_scores_{i}"]).max() < 1e-5
        same = (res[i]["labels"].numpy() == g[f"det_labels_{i}"]).mean()
        assert same > 0.98
        ok = res[i]["labels"].numpy() == g[f"det_labels_{i}"]
        assert np.abs(res[i]["segments"].numpy()[ok] - g[f"det_segments_{i}"][ok]).max() < 1e-2


def test_model_losses(small):
    c, P, videos, g = small
    with torch.no_grad():
        losses, ln = O.model_train_losses(P, c, videos)
    for k in ("cls_loss", "reg_loss", "al_loss", "final_loss"):
        assert abs(float(losses[k]) - float(g["loss_" + k])) <= 2e-5 * max(1.0, abs(float(g["loss_" + k]))), k


@pytest.mark.parametrize("name,window", [("w9", 9), ("w5", 5)])
def test_local_attention(name, window):
    g = np.load(os.path.join(GOLDEN, "local_attn.npz"))
    P = {k[len(name) + 3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(name + "_p_")}
    x = torch.from_numpy(g[name + "_x"])
    valid = torch.from_numpy(g[name + "_valid"])
    T = x.shape[-1]
    mask = (torch.arange(T)[None, :] < valid[:, None]).unsqueeze(1)
    y, _ = O.local_masked_mhca(P, "", x, mask, 2, window)
    assert _rel(y.numpy(), g[name + "_y"]) < 1e-5


CASES = ["n1", "n17", "n300", "n2000", "ties"]
SETTINGS = [(2, 0.99, 1e-4), (2, 0.5, 0.01), (1, 0.5, 0.001), (0, 0.5, 0.001)]


@pytest.mark.parametrize("name", ["n1", "n17", "n300", "ties"])
def test_softnms_numpy_restatement(name):
    g = np.load(os.path.join(GOLDEN, "nms.npz"))
    for method, sigma, ms in SETTINGS:
        dets, inds = O.softnms_1d(g[name + "_segs"], g[name + "_scores"], 0.1, sigma, ms, method)
        tag = f"{name}_m{method}_s{sigma}_t{ms}"
        assert (inds == g[tag + "_inds"]).all(), tag
        assert (dets == g[tag + "_dets"]).all(), tag  # bit-exact


@pytest.mark.parametrize("name", CASES)
def test_hard_nms_restatement(name):
    g = np.load(os.path.join(GOLDEN, "nms.npz"))
    keep = O.nms_1d(g[name + "_segs"], g[name + "_scores"], 0.4)
    assert (keep == g[name + "_hard_keep"]).all()


@pytest.mark.parametrize("name", CASES)
def test_softnms_c_restatement(name):
    from oracle import nms_c
    g = np.load(os.path.join(GOLDEN, "nms.npz"))
    for method, sigma, ms in SETTINGS:
        dets, inds = nms_c.softnms_1d(g[name + "_segs"], g[name + "_scores"], 0.1, sigma, ms, method)
        tag = f"{name}_m{method}_s{sigma}_t{ms}"
        assert (inds == g[tag + "_inds"]).all(), tag
        assert (dets == g[tag + "_dets"]).all(), tag  # bit-exact


def test_batched_nms_restatement():
    from oracle import nms_c
    g = np.load(os.path.join(GOLDEN, "nms.npz"))
    s, sc, lb = O.batched_nms(torch.from_numpy(g["b_segs"]), torch.from_numpy(g["b_scores"]),
                              torch.from_numpy(g["b_labels"]), 0.1, 1e-4, 200, True, True, 0.99, 0.9,
                              softnms_fn=nms_c.softnms_1d)
    assert (s.numpy() == g["b_out_segs"]).all()
    assert (sc.numpy() == g["b_out_scores"]).all()
    assert (lb.numpy() == g["b_out_labels"]).all()


def test_vilco_config_oracle_vs_reference_golden():
    """mq_vilco.yaml branches at inference (L2P prompts, adapters, EMA-adapter ensemble) — golden from the reference."""
    from oracle.gen_golden import vilco_cfg
    g = np.load(os.path.join(GOLDEN, "model_vilco.npz"))
    c = vilco_cfg()
    P = PR.random_state(PR.param_spec(c), 1)
    videos = PR.synth_video_list(c, 1, seed=5, lens=[900], text_lens=[57], n_gt=[3])
    with torch.no_grad():
        res, raw = O.model_infer(P, c, videos, return_raw=True)
    assert _rel(torch.cat(raw[0][0], 1)[0].numpy(), g["logits_0"]) < 2e-5
    assert _rel(torch.cat(raw[0][1], 1)[0].numpy(), g["offsets_0"]) < 2e-5
    assert np.abs(res[0]["scores"].numpy() - g["det_scores_0"]).max() < 1e-5


def test_oracle_gradients_match_reference_autograd():
    """d final_loss / d every parameter: torch autograd through the oracle vs the reference's own autograd
    (tests/golden/grads_small.npz, generated by oracle/gen_golden.py gen_grad_golden)."""
    import torch
    from oracle import gen_golden as GG
    g = np.load(os.path.join(GOLDEN, "grads_small.npz"))
    cfg = GG.small_cfg()
    P = {k: v.clone().requires_grad_(True) for k, v in PR.random_state(PR.param_spec(cfg), 0).items()}
    videos = PR.synth_video_list(cfg, 2, seed=0, lens=[128, 100], text_lens=[40, 57], n_gt=[3, 2])
    lo, _ = O.model_train_losses(P, cfg, videos)
    lo["final_loss"].backward()
    assert abs(float(lo["final_loss"]) - float(g["final_loss"])) < 1e-5
    n = 0
    for key in g.files:
        if not key.startswith("g:"):
            continue
        ref = g[key]
        mine = P[key[2:]].grad.reshape(-1).double()
        got = np.concatenate([[mine.norm().item(), mine.sum().item()], mine[:8].numpy()])
        scale = max(ref[0], 1e-12)
        assert np.abs(got - ref).max() <= 2e-4 * scale + 1e-7, (key, got[:3], ref[:3])
        n += 1
    assert n == 338
