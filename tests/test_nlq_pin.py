"""First pin of the NLQ row (not built yet — docs/NEXT_NLQ.md): the committed fixtures produced by the
reference's own NLQ model (oracle/gen_golden_nlq.py) are well-formed, the seeded weight recipe is a pure function of
(name, shape, seed), and — in the authoring container, in a subprocess because NLQ's package is also called `libs` — the
reference still reproduces the committed vectors."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT


def test_state_spec_and_golden_are_well_formed():
    spec = json.load(open(os.path.join(GOLDEN, "nlq_state_spec.json")))
    sd = spec["state_dict"]
    assert len(sd) == 467 and spec["n_parameters"] == 30168202
    assert sd["backbone.vid_embd.0.conv.weight"] == [384, 256, 3]                      # video 256-d -> C = 384, k = 3
    assert sd["backbone.vid_stem.0.attn.query.weight"][:2] == [384, 384]                # 4 heads of 96
    assert any(k.startswith("backbone.vid_stem.3.cross_attn.") for k in sd)              # text cross-attention in the video stem
    g = np.load(os.path.join(GOLDEN, "nlq_small.npz"))
    lens = [512 >> l for l in range(7)]
    for i in range(2):
        for l, n in enumerate(lens):
            assert g[f"logits_{i}_{l}"].shape == (n, 1) and g[f"offsets_{i}_{l}"].shape == (n, 2) and g[f"mask_{i}_{l}"].shape == (n,)
            assert np.isfinite(g[f"logits_{i}_{l}"]).all() and (g[f"offsets_{i}_{l}"] >= 0).all()
        assert g[f"det_segments_{i}"].shape == (5, 2) and (np.diff(g[f"det_scores_{i}"]) <= 0).all()     # max_seg_num 5, sorted
    assert g["mask_1_0"].sum() == 512 - 37                                                 # the second clip is 37 features shorter


def test_seeded_weights_are_a_pure_function_of_name_shape_seed():
    from oracle.gen_golden_nlq import nlq_random_state
    shapes = {"a.conv.weight": (4, 3, 3), "a.norm.weight": (4,), "a.norm.bias": (4,), "s.scale": (1, 4, 1), "b.bias": (4,)}
    a, b = nlq_random_state(shapes, 3), nlq_random_state(shapes, 3)
    assert all(torch.equal(a[k], b[k]) for k in shapes)
    assert not torch.equal(a["a.conv.weight"], nlq_random_state(shapes, 4)["a.conv.weight"])
    assert abs(float(a["a.norm.weight"].mean()) - 1.0) < 0.3 and abs(float(a["s.scale"].mean()) - 1.0) < 0.3


@pytest.mark.skipif(not os.path.isdir("/root/reference/NLQ/libs/modeling"), reason="NLQ reference not present (GPU box)")
def test_reference_reproduces_the_committed_vectors():
    code = (
        "import numpy as np, torch, os\n"
        "from oracle import nlq_shim\n"
        "from oracle.gen_golden_nlq import nlq_random_state, small, synth_clips, GOLDEN\n"
        "m, cfg = nlq_shim.build_model(small)\n"
        "shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if torch.is_floating_point(v)}\n"
        "m.load_state_dict(nlq_random_state(shapes, 0), strict=False); m.eval()\n"
        "g = np.load(os.path.join(GOLDEN, 'nlq_small.npz'))\n"
        "with torch.no_grad():\n"
        "    lg, of, mk = m([synth_clips(cfg, 2, 0)[1]], is_training=False, get_emb=True)\n"
        "err = max(float(np.abs(lg[l][0].numpy() - g[f'logits_1_{l}']).max()) for l in range(7))\n"
        "print('MAXERR', err)\n")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    err = float([ln for ln in out.stdout.splitlines() if ln.startswith("MAXERR")][-1].split()[1])
    assert err < 1e-5, err


def test_nlq_oracle_matches_reference_golden():
    """oracle/nlq_oracle.py (the MQ operator restatements re-composed for NLQ: window-9 attention live, head dim 96, text
    cross-attention in the video stem) against the reference's own NLQ forward — logits / offsets within 2e-5, masks equal."""
    from oracle import nlq_oracle as N
    from oracle.gen_golden_nlq import nlq_random_state, synth_clips
    spec = json.load(open(os.path.join(GOLDEN, "nlq_state_spec.json")))["state_dict"]
    shapes = {k: tuple(v) for k, v in spec.items() if not k.endswith("num_batches_tracked")}
    g = np.load(os.path.join(GOLDEN, "nlq_small.npz"))
    cfg = N.NlqCfg(max_seq_len=512)
    # the small golden model differs from the full-size spec only in buffers that depend on T (none are parameters)
    P = nlq_random_state({k: v for k, v in shapes.items()}, 0)
    clips = synth_clips({"dataset": {"max_seq_len": 512, "input_vid_dim": 256, "input_txt_dim": 512}}, 2, 0)
    with torch.no_grad():
        for i, clip in enumerate(clips):
            logits, offsets, masks = N.forward_heads(P, cfg, *N.preprocess_eval(cfg, clip))
            assert len(logits) == 7
            for l in range(7):
                assert np.array_equal(masks[l][0].numpy(), g[f"mask_{i}_{l}"])
                assert np.abs(logits[l][0].numpy() - g[f"logits_{i}_{l}"]).max() < 2e-5, (i, l)
                assert np.abs(offsets[l][0].numpy() - g[f"offsets_{i}_{l}"]).max() < 2e-5 * max(1.0, np.abs(g[f"offsets_{i}_{l}"]).max()), (i, l)


def test_nlq_oracle_detections_match_reference_golden():
    """decode + soft-NMS (sigma 0.75, 5 segments per query) of the oracle on its own logits == the reference's detections."""
    from oracle import nlq_oracle as N
    from oracle import nms_c
    from oracle.gen_golden_nlq import nlq_random_state, synth_clips
    spec = json.load(open(os.path.join(GOLDEN, "nlq_state_spec.json")))["state_dict"]
    P = nlq_random_state({k: tuple(v) for k, v in spec.items()}, 0)
    g = np.load(os.path.join(GOLDEN, "nlq_small.npz"))
    cfg = N.NlqCfg(max_seq_len=512)
    clips = synth_clips({"dataset": {"max_seq_len": 512, "input_vid_dim": 256, "input_txt_dim": 512}}, 2, 0)
    with torch.no_grad():
        for i, clip in enumerate(clips):
            s, sc, lb = N.infer(P, cfg, clip, softnms_fn=nms_c.softnms_1d)
            assert s.shape == (5, 2) and np.array_equal(lb.numpy(), g[f"det_labels_{i}"])
            assert np.abs(sc.numpy() - g[f"det_scores_{i}"]).max() < 1e-5
            assert np.abs(s.numpy() - g[f"det_segments_{i}"]).max() < 1e-3


def test_nlq_oracle_training_losses_match_reference_golden():
    """label assignment (centre sampling, radius 1.5) + focal (label smoothing 0.1) + DIoU of a two-clip batch with ragged
    video and text lengths, deterministic train mode — against the reference's own losses."""
    from oracle import nlq_oracle as N
    from oracle.gen_golden_nlq import nlq_random_state, synth_clips
    spec = json.load(open(os.path.join(GOLDEN, "nlq_state_spec.json")))["state_dict"]
    P = nlq_random_state({k: tuple(v) for k, v in spec.items()}, 0)
    g = np.load(os.path.join(GOLDEN, "nlq_small.npz"))
    cfg = N.NlqCfg(max_seq_len=512)
    clips = synth_clips({"dataset": {"max_seq_len": 512, "input_vid_dim": 256, "input_txt_dim": 512}}, 2, 0)
    with torch.no_grad():
        losses, norm = N.train_losses(P, cfg, clips, loss_normalizer=float(g["loss_normalizer_before"]))
    assert abs(norm - float(g["loss_normalizer_after"])) < 1e-3
    for k in ("cls_loss", "reg_loss", "final_loss"):
        ref = float(g["loss_" + k])
        assert abs(float(losses[k]) - ref) <= 2e-5 * max(1.0, abs(ref)), (k, float(losses[k]), ref)


def test_nlq_mirror_state_dict_equals_reference_layout():
    """vilco_b200.modeling.nlq.NlqPtTransformer registers exactly the reference NLQ model's state_dict: names, order, shapes
    (tests/golden/nlq_state_spec.json, written from the reference's own model) — a reference checkpoint loads strictly."""
    from vilco_b200.modeling import make_meta_arch
    spec = json.load(open(os.path.join(GOLDEN, "nlq_state_spec.json")))
    m = make_meta_arch("NlqLocPointTransformer", regression_range=[[0, 4], [2, 8], [4, 16], [8, 32], [16, 64], [32, 128], [64, 10000]])
    sd = m.state_dict()
    assert list(sd.keys()) == list(spec["state_dict"].keys())
    assert all(list(sd[k].shape) == spec["state_dict"][k] for k in sd)
    assert sum(p.numel() for p in m.parameters()) == spec["n_parameters"]
    assert m.fpn_strides == [1, 2, 4, 8, 16, 32, 64] and m.mha_win_size == [9] * 7 and m.max_div_factor == 512
    with pytest.raises(NotImplementedError):
        m([], is_training=True)
