"""TEST INFRASTRUCTURE ONLY — first pin for the NLQ row: the reference's own NLQ model
(`NLQ/libs/modeling`, config ego4d_nlq_v2_egovlp_1e-4.yaml: C = 384, 4 heads of 96, window-9 local attention, text
cross-attention in the video stem, 7 levels) run on seeded weights and seeded synthetic inputs.

Run in the authoring container only, in its own process:   python -m oracle.gen_golden_nlq
Writes
* tests/golden/nlq_state_spec.json — parameter / buffer names and shapes of the FULL-SIZE model (T = 2560): the state_dict
  layout a mirror has to reproduce (467 entries, 30.2 M parameters);
* tests/golden/nlq_small.npz — logits / offsets / masks per level (`get_emb=True`, eval mode) and the final detections for
  two clips at T = 512 (same widths and depth as the real config, shorter sequence), weights from `nlq_random_state`.
"""
import json
import os
import zlib

import numpy as np
import torch

from . import nlq_shim

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def nlq_random_state(shapes, seed=0):
    """seeded O(1) weights as a function of (name, shape) only, so that a mirror can rebuild them without the reference:
    norm weights and per-channel scales around 1 (this also lifts the AffineDropPath scales from their 1e-4 init, which
    would hide every residual branch), biases small, everything else N(0, 1/sqrt(fan_in))."""
    out = {}
    for name, shape in shapes.items():
        g = torch.Generator().manual_seed(seed * 100003 + zlib.crc32(name.encode()))
        leaf = name.rsplit(".", 1)[-1]
        if len(shape) == 0:
            out[name] = torch.tensor(1.0)
        elif "norm" in name or ".ln" in name or leaf == "scale":
            out[name] = (1.0 + 0.1 * torch.randn(shape, generator=g)) if leaf != "bias" else 0.1 * torch.randn(shape, generator=g)
        elif leaf == "bias":
            out[name] = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else int(shape[0])
            out[name] = torch.randn(shape, generator=g) / max(1.0, fan_in) ** 0.5
    return out


def small(cfg):
    cfg["dataset"]["max_seq_len"] = 512


def synth_clips(cfg, n, seed):
    T, Cv, Ct = cfg["dataset"]["max_seq_len"], cfg["dataset"]["input_vid_dim"], cfg["dataset"]["input_txt_dim"]
    out = []
    for i in range(n):
        g = torch.Generator().manual_seed(1000 * seed + i)
        tv = T - 37 * i
        nq = 9 + 4 * i
        out.append({"video_id": f"nlq_{seed}_{i}", "feats": torch.randn(Cv, tv, generator=g),
                    "query_feats": torch.randn(Ct, nq, generator=g), "fps": 30.0, "duration": tv * 16.043 / 30.0,
                    "feat_stride": 16.043, "feat_num_frames": 16.043, "query_id": f"q{i}",
                    "segments": torch.tensor([[10.0 + 5 * i, 60.0 + 9 * i]]), "one_hot_labels": torch.ones(1, 1)})
    return out


def main():
    full, cfg_full = nlq_shim.build_model()
    spec = {k: list(v.shape) for k, v in full.state_dict().items()}
    with open(os.path.join(GOLDEN, "nlq_state_spec.json"), "w") as f:
        json.dump({"config": "ego4d_nlq_v2_egovlp_1e-4.yaml", "n_parameters": sum(p.numel() for p in full.parameters()),
                   "state_dict": spec}, f, indent=0)
    print("full-size state_dict:", len(spec), "entries,", sum(p.numel() for p in full.parameters()), "parameters")
    del full

    model, cfg = nlq_shim.build_model(small)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items() if torch.is_floating_point(v)}
    missing, unexpected = model.load_state_dict(nlq_random_state(shapes, 0), strict=False)
    assert not unexpected and not [k for k in missing if k in shapes], (missing, unexpected)
    model.eval()
    out = {}
    with torch.no_grad():
        for i, clip in enumerate(synth_clips(cfg, 2, 0)):
            logits, offsets, masks = model([clip], is_training=False, get_emb=True)
            for l in range(len(logits)):
                out[f"logits_{i}_{l}"] = logits[l][0].numpy()
                out[f"offsets_{i}_{l}"] = offsets[l][0].numpy()
                out[f"mask_{i}_{l}"] = masks[l][0].numpy()
            res = model([clip], is_training=False)[0]
            for k in ("segments", "scores", "labels"):
                out[f"det_{k}_{i}"] = res[k].numpy()
    # training losses (meta_archs.py:746-776, 1094-1155) of the two clips as one batch: train mode with every dropout /
    # drop-path probability set to 0 (the scales of AffineDropPath stay), so that the pass is deterministic
    model.train()
    for m in model.modules():
        if hasattr(m, "drop_prob"):
            m.drop_prob = 0.0
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    out["loss_normalizer_before"] = np.float32(model.loss_normalizer)
    losses = model(synth_clips(cfg, 2, 0), is_training=True)
    for k, v in losses.items():
        out["loss_" + k] = np.float32(v.detach().item())
    out["loss_normalizer_after"] = np.float32(model.loss_normalizer)
    print({k: float(v) for k, v in out.items() if k.startswith("loss_")})
    np.savez_compressed(os.path.join(GOLDEN, "nlq_small.npz"), **out)
    print({k: v.shape for k, v in out.items() if k.endswith(("_0_0", "_0_6")) or k.startswith("det_")})
    print("logit range", float(out["logits_0_0"].min()), float(out["logits_0_0"].max()),
          "size", os.path.getsize(os.path.join(GOLDEN, "nlq_small.npz")))


if __name__ == "__main__":
    main()
