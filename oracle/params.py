"""TEST INFRASTRUCTURE ONLY — state_dict spec (reference key names / shapes) and a
version-stable seeded filler (numpy legacy RandomState), so the reference, the oracle and the CUDA path can
be loaded with bit-identical weights on any box without shipping weight files.

Values are deliberately O(1) where the reference initialises to ~0 (AffineDropPath.scale = 1e-4, zero biases,
unit LN): at default init every residual branch contributes ~1e-4 of the stream and parity would be blind to a
wrong attention / MLP / XLNet kernel (SURVEY.md §4 "test-design trap").
"""
import numpy as np
import torch


def param_spec(cfg):
    """key -> shape for the live parameters of PtTransformer (cl_cfg.name None).  Dead parameters the reference
    also registers (xlnet.word_embedding / mask_emb / seg_embed / r_s_bias, channel_attn.norm1, the unused
    cross_attn / channel_attn of some blocks) are listed too when cheap so state_dict loading is strict-ish;
    the 32000 x C word embedding is omitted."""
    C, Cin, K, Ct = cfg.embd_dim, cfg.input_dim, cfg.num_classes, cfg.n_txt_in
    H = cfg.n_head
    s = {}
    for n in ("mu", "sigma", "mu_reg_left", "sigma_reg_left", "mu_reg_right", "sigma_reg_right"):
        s[n] = (K, 1)
    b = "backbone."
    s[b + "proj.0.conv.weight"] = (C, Cin, 1)
    s[b + "proj.0.conv.bias"] = (C,)
    for i in range(cfg.arch[0]):
        s[b + f"embd.{i}.conv.weight"] = (C, C, 3)
        s[b + f"embd_norm.{i}.weight"] = (1, C, 1)
        s[b + f"embd_norm.{i}.bias"] = (1, C, 1)
        if cfg.use_cross_modal:
            s[b + f"txt_embd.{i}.conv.weight"] = (C, Ct if i == 0 else C, 1)
            s[b + f"txt_embd_norm.{i}.weight"] = (1, C, 1)
            s[b + f"txt_embd_norm.{i}.bias"] = (1, C, 1)

    def block(p, cross, chan):
        for ln in ("ln1", "ln2") + (("ln3",) if cross else ()):
            s[p + ln + ".weight"] = (1, C, 1)
            s[p + ln + ".bias"] = (1, C, 1)
        for n in ("query", "key", "value"):
            s[p + f"attn.{n}_conv.conv.weight"] = (C, 1, 3)
            s[p + f"attn.{n}_norm.weight"] = (1, C, 1)
            s[p + f"attn.{n}_norm.bias"] = (1, C, 1)
        for n in ("query", "key", "value", "proj"):
            s[p + f"attn.{n}.weight"] = (C, C, 1)
            s[p + f"attn.{n}.bias"] = (C,)
            if cross:
                s[p + f"cross_attn.{n}.weight"] = (C, C, 1)
                s[p + f"cross_attn.{n}.bias"] = (C,)
        s[p + "mlp.0.weight"] = (4 * C, C, 1)
        s[p + "mlp.0.bias"] = (4 * C,)
        s[p + "mlp.3.weight"] = (C, 4 * C, 1)
        s[p + "mlp.3.bias"] = (C,)
        s[p + "drop_path_attn.scale"] = (1, C, 1)
        s[p + "drop_path_mlp.scale"] = (1, C, 1)
        if chan:
            s[p + "channel_attn.attn.qkv.weight"] = (3 * C, C)
            s[p + "channel_attn.attn.proj.weight"] = (C, C)
            s[p + "channel_attn.attn.proj.bias"] = (C,)
            s[p + "channel_attn.norm2.weight"] = (C,)
            s[p + "channel_attn.norm2.bias"] = (C,)
            s[p + "channel_attn.mlp.0.weight"] = (4 * C, C)
            s[p + "channel_attn.mlp.0.bias"] = (4 * C,)
            s[p + "channel_attn.mlp.2.weight"] = (C, 4 * C)
            s[p + "channel_attn.mlp.2.bias"] = (C,)

    for i in range(cfg.arch[1]):
        block(b + f"stem.{i}.", False, True)
        if cfg.use_cross_modal:
            block(b + f"txt_stem.{i}.", False, True)
    for i in range(cfg.arch[2]):
        block(b + f"branch.{i}.", cfg.use_cross_modal and i not in (1, 2), False)
    if cfg.use_xl:
        x = b + "xlnet.layer.0."
        d = C // H
        for n in "qkvor":
            s[x + "rel_attn." + n] = (C, H, d)
        s[x + "rel_attn.r_r_bias"] = (H, d)
        s[x + "rel_attn.r_w_bias"] = (H, d)
        s[x + "rel_attn.layer_norm.weight"] = (C,)
        s[x + "rel_attn.layer_norm.bias"] = (C,)
        dff = {256: 1024, 512: 1024, 1024: 2048, 1536: 3072}.get(C, 2 * C)
        s[x + "ff.layer_norm.weight"] = (C,)
        s[x + "ff.layer_norm.bias"] = (C,)
        s[x + "ff.layer_1.weight"] = (dff, C)
        s[x + "ff.layer_1.bias"] = (dff,)
        s[x + "ff.layer_2.weight"] = (C, dff)
        s[x + "ff.layer_2.bias"] = (C,)
    for l in range(cfg.n_levels):
        s[f"neck.fpn_norms.{l}.weight"] = (1, C, 1)
        s[f"neck.fpn_norms.{l}.bias"] = (1, C, 1)
        s[f"reg_head.scale.{l}.scale"] = ()
    for h in ("cls_head.", "reg_head."):
        for i in range(2):
            s[h + f"head.{i}.conv.weight"] = (C, C, 3)
            s[h + f"norm.{i}.weight"] = (1, C, 1)
            s[h + f"norm.{i}.bias"] = (1, C, 1)
    s["cls_head.cls_head.conv.weight"] = (K, C, 3)
    s["cls_head.cls_head.conv.bias"] = (K,)
    s["reg_head.offset_head.conv.weight"] = (2, C, 3)
    s["reg_head.offset_head.conv.bias"] = (2,)
    prefixes = ["pets."] + [f"pets_emas.{e}.module." for e in range(getattr(cfg, "n_emas", 0))]
    for j, blk in enumerate(cfg.adapt_blocks):
        T = cfg.max_seq_len >> blk
        for pre in prefixes:
            s[f"{pre}{j}.layer.0.weight"] = (5 * T, T)
            s[f"{pre}{j}.layer.0.bias"] = (5 * T,)
            s[f"{pre}{j}.layer.2.weight"] = (T // 2, 5 * T)
            s[f"{pre}{j}.layer.2.bias"] = (T // 2,)
    if getattr(cfg, "narration_dim", 0):      # narration SSL branch (meta_archs.py:650-652): Linear(narration_dim, 1024)
        s["narration_encoder.weight"] = (1024, cfg.narration_dim)
        s["narration_encoder.bias"] = (1024,)
    if getattr(cfg, "prompt_pool", None):
        pp = cfg.prompt_pool
        s["prompt.prompt"] = (pp["pool_size"], pp["length"], cfg.n_txt_in)
        s["prompt.prompt_key"] = (pp["pool_size"], cfg.n_txt_in)
    return s


def random_state(spec, seed=0):
    """Fill `spec` with seeded O(1) values (float32 torch CPU tensors)."""
    rs = np.random.RandomState(seed)
    out = {}
    for key in sorted(spec):
        shape = spec[key]
        n = rs.standard_normal(shape).astype(np.float32) if shape != () else np.float32(rs.standard_normal())
        last = key.split(".")[-1]
        if key in ("mu",):
            v = 0.1 * n
        elif key == "mu_reg_left":
            v = -0.5 + 0.1 * n
        elif key == "mu_reg_right":
            v = 0.5 + 0.1 * n
        elif key.startswith("sigma"):
            v = 1.0 + 0.2 * np.abs(n)
        elif last == "scale" and key.startswith("reg_head.scale"):
            v = 1.0 + 0.1 * n
        elif last == "scale":  # AffineDropPath
            v = 0.5 + 0.1 * n
        elif "norm" in key or ".ln" in key:
            v = (1.0 + 0.1 * n) if last == "weight" else 0.1 * n
        elif key.endswith("_conv.conv.weight"):  # depthwise
            v = 0.5 * n
        elif key == "cls_head.cls_head.conv.bias":
            v = -3.0 + 0.3 * n
        elif key == "reg_head.offset_head.conv.bias":
            v = 1.0 + 0.3 * n
        elif last == "bias" or key.endswith("_bias"):
            v = 0.1 * n
        elif key.startswith("prompt."):
            v = 0.5 * n
        elif "rel_attn." in key:  # (C,H,d) einsum weights: fan_in = C
            v = n / np.sqrt(shape[0])
        else:  # conv (out,in,k) / linear (out,in)
            fan_in = int(np.prod(shape[1:]))
            v = n / np.sqrt(fan_in)
        out[key] = torch.from_numpy(np.asarray(v, dtype=np.float32).reshape(shape)).clone()
    return out


def synth_video_list(cfg, n_videos, seed=0, lens=None, text_lens=None, n_gt=None):
    """Synthetic reference-style `video_list` (schema MQ/libs/datasets/ego4d.py:820-837)."""
    rs = np.random.RandomState(1000 + seed)
    T = cfg.max_seq_len
    out = []
    for i in range(n_videos):
        Ti = int(lens[i]) if lens is not None else T
        Li = int(text_lens[i]) if text_lens is not None else int(rs.randint(20, 129))
        ng = int(n_gt[i]) if n_gt is not None else int(rs.randint(1, 9))
        feats = rs.standard_normal((cfg.input_dim, Ti)).astype(np.float32)
        text = rs.standard_normal((cfg.n_txt_in, Li)).astype(np.float32)
        centre = rs.uniform(0, Ti, ng)
        length = np.exp(rs.uniform(np.log(4.0), np.log(min(512.0, Ti)), ng))
        s0 = np.clip(centre - length / 2, 0, Ti).astype(np.float32)
        s1 = np.clip(centre + length / 2, 0, Ti).astype(np.float32)
        s1 = np.maximum(s1, s0 + 1.0).astype(np.float32)
        labels = rs.randint(0, cfg.num_classes, ng).astype(np.int64)
        out.append({
            "video_id": f"synthetic_{seed}_{i}",
            "feats": torch.from_numpy(feats),
            "segments": torch.from_numpy(np.stack([s0, s1], 1)),
            "labels": torch.from_numpy(labels),
            "fps": 30.0, "duration": 480.0, "feat_stride": 480.0 * 30.0 / T, "feat_num_frames": 480.0 * 30.0 / T,
            "segmentation_labels": torch.zeros(Ti, cfg.num_classes),
            "prompt_feature": torch.from_numpy(text),
        })
    return out
