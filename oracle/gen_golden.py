"""TEST INFRASTRUCTURE ONLY — generate tests/golden/*.npz by running the REFERENCE ITSELF (imported from
/root/reference through oracle/ref_shim.py) on seeded inputs and seeded weights.

Run in the authoring container only:   python -m oracle.gen_golden
The GPU box never runs this (no /root/reference there); it only reads the committed vectors.
"""
import os
import sys

import numpy as np
import torch

from . import mq_oracle as O
from . import params as PR
from . import ref_shim

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def small_cfg():
    return O.ModelCfg(input_dim=192, embd_dim=256, n_head=4, max_seq_len=128, arch=(2, 2, 4), num_classes=6,
                      n_txt_in=96, regression_range=[[0, 4], [2, 8], [4, 16], [8, 32], [16, 10000]])


def vilco_cfg():
    """small mq_vilco.yaml-like config: L2P prompts + temporal adapters (sized for T = 1024 by the reference) + EMA."""
    return O.ModelCfg(input_dim=192, embd_dim=256, n_head=4, max_seq_len=1024, arch=(2, 2, 5), num_classes=6, n_txt_in=96,
                      regression_range=[[0, 4], [2, 8], [4, 16], [8, 32], [16, 64], [32, 10000]],
                      adapt_blocks=(0, 1, 2, 3, 4), n_emas=1, prompt_pool=dict(pool_size=10, top_k=4, length=20))


def vilco_train_cfg():
    """mq_vilco.yaml training at reduced depth: prompts + adapters + narration SSL need C = 1024 (narration_encoder output,
    meta_archs.py:650) and T = 1024 (adapter sizes, :687)."""
    return O.ModelCfg(input_dim=64, embd_dim=1024, n_head=16, max_seq_len=1024, arch=(2, 1, 5), num_classes=6, n_txt_in=96,
                      regression_range=[[0, 4], [2, 8], [4, 16], [8, 32], [16, 64], [32, 10000]],
                      adapt_blocks=(0, 1, 2, 3, 4), n_emas=1, prompt_pool=dict(pool_size=10, top_k=4, length=20),
                      narration_dim=32)


def _override(c):
    def ov(cfg):
        cfg["dataset"]["input_dim"] = [c.input_dim]
        cfg["dataset"]["num_classes"] = c.num_classes
        cfg["dataset"]["max_seq_len"] = c.max_seq_len
        m = cfg["model"]
        m["backbone_arch"] = list(c.arch)
        m["regression_range"] = c.regression_range
        m["n_head"] = c.n_head
        m["embd_dim"] = [c.embd_dim]
        m["fpn_dim"] = c.embd_dim
        m["head_dim"] = c.embd_dim
        m["n_txt_in"] = c.n_txt_in
        if c.prompt_pool:
            cfg["cl_cfg"].update(prompt_pool=True, pool_size=c.prompt_pool["pool_size"], topk=c.prompt_pool["top_k"],
                                 length=c.prompt_pool["length"], embed_dim=c.n_txt_in)
        if c.adapt_blocks:
            cfg["cl_cfg"].update(use_adapt=True, adapt_blocks=list(c.adapt_blocks))
        if getattr(c, "narration_dim", 0):
            cfg["cl_cfg"].update(narration_ssl=True, narration_dim=c.narration_dim, memory_size=48, ssl_factor=0.01)
            cfg["train_cfg"].update(dropout=0.0, droppath=1e-12)   # deterministic (keep_prob rounds to 1.0) but keeps AffineDropPath; the SSL branch needs model.train()
        else:
            cfg["cl_cfg"].update(narration_ssl=False)
    return ov


def build_reference_model(c, seed=0, yaml_name="mq_no_cl.yaml"):
    """Reference PtTransformer at config `c`, loaded with oracle.params.random_state(seed)."""
    torch.manual_seed(0)
    if yaml_name == "mq_vilco.yaml":
        torch.Tensor.cuda = lambda self, *a, **k: self   # MemoryBank / contrastive loss call .cuda() (meta_archs.py:42)
    model, cfg = ref_shim.build_model(_override(c), yaml_name)
    spec = PR.param_spec(c)
    sd = model.state_dict()
    for k, shp in spec.items():
        assert k in sd, f"spec key {k} missing from the reference state_dict"
        assert tuple(sd[k].shape) == tuple(shp), (k, tuple(sd[k].shape), shp)
    state = PR.random_state(spec, seed)
    missing, unexpected = model.load_state_dict(state, strict=False)
    assert not unexpected, unexpected
    model.eval()
    return model, state


def run_reference(c, model, videos):
    """Returns dict of numpy arrays: per-video logits/offsets (concatenated over levels), final detections,
    and the training losses of the whole batch (model.eval() so dropout/drop-path are identity)."""
    out = {}
    with torch.no_grad():
        for i, v in enumerate(videos):
            logits, offs, masks = model([v], is_training=False, get_emb=True)
            out[f"logits_{i}"] = torch.cat(logits, 1)[0].numpy()
            out[f"offsets_{i}"] = torch.cat(offs, 1)[0].numpy()
            out[f"masks_{i}"] = torch.cat(masks, 1)[0].numpy()
            res = model([v], is_training=False)[0]
            out[f"det_segments_{i}"] = res["segments"].numpy()
            out[f"det_scores_{i}"] = res["scores"].numpy()
            out[f"det_labels_{i}"] = res["labels"].numpy()
    model.loss_normalizer = c.init_loss_norm
    losses = model(videos, is_training=True)
    for k, val in losses.items():
        out["loss_" + k] = np.float32(val.detach().reshape(-1)[0].item())
    # gradients of final_loss w.r.t. a few probes (used later by the backward parity tests)
    return out


def full_cfg(num_classes=22):
    """mq_no_cl.yaml at full size (C=1024, T=1024, input 4096, 10 levels); num_classes=110 = after augment_classification
    over the 5 sub-tasks (train_cl.py:378)."""
    return O.ModelCfg(num_classes=num_classes)


FULL_VIDEOS = dict(seed=8, lens=[1024, 700], text_lens=[57, 33], n_gt=[4, 3])


def gen_full_golden():
    """The north-star configuration itself: the REFERENCE at mq_no_cl.yaml full size on 2 clips (T = 1024 and a ragged 700),
    K = 22 and K = 110 -> tests/golden/model_full.npz: per clip the logits / offsets / masks over all 2046 pyramid points, the
    final detections (decode + soft-NMS, 200 kept segments), and the training losses of the 2-clip batch."""
    out = {}
    for K in (22, 110):
        c = full_cfg(K)
        model, _ = build_reference_model(c, seed=4)
        videos = PR.synth_video_list(c, 2, **FULL_VIDEOS)
        r = run_reference(c, model, videos)
        for k, v in r.items():
            out[f"k{K}_{k}"] = v
        del model
    np.savez_compressed(os.path.join(GOLDEN, "model_full.npz"), **out)
    print("model_full:", {k: (v.shape if hasattr(v, "shape") and v.shape else float(v)) for k, v in out.items()})


def gen_model_golden():
    c = small_cfg()
    model, _ = build_reference_model(c, seed=0)
    videos = PR.synth_video_list(c, 2, seed=0, lens=[128, 100], text_lens=[40, 57], n_gt=[3, 2])
    out = run_reference(c, model, videos)
    np.savez_compressed(os.path.join(GOLDEN, "model_small.npz"), **out)
    print("model_small:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


def gen_grad_golden():
    """d final_loss / d parameter from the reference's own autograd (model.eval(): dropout / drop-path identity) on the
    model_small inputs.  Stored compactly per parameter: L2 norm, sum, and the first 8 entries of the flattened gradient."""
    c = small_cfg()
    model, _ = build_reference_model(c, seed=0)
    videos = PR.synth_video_list(c, 2, seed=0, lens=[128, 100], text_lens=[40, 57], n_gt=[3, 2])
    model.loss_normalizer = c.init_loss_norm
    model.zero_grad()
    losses = model(videos, is_training=True)
    losses["final_loss"].backward()
    out = {"final_loss": np.float32(losses["final_loss"].item())}
    spec = PR.param_spec(c)
    for k, p_ in model.named_parameters():
        if k in spec and p_.grad is not None:
            g = p_.grad.detach().reshape(-1).double()
            out["g:" + k] = np.concatenate([[g.norm().item(), g.sum().item()], g[:8].numpy()]).astype(np.float64)
    np.savez_compressed(os.path.join(GOLDEN, "grads_small.npz"), **out)
    print("grads_small:", len(out) - 1, "parameters")


def narration_inputs(c, videos, seed=9):
    """narration_feats (narration_dim, Ln) / narration_mask per video (schema ego4d.py:820-837), seeded."""
    rs = np.random.RandomState(seed)
    for i, v in enumerate(videos):
        ln = int(rs.randint(2, 9))
        v["narration_feats"] = torch.from_numpy(rs.standard_normal((c.narration_dim, ln)).astype(np.float32))
        v["narration_mask"] = 1.0 if i != 1 else 0.0        # one clip without narration
    return videos


def seeded_memory_bank(size, dim, seed=17):
    m = np.random.RandomState(seed).standard_normal((size, dim)).astype(np.float32)
    return torch.from_numpy(m / np.linalg.norm(m, axis=1, keepdims=True))


def gen_vilco_train_golden():
    """One training forward / backward of the mq_vilco branches in the REFERENCE (model.train(), every dropout p = 0): L2P
    prompts selected by task id + pull constraint (n_known > 0), temporal adapters, narration SSL with a seeded memory
    bank.  Stored: the losses and, per parameter, L2 norm / sum / first 8 entries of d final_loss / d parameter."""
    c = vilco_train_cfg()
    model, _ = build_reference_model(c, seed=2, yaml_name="mq_vilco.yaml")
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    model.train()
    model.n_known = 1
    model.memory_bank.memory = seeded_memory_bank(48, 1024)
    model.memory_bank.ptr = 0
    videos = narration_inputs(c, PR.synth_video_list(c, 3, seed=6, lens=[1024, 900, 700], text_lens=[40, 57, 33], n_gt=[3, 2, 4]))
    model.loss_normalizer = c.init_loss_norm
    model.zero_grad()
    losses = model(videos, task_id=1, is_training=True)
    losses["final_loss"].backward()
    out = {"loss_" + k: np.float32(v.detach().reshape(-1)[0].item()) for k, v in losses.items()}
    n = 0
    for k, p_ in model.named_parameters():
        if p_.grad is not None:      # incl. the temporal adapters, listed under backbone.branch.<b>.adapters.attn.*
            g = p_.grad.detach().reshape(-1).double()
            out["g:" + k] = np.concatenate([[g.norm().item(), g.sum().item()], g[:8].numpy()]).astype(np.float64)
            n += 1
    out["memory_after"] = model.memory_bank.memory[:4].numpy()
    np.savez_compressed(os.path.join(GOLDEN, "train_vilco.npz"), **out)
    print("train_vilco:", {k: float(v) for k, v in out.items() if k.startswith("loss_")}, n, "parameter gradients")


def prev_logits(c, seed=4, probs=False):
    """per-level (T_l, K) arrays standing in for the previous task's recorded classification outputs"""
    rs = np.random.RandomState(seed)
    out = []
    for l in range(c.n_levels):
        a = rs.standard_normal((c.max_seq_len >> l, c.num_classes)).astype(np.float32)
        out.append(1.0 / (1.0 + np.exp(-a)) if probs else a)
    return out


def gen_distill_golden():
    """BiC (bias layers on the class slices + soft-target distillation) and iCaRL (BCE distillation) terms of the training
    loss (meta_archs.py:823-836, 1482-1519) from the REFERENCE: losses and gradient summaries, model.eval() (no dropout)."""
    c = small_cfg()
    out = {}
    for name, yaml_name in (("bic", "mq_bic_all.yaml"), ("icarl", "mq_icarl_all.yaml")):
        model, _ = build_reference_model(c, seed=0, yaml_name=yaml_name)
        ns = ref_shim.load()
        model.n_known = 3
        if name == "bic":
            model.list_splits = [3, 6]
            model.list_bias_layers = [ns.meta_archs.BiasLayer(), ns.meta_archs.BiasLayer()]
            for bl, (a, b) in zip(model.list_bias_layers, ((1.1, 0.05), (0.9, -0.02))):
                bl.alpha.data.fill_(a)
                bl.beta.data.fill_(b)
        videos = PR.synth_video_list(c, 2, seed=0, lens=[128, 100], text_lens=[40, 57], n_gt=[3, 2])
        prev = prev_logits(c, probs=True)
        model.loss_normalizer = c.init_loss_norm
        model.zero_grad()
        losses = model(videos, is_training=True, prev_out_cls_logits=prev if name == "bic" else [prev])
        losses["final_loss"].backward()
        for k, v in losses.items():
            out[f"{name}_loss_{k}"] = np.float32(torch.as_tensor(v).detach().reshape(-1)[0].item())
        spec = PR.param_spec(c)
        for k, p_ in model.named_parameters():
            if k in spec and p_.grad is not None:
                g = p_.grad.detach().reshape(-1).double()
                out[f"{name}_g:{k}"] = np.concatenate([[g.norm().item(), g.sum().item()], g[:8].numpy()]).astype(np.float64)
        if name == "bic":
            out["bic_bias_grads"] = np.array([[bl.alpha.grad.item(), bl.beta.grad.item()] for bl in model.list_bias_layers])
    np.savez_compressed(os.path.join(GOLDEN, "distill_small.npz"), **out)
    print("distill_small:", {k: float(v) for k, v in out.items() if "_loss_" in k})


def gen_nms_voting_golden():
    """class-agnostic batched_nms with segment voting (nms.py:159-181) from the reference python wrapper + C++ extension."""
    ns = ref_shim.load()
    rs = np.random.RandomState(11)
    out = {}
    for name, n in (("a", 500), ("b", 4000)):
        centre = rs.uniform(0, 1024, n).astype(np.float32)
        length = np.exp(rs.uniform(np.log(2.0), np.log(400.0), n)).astype(np.float32)
        segs = np.stack([centre - length / 2, centre + length / 2], 1).astype(np.float32)
        scores = rs.beta(0.5, 8, n).astype(np.float32) + np.float32(1e-3)
        labels = rs.randint(0, 5, n).astype(np.int64)
        for soft in (True, False):
            s, sc, lb = ns.nms.batched_nms(torch.from_numpy(segs), torch.from_numpy(scores), torch.from_numpy(labels), 0.3,
                                           1e-3, 100, use_soft_nms=soft, multiclass=False, sigma=0.75, voting_thresh=0.75)
            tag = f"{name}_{'soft' if soft else 'hard'}"
            out[tag + "_segs"], out[tag + "_scores"], out[tag + "_labels"] = s.numpy(), sc.numpy(), lb.numpy()
        out[name + "_in_segs"], out[name + "_in_scores"], out[name + "_in_labels"] = segs, scores, labels
    np.savez_compressed(os.path.join(GOLDEN, "nms_voting.npz"), **out)
    print("nms_voting ok", {k: v.shape for k, v in out.items() if k.endswith("_segs") and "_in_" not in k})


def gen_optimizer_groups_golden():
    """parameter grouping of the reference's make_optimizer (train_utils.py:68-143) on the small and the vilco model:
    names per group (decay / no_decay / remain) — pins trainer.make_optimizer."""
    import json
    ref_shim.load()
    import libs.utils.train_utils as TU
    out = {}
    for tag, c, yaml_name in (("small", small_cfg(), "mq_no_cl.yaml"), ("vilco", vilco_cfg(), "mq_vilco.yaml")):
        model, _ = build_reference_model(c, seed=0 if tag == "small" else 1, yaml_name=yaml_name)
        opt = TU.make_optimizer(model, {"type": "AdamW", "learning_rate": 1e-4, "weight_decay": 0.05, "momentum": 0.9})
        names = {id(p): k for k, p in model.named_parameters()}
        out[tag] = [{"weight_decay": g["weight_decay"], "params": sorted(names[id(p)] for p in g["params"])}
                    for g in opt.param_groups]
    with open(os.path.join(GOLDEN, "optimizer_groups.json"), "w") as f:
        json.dump(out, f)
    print("optimizer_groups:", {k: [len(g["params"]) for g in v] for k, v in out.items()})


def gen_vilco_golden():
    """mq_vilco.yaml branches at inference: prompts prepended to the text, adapters on branch 0-4, EMA-adapter ensemble."""
    c = vilco_cfg()
    model, _ = build_reference_model(c, seed=1, yaml_name="mq_vilco.yaml")
    videos = PR.synth_video_list(c, 1, seed=5, lens=[900], text_lens=[57], n_gt=[3])
    out = {}
    with torch.no_grad():
        logits, offs, masks = model(videos, is_training=False, get_emb=True)
        out["logits_0"] = torch.cat(logits, 1)[0].numpy()
        out["offsets_0"] = torch.cat(offs, 1)[0].numpy()
        # quirk: the EMA-ensemble loop re-binds fpn_masks to the un-squeezed (B,1,T_l) tensors (meta_archs.py:864)
        out["mask_shape_l0"] = np.asarray(masks[0].shape)
        out["masks_0"] = torch.cat([m.reshape(m.shape[0], -1) for m in masks], 1)[0].numpy()
        res = model(videos, is_training=False)[0]
        out["det_segments_0"], out["det_scores_0"], out["det_labels_0"] = (res[k].numpy() for k in ("segments", "scores", "labels"))
    out["state_keys"] = np.array(list(model.state_dict().keys()))
    np.savez_compressed(os.path.join(GOLDEN, "model_vilco.npz"), **out)
    print("model_vilco:", {k: v.shape for k, v in out.items()})


def gen_local_attn_golden():
    """LocalMaskedMHCA standalone (unused by MQ configs, live in NLQ) — MQ/libs/modeling/blocks.py:871-1207."""
    ns = ref_shim.load()
    out = {}
    for name, (C, H, W, T, valid) in {"w9": (128, 2, 9, 64, [64, 41]), "w5": (128, 2, 5, 32, [32, 20])}.items():
        torch.manual_seed(1)
        m = ns.blocks.LocalMaskedMHCA(C, H, W).eval()
        rs = np.random.RandomState(7)
        sd = {k: torch.from_numpy(rs.standard_normal(tuple(v.shape)).astype(np.float32) * (0.3 if v.dim() == 3 else 0.1) + (1.0 if "norm.weight" in k else 0.0))
              for k, v in sorted(m.state_dict().items())}
        m.load_state_dict(sd)
        x = torch.from_numpy(rs.standard_normal((2, C, T)).astype(np.float32))
        mask = (torch.arange(T)[None, :] < torch.tensor(valid)[:, None]).unsqueeze(1)
        with torch.no_grad():
            y, _ = m(x * mask, mask)
        out[name + "_x"] = (x * mask).numpy()
        out[name + "_valid"] = np.asarray(valid)
        out[name + "_y"] = y.numpy()
        for k, v in sd.items():
            out[name + "_p_" + k] = v.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "local_attn.npz"), **out)
    print("local_attn ok")


def gen_nms_golden():
    """soft-NMS / hard-NMS known answers from the reference's own C++ (MQ/libs/utils/csrc/nms_cpu.cpp)."""
    ns = ref_shim.load()
    out = {}
    rs = np.random.RandomState(3)
    cases = {"n1": 1, "n17": 17, "n300": 300, "n2000": 2000, "ties": 64}
    for name, n in cases.items():
        centre = rs.uniform(0, 1024, n).astype(np.float32)
        length = np.exp(rs.uniform(np.log(2.0), np.log(400.0), n)).astype(np.float32)
        segs = np.stack([centre - length / 2, centre + length / 2], 1).astype(np.float32)
        scores = rs.beta(0.5, 8, n).astype(np.float32)
        if name == "ties":
            scores = np.round(scores * 8) / 8 + np.float32(0.01)  # many exactly equal scores
            scores = scores.astype(np.float32)
        out[name + "_segs"], out[name + "_scores"] = segs, scores
        for method, sigma, min_score in ((2, 0.99, 1e-4), (2, 0.5, 0.01), (1, 0.5, 0.001), (0, 0.5, 0.001)):
            dets = torch.zeros(n, 3)
            inds = ns.nms_1d_cpu.softnms(torch.from_numpy(segs), torch.from_numpy(scores), dets, iou_threshold=0.1,
                                         sigma=float(sigma), min_score=float(min_score), method=int(method))
            tag = f"{name}_m{method}_s{sigma}_t{min_score}"
            out[tag + "_dets"] = dets[:len(inds)].numpy()
            out[tag + "_inds"] = inds.numpy()
        keep = ns.nms_1d_cpu.nms(torch.from_numpy(segs), torch.from_numpy(scores), iou_threshold=0.4)
        out[name + "_hard_keep"] = keep.numpy()
    # batched_nms end-to-end (python wrapper + extension), multi-class
    n, K = 3000, 7
    centre = rs.uniform(0, 1024, n).astype(np.float32)
    length = np.exp(rs.uniform(np.log(2.0), np.log(400.0), n)).astype(np.float32)
    segs = np.stack([centre - length / 2, centre + length / 2], 1).astype(np.float32)
    scores = rs.beta(0.5, 8, n).astype(np.float32)
    labels = rs.randint(0, K, n).astype(np.int64)
    s, sc, lb = ns.nms.batched_nms(torch.from_numpy(segs), torch.from_numpy(scores), torch.from_numpy(labels), 0.1,
                                   1e-4, 200, use_soft_nms=True, multiclass=True, sigma=0.99, voting_thresh=0.9)
    out.update(b_segs=segs, b_scores=scores, b_labels=labels, b_out_segs=s.numpy(), b_out_scores=sc.numpy(),
               b_out_labels=lb.numpy())
    np.savez_compressed(os.path.join(GOLDEN, "nms.npz"), **out)
    print("nms ok")


if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    what = sys.argv[1:] or ["nms", "local", "model"]
    if "nms" in what:
        gen_nms_golden()
    if "local" in what:
        gen_local_attn_golden()
    if "model" in what:
        gen_model_golden()
    if "vilco" in what:
        gen_vilco_golden()
    if "grads" in what:
        gen_grad_golden()
    if "vilco_train" in what:
        gen_vilco_train_golden()
    if "distill" in what:
        gen_distill_golden()
    if "voting" in what:
        gen_nms_voting_golden()
    if "optimizer" in what:
        gen_optimizer_groups_golden()
    if "full" in what:
        gen_full_golden()
