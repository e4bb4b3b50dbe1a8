"""GPU probe: engine forward (backbone + neck + heads) vs the CPU oracle on the same seeded weights/inputs.
usage: python tools/model_probe.py [small|full] [B]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mq_oracle as O  # noqa: E402
from oracle import params as PR  # noqa: E402
from oracle.gen_golden import small_cfg  # noqa: E402
from vilco_b200 import engine as E  # noqa: E402
from vilco_b200 import ops  # noqa: E402


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item(), ((a - b).norm() / (b.norm() + 1e-12)).item()


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    cfg = small_cfg() if which == "small" else O.ModelCfg()
    if len(sys.argv) > 3:
        ops.set_precision(sys.argv[3])
    print("precision", ops.precision())
    T = cfg.max_seq_len
    lens = [T, T * 3 // 4 + 1][:B] + [T] * max(0, B - 2)
    tl = [40, 57][:B] + [33] * max(0, B - 2)
    vids = PR.synth_video_list(cfg, B, seed=0, lens=lens, text_lens=tl)
    P = PR.random_state(PR.param_spec(cfg), 0)
    dev = "cuda"
    W = E.pack_weights(P, dev)
    x, mask, text, tmask = O.preprocess(cfg, vids, False)
    t0 = time.time()
    with torch.no_grad():
        feats_o, masks_o = O.backbone(P, cfg, x, mask, text, tmask, training=False)
        fpn_o, _ = O.fpn_identity(P, feats_o, masks_o)
        offs_o = O.reg_head(P, fpn_o, masks_o)
        logits_o = O.cls_head(P, fpn_o, masks_o)
    print(f"oracle cpu forward {time.time() - t0:.2f}s", flush=True)

    x16 = ops.pack_feats(x.to(dev))
    t16 = ops.pack_feats(text.to(dev))
    m = mask.squeeze(1).float().to(dev).contiguous()
    tm = tmask.squeeze(1).float().to(dev).contiguous()
    pe = E.sinusoid_pe_table(cfg.max_seq_len, cfg.embd_dim, dev)
    feats, masks = E.backbone_fwd(W, cfg, x16, m, t16, tm, pe)
    logits, offs, pmask, pyr = E.neck_heads_fwd(W, cfg, feats, masks)
    torch.cuda.synchronize()
    for l, (f, fo) in enumerate(zip(feats, feats_o)):
        mk = masks_o[l].transpose(1, 2).float()
        a = f.cpu() * mk
        b = fo.transpose(1, 2) * mk
        print(f"level {l} feat   max-rel {rel(a, b)[0]:.3e}  l2-rel {rel(a, b)[1]:.3e}  finite {torch.isfinite(f).all().item()}", flush=True)
    for l in range(len(feats)):
        o, n = pyr.off[l], pyr.lens[l]
        a = logits[:, o:o + n].cpu()
        b = logits_o[l].transpose(1, 2)
        c = offs[:, o:o + n].cpu()
        d = offs_o[l].transpose(1, 2)
        print(f"level {l} logits max-rel {rel(a, b)[0]:.3e} l2-rel {rel(a, b)[1]:.3e} | offsets max-rel {rel(c, d)[0]:.3e} l2-rel {rel(c, d)[1]:.3e}", flush=True)
    # timing of the GPU forward (eager, python launch overhead included)
    for _ in range(2):
        feats, masks = E.backbone_fwd(W, cfg, x16, m, t16, tm, pe)
        E.neck_heads_fwd(W, cfg, feats, masks, pyr)
    torch.cuda.synchronize()
    from vilco_b200 import lib as L
    n0 = L.launch_count()
    t0 = time.time()
    for _ in range(5):
        feats, masks = E.backbone_fwd(W, cfg, x16, m, t16, tm, pe)
        E.neck_heads_fwd(W, cfg, feats, masks, pyr)
    torch.cuda.synchronize()
    print(f"gpu eager forward {(time.time() - t0) / 5 * 1e3:.2f} ms per batch of {B}; launches/forward {(L.launch_count() - n0) // 5}")


if __name__ == "__main__":
    main()
