"""GPU probe: gradients of final_loss from the CUDA training path vs torch autograd through the oracle (CPU).
usage: python tools/train_probe.py [small|vilco] [B]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import gen_golden as GG  # noqa: E402
from oracle import mq_oracle as O  # noqa: E402
from oracle import params as PR  # noqa: E402
import util  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    cfg = GG.small_cfg()
    model, P = util.build_pair(cfg, 0)
    videos = PR.synth_video_list(cfg, 2, seed=0, lens=[128, 100], text_lens=[40, 57], n_gt=[3, 2])
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    t0 = time.time()
    x, mask, text, tmask = O.preprocess(cfg, videos, True)
    logits, offs, masks, _ = O.forward_heads(Pg, cfg, x, mask, text, tmask, training=True)
    for t in list(logits) + list(offs):
        t.retain_grad()
    lo, _ = O.losses(Pg, cfg, masks, logits, offs, [v["segments"] for v in videos], [v["labels"] for v in videos],
                     cfg.init_loss_norm)
    lo["final_loss"].backward()
    print("oracle losses", {k: float(v) for k, v in lo.items()}, f"{time.time() - t0:.1f}s")
    model.eval()   # dropout / drop-path off, as in the golden generation
    model.loss_normalizer = cfg.init_loss_norm
    out = model(videos, is_training=True)
    print("cuda losses  ", {k: float(v) for k, v in out.items()})
    import vilco_b200.train_engine as TE
    keep = {}
    orig = TE.Tape.backward

    def bw(self):
        keep["tape"] = self
        orig(self)
    TE.Tape.backward = bw
    out["final_loss"].backward()
    torch.cuda.synchronize()
    dl = torch.cat([t.grad for t in logits], 1)
    do = torch.cat([t.grad for t in offs], 1)
    print("d logits max", float(dl.abs().max()), "d offsets max", float(do.abs().max()))
    mdl, mdo, pyr = model._last_head_grads
    mdl = torch.cat([mdl[:, o:o + n] for o, n in zip(pyr.off, pyr.lens)], 1).cpu()
    mdo = torch.cat([mdo[:, o:o + n] for o, n in zip(pyr.off, pyr.lens)], 1).cpu()
    print("dlogits rel err", util.rel_max(mdl, dl), "doffsets rel err", util.rel_max(mdo, do))
    named = dict(model.named_parameters())
    rows = []
    for k, g in Pg.items():
        if g.grad is None:
            continue
        p = named.get(k)
        if p is None:
            continue
        if p.grad is None:
            rows.append((float("inf"), k, float(g.grad.abs().max()), 0.0))
            continue
        a, b = p.grad.detach().cpu().double(), g.grad.double()
        err = float((a - b).abs().max() / (b.abs().max() + 1e-20))
        nbad = int(((a - b).abs() > 1e-3 * b.abs().max()).sum())
        rows.append((err, k + f"  nbad {nbad}/{b.numel()}", float(b.abs().max()), float(a.abs().max())))
    rows.sort(reverse=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/train_probe_all.txt", "w") as f:
        for err, k, mb, ma in rows:
            f.write(f"{err:10.3e}  ref_max {mb:9.3e}  got_max {ma:9.3e}  {k}\n")
    for err, k, mb, ma in rows[:12]:
        print(f"{err:10.3e}  ref_max {mb:9.3e}  got_max {ma:9.3e}  {k}")
    print("n params compared", len(rows), "worst", rows[0][0], "median", rows[len(rows) // 2][0])
    extra = [k for k, p in named.items() if p.grad is not None and (k not in Pg or Pg[k].grad is None)]
    print("grads only on the cuda side:", extra[:10])


if __name__ == "__main__":
    main()
