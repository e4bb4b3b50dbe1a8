"""Training path on the B200: hand-written backward (vilco_b200/train_engine.py) vs torch autograd through the oracle.

Three layers of evidence:
  * block level (TransformerBlock in all its variants, XLNetLayer): gradients w.r.t. inputs and every parameter within
    2e-4 of CPU autograd through the oracle block.  These blocks contain no ReLU and the max-pool acts on the given input,
    so both sides take identical discrete decisions and the comparison is exact up to rounding.
  * whole model: d final_loss / d parameter vs the oracle's autograd (itself pinned to the reference's gradients,
    tests/golden/grads_small.npz).  The model has ~0.7 M ReLU gates; a pre-activation within ~1e-5 of zero may take the
    other side on the GPU (different summation order), which moves that element's whole gradient.  Such a flip perturbs
    everything below it by O(1e-2), so the model-level bound is a relative L2 bound per parameter, with the strict bound
    kept for the parameters no gate sits above (final head convolutions, mu / sigma).
  * a central finite difference of the CUDA forward loss along the computed gradient direction (validates the backward
    against the parity-checked forward, independent of gate flips).
"""
import math

import numpy as np
import pytest
import torch

from oracle import gen_golden as GG
from oracle import mq_oracle as O
from oracle import params as PR
from util import build_pair, rel_max

pytestmark = pytest.mark.gpu
PRECISION = "fp16x3"   # the tolerances below state the exact (split-operand) arithmetic; see conftest._precision_mode


def _grads_block(pre, stride, cross, adapter=False, T=64, L=24, seed=0):
    from vilco_b200 import engine as E
    from vilco_b200 import train_engine as TE
    cfg = GG.small_cfg()
    C, H, B = cfg.embd_dim, cfg.n_head, 2
    spec = {k: v for k, v in PR.param_spec(cfg).items() if k.startswith(pre)}
    if adapter:
        spec.update({"pets.0.layer.0.weight": (5 * T, T), "pets.0.layer.0.bias": (5 * T,),
                     "pets.0.layer.2.weight": (T // 2, 5 * T), "pets.0.layer.2.bias": (T // 2,)})
    P = PR.random_state(spec, seed)
    g = torch.Generator().manual_seed(seed)
    valid, tvalid = [T, T - 13], [L, L - 7]
    mask = torch.arange(T)[None, :] < torch.tensor(valid)[:, None]
    tmask = torch.arange(L)[None, :] < torch.tensor(tvalid)[:, None]
    x = torch.randn(B, C, T, generator=g) * mask[:, None]
    y = torch.randn(B, C, L, generator=g) * tmask[:, None]
    # oracle autograd (CPU fp32)
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    xg, yg = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    out, om = O.transformer_block(Pg, pre, xg, mask[:, None], H, stride, cross_y=yg if cross else None,
                                  cross_y_mask=tmask.long() if cross else None, t_c_alpha=0.8,
                                  adapter_pre="pets.0." if adapter else None)
    R = torch.randn(out.shape, generator=g)
    (out * R).sum().backward()
    # CUDA tape
    W = E.pack_weights(P, "cuda")
    tp = TE.Tape(W)
    xv = TE.V(x.permute(0, 2, 1).contiguous().cuda())
    yv = TE.V(y.permute(0, 2, 1).contiguous().cuda())
    o, _ = TE.transformer_block(tp, pre, xv, mask.float().cuda(), H, stride, cross=(yv, tmask.float().cuda()) if cross else None,
                                t_c_alpha=0.8, adapter_pre="pets.0." if adapter else None)
    assert rel_max(o.v.permute(0, 2, 1), out.detach()) < 1e-4
    o.g = R.permute(0, 2, 1).contiguous().cuda()
    tp.backward()
    errs = {"dx": rel_max(xv.g.permute(0, 2, 1), xg.grad)}
    if cross:
        errs["dtext"] = rel_max(yv.g.permute(0, 2, 1), yg.grad)
    scale = max(float(v.grad.abs().max()) for v in Pg.values() if v.grad is not None)
    for k, v in Pg.items():
        if v.grad is None:
            continue
        assert k in tp.G, f"no gradient for {k}"
        got = TE.unpack_grad(k, tp.G[k], v).cpu()
        if float(v.grad.abs().max()) < 1e-6 * scale:      # analytically zero (key bias under softmax): absolute check
            assert float(got.abs().max()) < 1e-5 * scale, k
            continue
        errs[k] = rel_max(got, v.grad)
    return errs


@pytest.mark.parametrize("pre,stride,cross,adapter", [
    ("backbone.stem.0.", 1, False, False),       # + ChannelBlock mix
    ("backbone.branch.1.", 2, False, False),     # strided, max-pool skip
    ("backbone.branch.0.", 2, True, False),      # + cross attention to the text
    ("backbone.branch.1.", 2, False, True),      # + temporal adapter
])
def test_transformer_block_gradients(pre, stride, cross, adapter):
    errs = _grads_block(pre, stride, cross, adapter)
    bad = {k: e for k, e in errs.items() if not e < 2e-4}
    assert not bad, bad


def test_xlnet_layer_gradients():
    from vilco_b200 import engine as E
    from vilco_b200 import train_engine as TE
    cfg = GG.small_cfg()
    C, H, B, T = cfg.embd_dim, cfg.n_head, 2, 128
    pre = "backbone.xlnet.layer.0."
    P = PR.random_state({k: v for k, v in PR.param_spec(cfg).items() if k.startswith(pre)}, 3)
    g = torch.Generator().manual_seed(3)
    mask = (torch.arange(T)[None, :] < torch.tensor([T, 90])[:, None]).float()
    x = torch.randn(B, T, C, generator=g)
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    xg = x.clone().requires_grad_(True)
    out = O.xlnet_layer(Pg, pre, xg, mask)
    R = torch.randn(out.shape, generator=g)
    (out * R).sum().backward()
    tp = TE.Tape(E.pack_weights(P, "cuda"))
    xv = TE.V(x.cuda())
    o = TE.xlnet_layer(tp, pre, xv, mask.cuda(), H)
    assert rel_max(o.v, out.detach()) < 1e-4
    o.g = R.cuda()
    tp.backward()
    errs = {"dx": rel_max(xv.g, xg.grad)}
    for k, v in Pg.items():
        if v.grad is not None:
            errs[k] = rel_max(TE.unpack_grad(k, tp.G[k], v).cpu(), v.grad)
    bad = {k: e for k, e in errs.items() if not e < 2e-4}
    assert not bad, bad


def _model_grads():
    cfg = GG.small_cfg()
    model, P = build_pair(cfg, 0)
    videos = PR.synth_video_list(cfg, 2, seed=0, lens=[128, 100], text_lens=[40, 57], n_gt=[3, 2])
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    lo, _ = O.model_train_losses(Pg, cfg, videos)
    lo["final_loss"].backward()
    model.eval()   # dropout / drop-path off, as in the golden generation
    model.loss_normalizer = cfg.init_loss_norm
    out = model(videos, is_training=True)
    out["final_loss"].backward()
    torch.cuda.synchronize()
    return cfg, model, videos, Pg, lo, out


def test_model_gradients_vs_oracle_autograd():
    cfg, model, videos, Pg, lo, out = _model_grads()
    for k in ("cls_loss", "reg_loss", "al_loss", "final_loss"):
        assert abs(float(out[k]) - float(lo[k])) <= 1e-3 * abs(float(lo[k])) + 1e-6, k
    named = dict(model.named_parameters())
    gmax = max(float(v.grad.abs().max()) for v in Pg.values())
    l2 = {}
    for k, v in Pg.items():
        p = named[k]
        assert p.grad is not None, f"{k} received no gradient"
        a, b = p.grad.detach().cpu().double(), v.grad.double()
        if float(b.abs().max()) < 1e-6 * gmax:
            assert float(a.abs().max()) < 1e-5 * gmax, k
            continue
        l2[k] = float((a - b).norm() / b.norm())
    # no ReLU gate above these: strict
    for k in ("cls_head.cls_head.conv.weight", "cls_head.cls_head.conv.bias", "reg_head.offset_head.conv.weight",
              "reg_head.offset_head.conv.bias", "mu", "sigma", "mu_reg_left", "sigma_reg_left", "mu_reg_right", "sigma_reg_right"):
        assert l2[k] < 1e-4, (k, l2[k])
    worst = max(l2.items(), key=lambda kv: kv[1])
    assert worst[1] < 1e-1, worst
    assert float(np.median(list(l2.values()))) < 2e-2


def test_model_gradient_matches_finite_difference_of_cuda_forward():
    cfg, model, videos, Pg, lo, out = _model_grads()
    params = [p for p in model.parameters() if p.grad is not None]

    def loss_at(step, direction):
        with torch.no_grad():
            for p, d in zip(params, direction):
                p.add_(d, alpha=step)
            model.loss_normalizer = cfg.init_loss_norm
            v = float(model(videos, is_training=True)["final_loss"])
            for p, d in zip(params, direction):
                p.sub_(d, alpha=step)
        return v

    for select in (lambda n: True, lambda n: n.startswith("backbone.")):
        names = [n for n, p in model.named_parameters() if p.grad is not None]
        direction = [p.grad.clone() if select(n) else torch.zeros_like(p.grad) for n, p in zip(names, params)]
        gnorm = math.sqrt(sum(float((d.double() ** 2).sum()) for d in direction))
        direction = [d / gnorm for d in direction]
        fds = {h: (loss_at(h, direction) - loss_at(-h, direction)) / (2 * h) for h in (4e-3, 2e-3, 1e-3)}
        print("finite differences", fds, "analytic", gnorm)
        assert abs(fds[2e-3] - gnorm) <= 2e-2 * gnorm, (fds, gnorm)
        extrap = 2 * fds[1e-3] - fds[2e-3]         # the O(h) term comes from the gates that switch inside [-h, h]
        assert abs(extrap - gnorm) <= 5e-3 * gnorm, (fds, extrap, gnorm)


def test_flat_adamw_matches_torch_adamw_and_keeps_planes_current():
    """FlatAdamW (csrc/optim.cu) == clip_grad_norm_ + torch.optim.AdamW on identical gradients; its bf16 planes always equal
    the hi / lo split of the updated parameters."""
    from vilco_b200 import ops
    from vilco_b200.trainer import FlatAdamW
    torch.manual_seed(0)
    shapes = [(64, 40, 1), (40,), (3, 17), (1, 24, 1), ()]
    ref = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    mine = [torch.nn.Parameter(p.detach().clone()) for p in ref]
    groups = lambda ps: [{"params": ps[:3], "weight_decay": 0.05}, {"params": ps[3:], "weight_decay": 0.0}]  # noqa: E731
    o_ref = torch.optim.AdamW(groups(ref), lr=1e-2)
    o_mine = FlatAdamW(groups(mine), lr=1e-2)
    for it in range(4):
        gs = [torch.randn(s, device="cuda") * (3.0 if it % 2 else 0.01) for s in shapes]   # clipped and un-clipped steps
        o_mine.zero_grad()
        for p, q, g in zip(ref, mine, gs):
            p.grad = g.clone()
            q.grad.add_(g)
        nrm = torch.nn.utils.clip_grad_norm_(ref, 1.0)
        o_ref.step()
        o_mine.step(clip_grad_l2norm=1.0)
        assert abs(float(o_mine.grad_norm()) - float(nrm)) <= 1e-5 * float(nrm)
        for p, q in zip(ref, mine):
            assert rel_max(q.detach(), p.detach()) < 2e-6
    for q in mine:
        pv = o_mine.plane_view(q, (q.numel(),))
        assert torch.equal(pv, ops.split16(q.detach().reshape(-1)))


def test_training_steps_flat_optimizer_equals_torch_optimizer():
    """Three Trainer.step iterations of the small model: FlatAdamW path (live plane views, no re-pack) vs torch AdamW path
    (full re-pack each step) end with the same parameters and losses."""
    from vilco_b200.trainer import Trainer, make_optimizer
    cfg = GG.small_cfg()
    videos = PR.synth_video_list(cfg, 2, seed=0, lens=[128, 100], text_lens=[40, 57], n_gt=[3, 2])
    out = []
    for flat in (False, True):
        model, P = build_pair(cfg, 0)
        model.eval()                      # deterministic: no dropout
        model.loss_normalizer = cfg.init_loss_norm
        model.loss_normalizer_momentum = 1.0     # frozen normaliser, so that the loss values of successive steps compare
        opt = make_optimizer(model, {"type": "AdamW", "learning_rate": 1e-3, "weight_decay": 0.05}, flat=flat)
        tr = Trainer(model, opt, clip_grad_l2norm=1.0)
        losses = [float(tr.step(videos)["final_loss"]) for _ in range(3)]
        out.append((losses, {k: v.detach().clone() for k, v in model.named_parameters()}))
    (l0, p0), (l1, p1) = out
    assert l0[0] > l0[2], "loss should go down on a repeated batch"
    for a, b in zip(l0, l1):
        assert abs(a - b) <= 1e-4 * abs(a), (l0, l1)
    # Adam normalises every element's gradient, so elements whose gradient is rounding noise (key biases are analytically
    # zero, a few weights are numerically ~0) move by +-lr in a direction the last bit decides; compare the UPDATES in L2
    # (parameters outside the seeded spec are dead weights with a random init per construction)
    errs = []
    for k in p0:
        if k not in P or k.endswith("key.bias") or k.endswith("key_norm.bias"):
            continue
        init = P[k].cuda().reshape(p0[k].shape)
        d0, d1 = (p0[k] - init).double(), (p1[k] - init).double()
        errs.append((float((d1 - d0).norm() / (d0.norm() + 1e-12)), k))
    errs.sort(reverse=True)
    assert errs[0][0] < 5e-2, errs[:8]


def test_dropout_mask_statistics_and_fused_residual_branch():
    """vilco_dropout: inverted dropout with a reproducible counter-based mask; vilco_resid_branch_fwd / _bwd (the fused
    `resid*mask + scale * dropout(proj + b) * mask * drop_path` of a TransformerBlock in training mode) against the same
    expression in torch using the mask that vilco_dropout produces for the same seed."""
    import ctypes as C
    from vilco_b200 import lib as L
    from vilco_b200 import ops
    torch.manual_seed(0)
    R, Cc, p, seed = 777, 256, 0.1, 12345
    ones = torch.ones(R, Cc, device="cuda")
    keep = ops.dropout(ones, p, seed)                       # 0 or 1 / (1 - p)
    vals = torch.unique(keep)
    assert vals.numel() == 2 and vals[0] == 0 and abs(float(vals[1]) - 1 / (1 - p)) < 1e-6
    frac = float((keep == 0).float().mean())
    assert abs(frac - p) < 4 * math.sqrt(p * (1 - p) / (R * Cc))
    assert torch.equal(keep, ops.dropout(ones, p, seed))    # same seed -> same mask (this is how the backward re-derives it)
    assert not torch.equal(keep, ops.dropout(ones, p, seed + 1))
    x = torch.randn(R, Cc, device="cuda")
    y32, y16 = ops.dropout(x, p, seed, out16=True)
    assert torch.equal(y32, x * keep) and rel_max(ops.merge16(y16), y32) < 2e-5

    resid, y, g = (torch.randn(R, Cc, device="cuda") for _ in range(3))
    rm = (torch.rand(R, device="cuda") > 0.2).float()
    ymul = rm * torch.tensor([0.0, 1 / 0.9], device="cuda")[torch.randint(0, 2, (R,), device="cuda")]
    bias, scale = torch.randn(Cc, device="cuda"), torch.randn(Cc, device="cuda")
    out = torch.empty_like(y)
    st = L.stream_ptr()
    L.check(L.lib().vilco_resid_branch_fwd(ops._p(resid), ops._p(rm), ops._p(y), ops._p(bias), ops._p(scale), ops._p(ymul),
                                           ops._p(out), ops._i64(R), Cc, C.c_float(p), C.c_uint64(seed), st))
    ref = resid * rm[:, None] + scale * ((y + bias) * keep) * ymul[:, None]
    assert rel_max(out, ref) < 1e-6
    dres = torch.empty_like(g)
    dz = ops.empty16(R, Cc, device="cuda", grad=True)
    db, ds = torch.zeros(Cc, device="cuda"), torch.zeros(Cc, device="cuda")
    L.check(L.lib().vilco_resid_branch_bwd(ops._p(g), ops._p(rm), ops._p(y), ops._p(bias), ops._p(scale), ops._p(ymul), ops._p(dres),
                                           ops._p(dz), ops._i64(ops.lo(dz)), ops._p(db), ops._p(ds), R, Cc, C.c_float(p),
                                           C.c_uint64(seed), st))
    t = g * ymul[:, None] * keep
    assert rel_max(dres, g * rm[:, None]) < 1e-6
    assert rel_max(ops.merge16(dz) * ops.ginv(), t * scale) < 2e-5        # gradient planes are stored times GRAD_SCALE
    assert rel_max(db, (t * scale).sum(0)) < 1e-5
    assert rel_max(ds, (t * (y + bias)).sum(0)) < 1e-5


def test_training_mode_step_runs_with_dropout_and_changes_with_the_seed():
    """model.train(): dropout / stochastic depth / XLNet dropout active.  Two calls draw different masks (different losses),
    every parameter still receives a finite gradient."""
    cfg = GG.small_cfg()
    model, P = build_pair(cfg, 0)
    videos = PR.synth_video_list(cfg, 2, seed=0, lens=[128, 100], text_lens=[40, 57], n_gt=[3, 2])
    model.train()
    vals = []
    for _ in range(2):
        model.zero_grad()
        model.loss_normalizer = cfg.init_loss_norm
        out = model(videos, is_training=True)
        out["final_loss"].backward()
        vals.append(float(out["final_loss"].detach()))
    assert vals[0] != vals[1]
    det = 0.4416                                           # the deterministic (eval-mode) loss of this batch
    assert all(abs(v - det) < 0.5 * det for v in vals)
    for k, prm in model.named_parameters():
        if k in P:
            assert prm.grad is not None and torch.isfinite(prm.grad).all(), k


def test_flat_adamw_compaction_parks_dead_parameters():
    """Parameters the backward never writes are moved behind the live region: like torch.optim.AdamW with .grad = None they
    are neither updated nor decayed, and the values / moments of the live ones survive the re-layout."""
    from vilco_b200.trainer import FlatAdamW
    torch.manual_seed(1)
    ps = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in [(16, 8), (5,), (7, 3), (11,)]]
    before = [p.detach().clone() for p in ps]
    opt = FlatAdamW([{"params": ps[:3], "weight_decay": 0.1}, {"params": ps[3:], "weight_decay": 0.0}], lr=1e-2)
    for p in ps:
        p.grad.add_(torch.randn_like(p))
    opt.step()
    m_before = [opt.exp_avg[opt.slots[id(p)][0]:opt.slots[id(p)][0] + p.numel()].clone() for p in ps]
    after1 = [p.detach().clone() for p in ps]
    opt.compact({id(ps[1])})
    assert opt.live_ranges()[0][1] - opt.live_ranges()[0][0] < opt.segments[0][2] - opt.segments[0][0]
    for p, a, m in zip(ps, after1, m_before):
        o, k = opt.slots[id(p)]
        assert torch.equal(p.detach(), a) and torch.equal(opt.exp_avg[o:o + k], m)
        assert p.grad.data_ptr() == opt.flat_grad[o:o + k].data_ptr()
    opt.zero_grad()
    for p in ps:
        p.grad.add_(torch.randn_like(p))
    opt.step()
    assert torch.equal(ps[1].detach(), after1[1])                      # parked: untouched (no update, no weight decay)
    assert all(not torch.equal(p.detach(), a) for p, a in zip([ps[0], ps[2], ps[3]], [after1[0], after1[2], after1[3]]))
    assert all(not torch.equal(a, b) for a, b in zip(after1, before))


def test_fast_bf16_backward_mode_is_bounded():
    """VILCO_BWD_PRECISION=bf16: backward GEMMs read one bf16 plane per operand.  Forward / losses are unchanged;
    gradients stay within a few 1e-3 (relative L2) of the split-precision ones."""
    from vilco_b200 import ops
    cfg, model, videos, Pg, lo_, out = _model_grads()
    ref = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    ref_loss = float(out["final_loss"].detach())
    old = ops.BWD_PRECISION
    ops.BWD_PRECISION = "bf16"
    try:
        model.zero_grad()
        model.loss_normalizer = cfg.init_loss_norm
        out2 = model(videos, is_training=True)
        out2["final_loss"].backward()
    finally:
        ops.BWD_PRECISION = old
    assert abs(float(out2["final_loss"].detach()) - ref_loss) <= 1e-6 * ref_loss     # same forward (sums use float atomics)
    gmax = max(float(g.abs().max()) for g in ref.values())
    errs = []
    for k, p in model.named_parameters():
        if k in ref and float(ref[k].abs().max()) > 1e-6 * gmax:
            errs.append(float((p.grad - ref[k]).norm() / ref[k].norm()))
    assert max(errs) < 5e-2 and float(np.median(errs)) < 1e-2, (max(errs), float(np.median(errs)))


def test_shipped_mixed_mode_training_step_is_within_the_bar():
    """The shipped operand policy ("mixed": single fp16 planes except the sensitive contractions; bf16 hi+lo gradient planes
    against single-plane activations / weights in the backward): losses within 1e-3 of the oracle; gradients against the exact
    split-operand ones within the bounds test_model_gradients_vs_oracle_autograd uses for GPU-vs-oracle (a ReLU pre-activation
    within rounding of zero takes the other side and moves that element's whole gradient)."""
    from util import precision
    cfg, model, videos, Pg, lo_, out = _model_grads()            # exact mode (module default)
    ref = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    with precision("mixed"):
        model.zero_grad()
        model.loss_normalizer = cfg.init_loss_norm
        out2 = model(videos, is_training=True)
        out2["final_loss"].backward()
        torch.cuda.synchronize()
    for k in ("cls_loss", "reg_loss", "al_loss", "final_loss"):
        assert abs(float(out2[k]) - float(lo_[k])) <= 1e-3 * abs(float(lo_[k])) + 1e-6, k
    gmax = max(float(g.abs().max()) for g in ref.values())
    errs = {}
    for k, p in model.named_parameters():
        if k in ref and float(ref[k].abs().max()) > 1e-6 * gmax:
            errs[k] = float((p.grad - ref[k]).norm() / ref[k].norm())
    worst = max(errs.items(), key=lambda kv: kv[1])
    print("mixed-mode gradients: median rel-L2", float(np.median(list(errs.values()))), "worst", worst)
    assert worst[1] < 1e-1 and float(np.median(list(errs.values()))) < 2e-2, worst


def test_vilco_training_step_matches_reference_golden():
    """mq_vilco.yaml training branches — L2P prompts chosen by task id + pull constraint, temporal adapters, narration SSL
    with a memory bank — against losses and gradients produced by the REFERENCE itself (tests/golden/train_vilco.npz)."""
    import os
    from conftest import GOLDEN
    from util import build_vilco_train_pair
    g = np.load(os.path.join(GOLDEN, "train_vilco.npz"))
    cfg = GG.vilco_train_cfg()
    model, P = build_vilco_train_pair(cfg)
    model.train()
    model.n_known = 1
    model.memory_bank.memory = GG.seeded_memory_bank(48, 1024).cuda()
    model.memory_bank.ptr = 0
    videos = GG.narration_inputs(cfg, PR.synth_video_list(cfg, 3, seed=6, lens=[1024, 900, 700], text_lens=[40, 57, 33],
                                                           n_gt=[3, 2, 4]))
    model.loss_normalizer = cfg.init_loss_norm
    out = model(videos, task_id=1, is_training=True)
    out["final_loss"].backward()
    for k in ("cls_loss", "reg_loss", "al_loss", "ssl_loss", "final_loss"):
        ref = float(g["loss_" + k])
        assert abs(float(out[k].detach()) - ref) <= 1e-3 * abs(ref) + 1e-6, (k, float(out[k].detach()), ref)
    assert rel_max(model.memory_bank.memory[:4], g["memory_after"]) < 1e-4
    named = dict(model.named_parameters())
    gmax = max(float(g[k][0]) for k in g.files if k.startswith("g:"))
    strict = ("cls_head.cls_head", "reg_head.offset_head", "mu", "sigma", "narration_encoder", "prompt.")
    n = 0
    for key in g.files:
        if not key.startswith("g:"):
            continue
        ref = g[key]
        p = named[key[2:]]
        if ref[0] < 1e-6 * gmax:
            continue
        assert p.grad is not None, key
        mine = p.grad.detach().reshape(-1).double().cpu()
        got = np.concatenate([[mine.norm().item(), mine.sum().item()], mine[:8].numpy()])
        tol = 1e-3 if key[2:].startswith(strict) else 5e-2     # (ReLU gate flips perturb everything below them, see above)
        assert abs(got[0] - ref[0]) <= tol * ref[0], (key, got[0], ref[0])
        assert np.abs(got[2:] - ref[2:]).max() <= tol * ref[0] + 1e-7, (key, got[2:4], ref[2:4])
        n += 1
    assert n > 250
    # the temporal adapters are trained (the reference un-freezes them and EMA-averages them): their gradients reach the
    # registered parameters although the engines address them as pets.<i>.*
    assert sum(1 for k in g.files if ".adapters.attn." in k and g[k][0] >= 1e-6 * gmax) >= 10
    for i in range(len(cfg.adapt_blocks)):
        assert model.pets[i].layer[0].weight.grad is not None and float(model.pets[i].layer[0].weight.grad.abs().max()) > 0


def test_flat_and_torch_optimizer_checkpoints_are_interchangeable():
    """make_optimizer(flat=True).state_dict() loads into make_optimizer(flat=False) and back with every moment on the SAME
    parameter: the state is indexed by the position in the (alphabetical) groups in both, whatever the flat layout order."""
    from vilco_b200.trainer import make_optimizer
    cfg = GG.small_cfg()
    oc = {"type": "AdamW", "learning_rate": 1e-3, "weight_decay": 0.05}
    m1, _ = build_pair(cfg, 0)
    m2, _ = build_pair(cfg, 0)
    o_flat, o_torch = make_optimizer(m1, oc, flat=True), make_optimizer(m2, oc, flat=False)
    n1 = {id(p): k for k, p in m1.named_parameters()}
    n2 = {id(p): k for k, p in m2.named_parameters()}
    assert [[n1[id(p)] for p in g["params"]] for g in o_flat.param_groups] == \
        [[n2[id(p)] for p in g["params"]] for g in o_torch.param_groups]
    # give every parameter of the flat optimizer a recognisable state: exp_avg = hash of its name
    o_flat.t = 3
    tag = {k: float(sum(map(ord, k)) % 997) for k in n1.values()}
    for pid, (o, k) in o_flat.slots.items():
        o_flat.exp_avg[o:o + k] = tag[n1[pid]]
        o_flat.exp_avg_sq[o:o + k] = 2 * tag[n1[pid]]
    for p in m2.parameters():
        p.grad = torch.zeros_like(p)
    o_torch.step()                                       # creates torch's state entries
    o_torch.load_state_dict(o_flat.state_dict())
    for p, st in o_torch.state.items():
        assert float(st["exp_avg"].flatten()[0]) == tag[n2[id(p)]] and float(st["exp_avg_sq"].flatten()[0]) == 2 * tag[n2[id(p)]]
    m3, _ = build_pair(cfg, 0)
    o3 = make_optimizer(m3, oc, flat=True)
    o3.load_state_dict(o_torch.state_dict())
    n3 = {id(p): k for k, p in m3.named_parameters()}
    for pid, (o, k) in o3.slots.items():
        assert float(o3.exp_avg[o]) == tag[n3[pid]], n3[pid]


def test_eval_graph_applies_bic_bias_layers():
    """BiC with n_known > 0: the captured evaluation graph corrects the logits like the eager forward (meta_archs.py:822-836)"""
    from util import match_detections
    from vilco_b200.modeling.meta_archs import BiasLayer
    cfg = GG.small_cfg()
    model, P = build_pair(cfg, 0)
    videos = PR.synth_video_list(cfg, 2, seed=0, lens=[128, 100], text_lens=[40, 57], n_gt=[3, 2])
    plain = model(videos, is_training=False)
    model.cl_name, model.n_known, model.list_splits = "bic", 3, [3, 6]
    model.list_bias_layers = [BiasLayer().cuda(), BiasLayer().cuda()]
    for bl, (a, b) in zip(model.list_bias_layers, ((1.3, 0.2), (0.7, -0.3))):
        bl.alpha.data.fill_(a)
        bl.beta.data.fill_(b)
    eager = model(videos, is_training=False)
    out = model.make_eval_graph(2, text_len=64).run(videos)
    assert float((eager[0]["scores"] - plain[0]["scores"]).abs().max()) > 1e-3          # the correction changes the result
    for i in range(2):
        ds, swaps, dseg, orphans = match_detections(out[i], eager[i]["segments"].numpy(), eager[i]["scores"].numpy(),
                                                    eager[i]["labels"].numpy())
        assert ds < 1e-5 and swaps <= 2 and orphans <= 1, (ds, swaps, orphans)


@pytest.mark.parametrize("name", ["bic", "icarl"])
def test_distillation_terms_match_reference_golden(name):
    """n_known > 0: BiC bias layers + soft-target distillation / iCaRL BCE distillation (meta_archs.py:823-836, 1482-1519)
    against losses and gradients produced by the reference (tests/golden/distill_small.npz)."""
    import os
    from conftest import GOLDEN
    from vilco_b200.modeling.meta_archs import BiasLayer
    g = np.load(os.path.join(GOLDEN, "distill_small.npz"))
    cfg = GG.small_cfg()
    model, P = build_pair(cfg, 0)
    model.eval()
    model.cl_name, model.n_known = name, 3
    if name == "bic":
        model.list_splits = [3, 6]
        model.list_bias_layers = [BiasLayer().cuda(), BiasLayer().cuda()]
        for bl, (a, b) in zip(model.list_bias_layers, ((1.1, 0.05), (0.9, -0.02))):
            bl.alpha.data.fill_(a)
            bl.beta.data.fill_(b)
    videos = PR.synth_video_list(cfg, 2, seed=0, lens=[128, 100], text_lens=[40, 57], n_gt=[3, 2])
    prev = GG.prev_logits(cfg, probs=True)
    model.loss_normalizer = cfg.init_loss_norm
    with torch.no_grad():                       # validate_loss path (train_utils.py:584-655): same values, no graph
        model.loss_normalizer = cfg.init_loss_norm
        quiet = model(videos, is_training=True, prev_out_cls_logits=prev if name == "bic" else [prev])
    model.loss_normalizer = cfg.init_loss_norm
    out = model(videos, is_training=True, prev_out_cls_logits=prev if name == "bic" else [prev])
    out["final_loss"].backward()
    for k in ("cls_loss", "reg_loss", "al_loss", "dist_loss", "final_loss"):
        ref = float(g[f"{name}_loss_{k}"])
        assert abs(float(out[k].detach()) - ref) <= 1e-3 * abs(ref) + 1e-6, (k, float(out[k].detach()), ref)
        assert abs(float(quiet[k]) - ref) <= 1e-3 * abs(ref) + 1e-6, ("no_grad", k, float(quiet[k]), ref)
    if name == "bic":
        got = np.array([[bl.alpha.grad.item(), bl.beta.grad.item()] for bl in model.list_bias_layers])
        assert np.abs(got - g["bic_bias_grads"]).max() <= 1e-3 * np.abs(g["bic_bias_grads"]).max()
    named = dict(model.named_parameters())
    for key in ("cls_head.cls_head.conv.weight", "cls_head.cls_head.conv.bias", "cls_head.norm.1.weight", "mu"):
        ref = g[f"{name}_g:{key}"]
        mine = named[key].grad.detach().reshape(-1).double().cpu()
        got = np.concatenate([[mine.norm().item(), mine.sum().item()], mine[:8].numpy()])
        assert np.abs(got - ref).max() <= 2e-3 * ref[0] + 1e-8, (key, got[:3], ref[:3])


def test_ragged_training_batch_matches_oracle():
    """clips and texts of very different lengths, 1 / 8 / 3 ground truths per clip: losses vs the oracle, and a backward
    pass that leaves finite gradients.  (A clip without labels cannot be part of a cross-modal training batch in the
    reference either: it drops the clip's features but not its text, meta_archs.py:1138 vs :1188.)"""
    cfg = GG.small_cfg()
    model, P = build_pair(cfg, 0)
    model.eval()
    videos = PR.synth_video_list(cfg, 3, seed=4, lens=[128, 64, 37], text_lens=[20, 128, 77], n_gt=[1, 8, 3])
    lo, _ = O.model_train_losses(P, cfg, videos)
    model.loss_normalizer = cfg.init_loss_norm
    out = model(videos, is_training=True)
    for k in ("cls_loss", "reg_loss", "al_loss", "final_loss"):
        assert abs(float(out[k].detach()) - float(lo[k])) <= 1e-3 * abs(float(lo[k])) + 1e-6, (k, float(out[k].detach()), float(lo[k]))
    out["final_loss"].backward()
    for k, prm in model.named_parameters():
        if k in P:
            assert prm.grad is not None and torch.isfinite(prm.grad).all(), k


def test_flat_adamw_checkpoint_round_trip_and_torch_compatibility():
    """FlatAdamW.state_dict() has torch.optim.AdamW's structure: resuming from it (in a new FlatAdamW or in torch's AdamW)
    continues with identical updates."""
    from vilco_b200.trainer import FlatAdamW
    torch.manual_seed(3)
    shapes = [(32, 16), (16,), (8, 4, 1)]
    base = [torch.randn(s, device="cuda") for s in shapes]
    grads = [[torch.randn(s, device="cuda") for s in shapes] for _ in range(4)]

    def mk(cls):
        ps = [torch.nn.Parameter(b.clone()) for b in base]
        return ps, cls([{"params": ps[:2], "weight_decay": 0.05}, {"params": ps[2:], "weight_decay": 0.0}], lr=1e-2)

    def run(ps, opt, gs):
        for g in gs:
            opt.zero_grad()
            for p, gi in zip(ps, g):
                if p.grad is None:
                    p.grad = gi.clone()
                else:
                    p.grad.add_(gi)
            opt.step()

    ps_a, opt_a = mk(FlatAdamW)
    run(ps_a, opt_a, grads[:2])
    sd = opt_a.state_dict()
    assert set(sd["state"].keys()) == {0, 1, 2} and float(sd["state"][0]["step"]) == 2.0
    run(ps_a, opt_a, grads[2:])                                  # uninterrupted run = the reference trajectory
    for cls in (FlatAdamW, torch.optim.AdamW):
        ps_b, opt_b = mk(cls)
        ps_c, opt_c = mk(FlatAdamW)
        run(ps_c, opt_c, grads[:2])                              # parameters after two steps ...
        with torch.no_grad():
            for p, q in zip(ps_b, ps_c):
                p.copy_(q)
        opt_b.load_state_dict(sd)                                # ... plus the checkpointed moments
        run(ps_b, opt_b, grads[2:])
        for p, q in zip(ps_b, ps_a):
            assert rel_max(p.detach(), q.detach()) < 5e-6, cls


def test_training_loop_drives_the_loss_down():
    """25 optimisation steps of trainer.Trainer (FlatAdamW, clip 1.0, dropout / drop-path ON) on one fixed batch."""
    from vilco_b200.trainer import Trainer, make_optimizer
    cfg = GG.small_cfg()
    model, P = build_pair(cfg, 0)
    model.train()
    model.loss_normalizer_momentum = 1.0
    videos = PR.synth_video_list(cfg, 4, seed=3, lens=[128, 100, 90, 128], text_lens=[40, 57, 33, 64], n_gt=[3, 2, 4, 1])
    tr = Trainer(model, make_optimizer(model, {"type": "AdamW", "learning_rate": 3e-4, "weight_decay": 0.05}, flat=True),
                 clip_grad_l2norm=1.0)
    hist = [float(tr.step(videos)["final_loss"].detach()) for _ in range(25)]
    assert all(np.isfinite(hist))
    assert np.mean(hist[-3:]) < 0.4 * np.mean(hist[:3]), hist


def test_eval_graph_follows_weight_updates():
    """An EvalGraph captured before training must not keep reading the old packed weights: after optimizer steps its
    results equal the eager evaluation of the updated model."""
    from vilco_b200.trainer import Trainer, make_optimizer
    cfg = GG.small_cfg()
    model, P = build_pair(cfg, 0)
    videos = PR.synth_video_list(cfg, 2, seed=0, lens=[128, 100], text_lens=[40, 57], n_gt=[3, 2])
    model.eval()
    g = model.make_eval_graph(2, text_len=64)
    before = g.run(videos)
    model.loss_normalizer = cfg.init_loss_norm
    tr = Trainer(model, make_optimizer(model, {"type": "AdamW", "learning_rate": 1e-2, "weight_decay": 0.05}, flat=True), 1.0)
    for _ in range(3):
        tr.step(videos)
    with torch.no_grad():
        eager = model(videos, is_training=False)
    after = g.run(videos)
    assert rel_max(after[0]["scores"], eager[0]["scores"]) < 1e-4
    assert rel_max(after[0]["scores"], before[0]["scores"]) > 1e-3      # the update is visible


@pytest.mark.parametrize("B,H,T,d,W,valid", [(2, 2, 64, 64, 9, [64, 41]), (1, 4, 200, 96, 19, [173]), (2, 2, 37, 32, 5, [37, 20])])
def test_local_attention_backward(B, H, T, d, W, valid):
    """LocalMaskedMHCA core (window W): dq / dk / dv of vilco_local_attention_bwd vs torch autograd through the same banded
    attention written densely (scale 1/sqrt(d), -1e4 on padded keys, -inf outside the band / sequence, padded query rows
    zeroed — blocks.py:1140-1200); head dim 96 = the NLQ configuration."""
    from vilco_b200 import backward as BW
    from vilco_b200 import ops
    torch.manual_seed(0)
    C, w = H * d, W // 2
    mask = (torch.arange(T, device="cuda")[None, :] < torch.tensor(valid, device="cuda")[:, None]).float().contiguous()
    q, k, v = (ops.merge16(ops.split16(torch.randn(B, T, C, device="cuda"))).requires_grad_(True) for _ in range(3))
    qh, kh, vh = (t.view(B, T, H, d).permute(0, 2, 1, 3) for t in (q, k, v))
    S = (qh / math.sqrt(d)) @ kh.transpose(-1, -2)
    i = torch.arange(T, device="cuda")
    band = (i[:, None] - i[None, :]).abs() <= w
    S = S + (-1e4) * (1 - mask)[:, None, None, :]
    S = S.masked_fill(~band[None, None], float("-inf"))
    P = torch.softmax(S, -1) * mask[:, None, :, None]
    O = (P @ vh).permute(0, 2, 1, 3).reshape(B, T, C)
    q16, k16, v16 = (ops.split16(t.detach()) for t in (q, k, v))
    out = ops.merge16(ops.local_attention(q16, k16, v16, mask, H, W))
    assert rel_max(out, O.detach()) < 2e-5
    dO = torch.randn(B, T, C, device="cuda")
    O.backward(dO)
    dq, dk, dv = BW.local_attention_bwd(dO, q16, k16, v16, mask, H, W)
    for name, got, ref in (("dq", dq, q.grad), ("dk", dk, k.grad), ("dv", dv, v.grad)):
        assert rel_max(got, ref) < 2e-5, name


def test_class_incremental_flow():
    """The query-incremental loop in miniature (what train_cl.py does between sub-tasks): train on 3 classes, evaluate,
    augment_classification(+3), new optimizer, train on all 6 classes, evaluate.  Checks that the grown classifier / the
    grown mu, sigma are live in the CUDA path, that old-class weights are carried over, and that inference follows."""
    from vilco_b200.config import mq_model_kwargs
    from vilco_b200.modeling import make_meta_arch
    from vilco_b200.trainer import Trainer, make_optimizer
    cfg = GG.small_cfg()
    torch.manual_seed(0)
    model = make_meta_arch("LocPointTransformer", **mq_model_kwargs(cfg.input_dim, cfg.embd_dim, cfg.n_head, cfg.max_seq_len,
                                                                    cfg.arch, 3, cfg.n_txt_in, cfg.regression_range)).cuda()
    videos = PR.synth_video_list(cfg, 4, seed=8, lens=[128, 100, 90, 128], text_lens=[40, 57, 33, 64], n_gt=[3, 2, 4, 1])
    task0 = [dict(v, labels=v["labels"] % 3) for v in videos]
    model.train()
    tr = Trainer(model, make_optimizer(model, {"type": "AdamW", "learning_rate": 3e-4, "weight_decay": 0.05}, flat=True), 1.0)
    l0 = [float(tr.step(task0)["final_loss"].detach()) for _ in range(6)]
    model.eval()
    with torch.no_grad():
        r0 = model(task0[:1], is_training=False)
    assert int(r0[0]["labels"].max()) <= 2
    w_old = model.cls_head.cls_head.conv.weight.detach().clone()
    mu_old = model.mu.detach().clone()
    model.augment_classification(3, "cuda")
    model.n_known = 3
    assert model.num_classes == 6 and model.cls_head.cls_head.conv.weight.shape[0] == 6 and model.mu.shape[0] == 6
    assert torch.equal(model.cls_head.cls_head.conv.weight[:3].detach(), w_old) and torch.equal(model.mu[:3].detach(), mu_old)
    model.train()
    tr = Trainer(model, make_optimizer(model, {"type": "AdamW", "learning_rate": 3e-4, "weight_decay": 0.05}, flat=True), 1.0)
    l1 = [float(tr.step(videos)["final_loss"].detach()) for _ in range(6)]
    assert all(np.isfinite(l0 + l1)) and l1[-1] < l1[0]
    assert not torch.equal(model.cls_head.cls_head.conv.weight[3:].detach(), torch.zeros_like(model.cls_head.cls_head.conv.weight[3:]))
    model.eval()
    with torch.no_grad():
        logits, _, _ = model(videos[:1], is_training=False, get_emb=True)
        r1 = model(videos[:1], is_training=False)
    assert logits[0].shape[-1] == 6 and len(r1[0]["scores"]) > 0


@pytest.mark.parametrize("B,T,H,valid", [(2, 256, 2, [256, 100]), (1, 1024, 4, [777])])
def test_attention_backward_with_fused_epilogues(B, T, H, valid):
    """backward.attention_bwd_lse — probabilities recomputed in the QK^T GEMM epilogue from the forward kernel's row
    log-sum-exp, softmax backward fused into the dO V^T GEMM epilogue — against torch autograd of the same attention on the same
    fp16 operands (float64), and against the materialised chain it replaces"""
    from util import precision
    from vilco_b200 import backward as BW
    from vilco_b200 import ops
    with precision("mixed"):
        gen = torch.Generator(device="cuda").manual_seed(B * T + H)
        C = H * 64
        q16, k16, v16 = (ops.split16(torch.randn(B, T, C, device="cuda", generator=gen) * s, planes=1) for s in (1.2, 1.2, 1.0))
        vl = torch.tensor(valid, device="cuda")
        kmask = (torch.arange(T, device="cuda")[None, :] < vl[:, None]).float()
        v16 = v16 * kmask[None, :, :, None].to(v16.dtype)                 # the forward multiplies v by the key mask (blocks.py:394)
        dO = torch.randn(B, T, C, device="cuda", generator=gen) * 1e-2
        O16, lse = ops.self_attention(q16, k16, v16, kmask, H, 0.125, want_lse=True)
        dq, dk, dv = BW.attention_bwd_lse(dO, O16, lse, q16, k16, v16, kmask, H, 0.125)
        dq0, dk0, dv0 = BW.attention_bwd(dO, q16, k16, v16, kmask, H, 0.125)
        q, k, v = (t[0].double().requires_grad_(True) for t in (q16, k16, v16))
        hd = lambda t: t.view(B, T, H, 64).transpose(1, 2)     # noqa: E731
        att = (hd(q) @ hd(k).transpose(-2, -1)) * 0.125
        att = att.masked_fill(~(kmask > 0)[:, None, None, :], float("-inf"))
        o = (torch.softmax(att, -1) @ hd(v)).transpose(1, 2).reshape(B, T, C)
        lse_ref = torch.logsumexp(att, -1) * 1.4426950408889634
        assert float((lse.double() - lse_ref).abs().max()) < 2e-3
        gq, gk, gv = torch.autograd.grad(o, (q, k, v), dO.double())
        for name, got, old, ref in (("dq", dq, dq0, gq), ("dk", dk, dk0, gk), ("dv", dv, dv0, gv)):
            e_new = float((got.double() - ref).norm() / ref.norm())
            e_old = float((old.double() - ref).norm() / ref.norm())
            print(f"attention bwd T{T} {name}: fused epilogues {e_new:.2e}, materialised chain {e_old:.2e}")
            assert e_new < 2e-3
