"""CPU emulation of the operand-format policy of the tensor-core GEMMs (TEST / DESIGN TOOL, not on the product path).

Runs the oracle at the full mq_no_cl.yaml configuration with the operands of every dense contraction rounded to a 16-bit
format exactly where the CUDA path rounds them (producer output planes and packed weights), fp32 accumulation, and reports
max|d|/max|ref| of logits / offsets against the un-rounded oracle.  A policy maps call sites (the oracle function the
contraction sits in) to a format:  "f32" (exact, = the split hi+lo modes to ~1e-5), "fp16", "bf16".

    python tools/precision_sweep.py                 # the canned policies
"""
import contextlib
import inspect
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mq_oracle as O  # noqa: E402
from oracle import params as PR  # noqa: E402


def rnd(t, fmt):
    if fmt == "f32" or not torch.is_floating_point(t):
        return t
    return t.to(torch.float16 if fmt == "fp16" else torch.bfloat16).to(t.dtype)


SITES = ("masked_conv1d", "_qkv_conv_norm", "masked_mhca", "masked_mha_cross", "channel_block", "transformer_block",
         "xlnet_layer", "adapter_time")


CTX = ("_head_tower", "cls_head", "reg_head", "backbone")


def _site():
    """(innermost oracle operator the contraction sits in, the model part that called it)"""
    site, ctx, blk = "other", "other", ""
    for fr in inspect.stack(0)[2:10]:
        if site == "other" and fr.function in SITES:
            site = fr.function
        if fr.function == "transformer_block":
            blk = fr.frame.f_locals.get("pre", "")
        if fr.function in CTX:
            ctx = fr.function
            break
    return site, ctx, blk


@contextlib.contextmanager
def policy(pol, default="fp16"):
    """pol: dict site -> fmt, or site+':'+tag -> fmt for finer keys (see the wrappers)."""
    conv1d, linear, einsum, matmul = F.conv1d, F.linear, torch.einsum, torch.Tensor.__matmul__

    def fmt_for(kind, w=None):
        s, ctx, blk = _site()
        shp = "x".join(str(int(v)) for v in w.shape) if w is not None else ""
        for key in (f"blk:{blk}{s}:{kind}:{shp}", f"blk:{blk}{s}:{kind}", f"blk:{blk}{s}", f"blk:{blk}", f"{ctx}>{s}:{shp}", f"{ctx}>{s}", f"{s}:{kind}:{shp}", f"{s}:{kind}", s):
            if key in pol:
                return pol[key]
        return default

    def q_conv1d(x, w, b=None, stride=1, padding=0, dilation=1, groups=1):
        if groups == 1:
            f = fmt_for("conv", w)
            x, w = rnd(x, f), rnd(w, f)
        return conv1d(x, w, b, stride, padding, dilation, groups)

    def q_linear(x, w, b=None):
        f = fmt_for("linear", w)
        return linear(rnd(x, f), rnd(w, f), b)

    def q_einsum(eq, *ops):
        f = fmt_for("einsum")
        return einsum(eq, *[rnd(o, f) for o in ops])

    def q_matmul(a, b):
        f = fmt_for("matmul")
        return matmul(rnd(a, f), rnd(b, f))

    F.conv1d, F.linear, torch.einsum, torch.Tensor.__matmul__ = q_conv1d, q_linear, q_einsum, q_matmul
    try:
        yield
    finally:
        F.conv1d, F.linear, torch.einsum, torch.Tensor.__matmul__ = conv1d, linear, einsum, matmul


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def main():
    K = int(os.environ.get("K", 22))
    c = O.ModelCfg(num_classes=K)
    P = PR.random_state(PR.param_spec(c), 4)
    vids = PR.synth_video_list(c, 2, seed=8, lens=[1024, 700], text_lens=[57, 33], n_gt=[4, 3])
    which = [int(a) for a in os.environ.get("CLIPS", "1").split(",")]

    def run():
        out = []
        with torch.no_grad():
            for i in which:
                x, m, t, tm = O.preprocess(c, [vids[i]], False)
                lg, of, _, _ = O.forward_heads(P, c, x, m, t, tm, False)
                out.append((torch.cat(lg, 1), torch.cat(of, 1)))
        return out

    ref = run()
    C = c.embd_dim
    qkv = f"channel_block:linear:{3 * C}x{C}"
    final = {f"cls_head>masked_conv1d:{K}x{C}x3": "f32", f"reg_head>masked_conv1d:2x{C}x3": "f32"}
    pre_exact = {"backbone>masked_conv1d": "f32", "blk:backbone.stem.0.": "f32", "blk:backbone.txt_stem.0.": "f32",
                 "blk:backbone.txt_stem.1.": "f32", **final}
    s1 = {"blk:backbone.stem.1.channel_block:linear:%dx%d" % (3 * C, C): "f32", "blk:backbone.stem.1.channel_block:matmul": "f32"}
    chan_all = {qkv: "f32", "channel_block:matmul": "f32"}
    vid_in = {f"backbone>masked_conv1d:{C}x4096x1": "f32", f"backbone>masked_conv1d:{C}x{C}x3": "f32"}
    pols = {
        "G: proj+embd, chan qkv/core, final exact; rest fp16": ({**vid_in, **final, **chan_all}, "fp16"),
        "H: G + stem.0 chan proj/mlp exact": ({**vid_in, **final, **chan_all, "blk:backbone.stem.0.channel_block": "f32"}, "fp16"),
        "I: G + txt_embd exact": ({"backbone>masked_conv1d": "f32", **final, **chan_all}, "fp16"),
        "D: A but proj+embd fp16": ({"blk:backbone.stem.0.": "f32", **final, **s1, **chan_all}, "fp16"),
        "E: only chan qkv/core + stem.0 chan block + final exact": ({"blk:backbone.stem.0.channel_block": "f32", **final, **chan_all}, "fp16"),
        "F: B with bf16 for the rest": ({"backbone>masked_conv1d": "f32", "blk:backbone.stem.0.": "f32", **final, **s1,
                                         qkv: "f32", "channel_block:matmul": "f32"}, "bf16"),
        "A: prefix (proj,embd,txt,stem.0) + stem.1 chan qkv/core + final exact": ({**pre_exact, **s1}, "fp16"),
        "B: A but txt path fp16 except chan qkv/core": ({"backbone>masked_conv1d": "f32", "blk:backbone.stem.0.": "f32", **final, **s1,
                                                         qkv: "f32", "channel_block:matmul": "f32"}, "fp16"),
        "C: A but stem.0 mhca+mlp fp16": ({**pre_exact, **s1, "blk:backbone.stem.0.masked_mhca": "fp16",
                                           "blk:backbone.stem.0._qkv_conv_norm": "fp16",
                                           "blk:backbone.stem.0.transformer_block": "fp16"}, "fp16"),
        "fp16, chan qkv+core exact": ({qkv: "f32", "channel_block:matmul": "f32"}, "fp16"),
        "fp16, chan qkv+core, final convs exact": ({qkv: "f32", "channel_block:matmul": "f32", **final}, "fp16"),
        "fp16, chan qkv+core, final, proj exact": ({qkv: "f32", "channel_block:matmul": "f32", **final,
                                                    f"backbone>masked_conv1d:{C}x4096x1": "f32"}, "fp16"),
        "only head tower fp16": ({"_head_tower>masked_conv1d": "fp16"}, "f32"),
        "only final convs fp16": ({k_: "fp16" for k_ in final}, "f32"),
        "only backbone convs (proj, embd, txt) fp16": ({"backbone>masked_conv1d": "fp16"}, "f32"),
        "only chan qkv fp16": ({qkv: "fp16"}, "f32"),
        "only chan core fp16": ({"channel_block:matmul": "fp16"}, "f32"),
        "only chan proj+mlp fp16": ({"channel_block:linear": "fp16", qkv: "f32"}, "f32"),
        "all fp16": ({}, "fp16"),
        "all bf16": ({}, "bf16"),
        "fp16, channel-attention core exact": ({"channel_block:matmul": "f32"}, "fp16"),
        "fp16, attention cores exact": ({"channel_block:matmul": "f32", "masked_mhca:matmul": "f32",
                                         "masked_mha_cross:matmul": "f32", "xlnet_layer:einsum": "f32"}, "fp16"),
        "fp16, head convs exact": ({"masked_conv1d": "f32"}, "fp16"),
        "only head/embed convs fp16": ({"masked_conv1d": "fp16"}, "f32"),
        "only xlnet fp16": ({"xlnet_layer": "fp16"}, "f32"),
        "only channel block fp16": ({"channel_block": "fp16"}, "f32"),
        "only mhca fp16": ({"masked_mhca": "fp16", "_qkv_conv_norm": "fp16", "masked_mha_cross": "fp16"}, "f32"),
        "only mlp fp16": ({"transformer_block": "fp16"}, "f32"),
    }
    sel = sys.argv[1:]
    for name, (pol, default) in pols.items():
        if sel and not any(s in name for s in sel):
            continue
        with policy(pol, default):
            got = run()
        e = [(rel(g[0], r[0]), rel(g[1], r[1])) for g, r in zip(got, ref)]
        print(f"{name:40s} " + "  ".join(f"clip{i}: logits {a:.2e} offsets {b:.2e}" for i, (a, b) in zip(which, e)), flush=True)


if __name__ == "__main__":
    main()
