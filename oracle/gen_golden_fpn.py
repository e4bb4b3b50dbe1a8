"""TEST INFRASTRUCTURE ONLY — golden vectors for FPN1D (`fpn_type: fpn`) from the REFERENCE's own module
(MQ/libs/modeling/necks.py:13-106, with ACConv / DenseAPP from modeling/utils.py).  Authoring container only:

    python -m oracle.gen_golden_fpn        ->  tests/golden/fpn1d.npz  (inputs, masks, outputs; the weights are the seeded
                                                function of name and shape `fpn_state`, so they are not stored)
"""
import os

import numpy as np
import torch

from . import params as PR

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
C, LEVELS, T0, B = 128, 4, 32, 2      # C % 128 == 0: what the depthwise-conv + LN kernel supports (MQ: 1024)
VALID = (32, 21)


def fpn_spec(C=C, levels=LEVELS, pre=""):
    """state_dict layout of the reference FPN1D (registration order)"""
    sp = {}
    for i in range(levels):
        sp[f"{pre}lateral_convs.{i}.conv.weight"] = (C, C, 1)
    d = pre + "ac_conv.denseapp."
    cin = C
    for r in (3, 6, 12, 18, 24):
        sp[f"{d}aspp{r}.conv1x1.weight"] = (512, cin, 1); sp[f"{d}aspp{r}.conv1x1.bias"] = (512,)
        sp[f"{d}aspp{r}.ConvGN.weight"] = (512,); sp[f"{d}aspp{r}.ConvGN.bias"] = (512,)
        sp[f"{d}aspp{r}.dilaconv.weight"] = (256, 512, 3); sp[f"{d}aspp{r}.dilaconv.bias"] = (256,)
        cin += 256
    sp[d + "conv1x1.weight"] = (C, 5 * 256, 1); sp[d + "conv1x1.bias"] = (C,)
    sp[d + "ConvGN.weight"] = (C,); sp[d + "ConvGN.bias"] = (C,)
    for m in ("CxAM", "CnAM"):
        a = f"{pre}ac_conv.{m}."
        for n in (("key_conv", "query_conv", "value_conv") if m == "CxAM" else ("query_conv", "key_conv", "value_conv")):
            co = C if n == "value_conv" else C // 8
            sp[f"{a}{n}.weight"] = (co, C, 1); sp[f"{a}{n}.bias"] = (co,)
    for i in range(levels):
        sp[f"{pre}fpn_convs.{i}.conv.weight"] = (C, 1, 3)
    for i in range(levels):
        sp[f"{pre}fpn_norms.{i}.weight"] = (1, C, 1); sp[f"{pre}fpn_norms.{i}.bias"] = (1, C, 1)
    return sp


def fpn_state(seed=11, pre=""):
    P = PR.random_state(fpn_spec(pre=pre), seed)
    for k in P:                                            # GroupNorm affine: around 1 / small, like the other norms
        if ".ConvGN." in k:
            P[k] = (1.0 + 0.1 * P[k]) if k.endswith("weight") else 0.1 * P[k]
    return P


def fpn_inputs(seed=12):
    g = torch.Generator().manual_seed(seed)
    feats, masks = [], []
    for l in range(LEVELS):
        T = T0 >> l
        m = (torch.arange(T0)[None, :] < torch.tensor(VALID)[:, None])[:, ::(1 << l)][:, None, :]
        feats.append(torch.randn(B, C, T, generator=g) * m)
        masks.append(m)
    return feats, masks


def main():
    from . import ref_shim
    ref = ref_shim.load()
    neck = ref.necks.FPN1D(in_channels=[C] * LEVELS, out_channel=C).eval()
    sd = neck.state_dict()
    spec = fpn_spec()
    assert list(sd.keys()) == list(spec.keys()), [k for k in sd if k not in spec][:5]
    assert all(tuple(sd[k].shape) == tuple(spec[k]) for k in sd)
    neck.load_state_dict(fpn_state())
    feats, masks = fpn_inputs()
    with torch.no_grad():
        out, out_masks = neck(feats, masks)
    save = {}
    for l in range(LEVELS):
        save[f"out_{l}"] = out[l].numpy()
        assert (out_masks[l] == masks[l]).all()
    np.savez_compressed(os.path.join(GOLDEN, "fpn1d.npz"), **save)
    print("fpn1d.npz", {k: v.shape for k, v in save.items()})


if __name__ == "__main__":
    main()
