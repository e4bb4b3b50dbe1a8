"""FPN1D (`fpn_type: fpn`): the oracle restatement against the golden produced by the reference's own module
(oracle/gen_golden_fpn.py), and the mirror's state_dict layout."""
import os

import numpy as np
import torch

from conftest import GOLDEN
from oracle import mq_oracle as O
from oracle.gen_golden_fpn import C, LEVELS, fpn_inputs, fpn_spec, fpn_state


def test_oracle_fpn1d_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, "fpn1d.npz"))
    P = fpn_state(pre="neck.")
    feats, masks = fpn_inputs()
    with torch.no_grad():
        out, _ = O.fpn1d(P, feats, masks)
    for l in range(LEVELS):
        ref = g[f"out_{l}"]
        assert np.abs(out[l].numpy() - ref).max() <= 2e-5 * np.abs(ref).max(), l


def test_fpn1d_mirror_state_dict_equals_reference_layout():
    from vilco_b200.modeling import make_neck
    neck = make_neck("fpn", in_channels=[C] * LEVELS, out_channel=C)
    sd, spec = neck.state_dict(), fpn_spec()
    assert list(sd.keys()) == list(spec.keys())
    assert all(tuple(sd[k].shape) == tuple(spec[k]) for k in sd)
