"""Data path: vilco_resize_feats against the reference's own call — F.interpolate(mode='linear',
align_corners=False) of the permuted clip on the CPU (MQ/libs/datasets/ego4d.py:644-651).  Tolerance: 1e-6 absolute on
N(0,1) features (the two products and their sum may round one ulp apart between ATen's vectorised CPU kernel and the GPU)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
PRECISION = "fp16x3"   # the tolerances below state the exact (split-operand) arithmetic; see conftest._precision_mode


def _ref(feats_tc, T_out):
    return F.interpolate(feats_tc.permute(1, 0).unsqueeze(0), size=T_out, mode="linear", align_corners=False).squeeze(0)


@pytest.mark.parametrize("T_in,C,T_out", [(1, 64, 128), (2, 64, 128), (37, 256, 1024), (511, 4096, 1024), (1000, 512, 1024),
                                          (1024, 256, 1024), (1500, 256, 1024), (2560, 64, 192)])
def test_resize_matches_interpolate(T_in, C, T_out):
    from vilco_b200 import data
    g = torch.Generator().manual_seed(1000 + T_in)
    x = torch.randn(T_in, C, generator=g)
    want = _ref(x, T_out)
    got = data.resize_feats(x, T_out)
    assert got.is_cuda and got.shape == (C, T_out) and got.dtype == torch.float32
    err = (got.cpu() - want).abs().max().item()
    assert err <= 1e-6, (T_in, C, T_out, err)
    if T_in == T_out:
        assert torch.equal(got.cpu(), x.t())            # scale 1: a copy
    # pinned host input and device input give the same bits
    assert torch.equal(data.resize_feats(x.pin_memory(), T_out), got)
    assert torch.equal(data.resize_feats(x.cuda(), T_out), got)


def test_ragged_batch_equals_single_clips_and_packed_planes():
    from vilco_b200 import data, ops
    g = torch.Generator().manual_seed(5)
    clips = [torch.randn(t, 1024, generator=g) for t in (300, 1, 1024, 77, 640)]
    outs = data.resize_feats(clips, 1024)
    assert len(outs) == 5
    for f, o in zip(clips, outs):
        assert torch.equal(o, data.resize_feats(f, 1024))
        assert (o.cpu() - _ref(f, 1024)).abs().max().item() <= 1e-6
    planes = data.resize_pack(clips, 1024)                              # (NP, B, T, C) operand of the first GEMM
    want = ops.pack_feats(torch.stack(outs).contiguous())               # resize -> (B, C, T) -> transpose + split
    assert planes.shape == want.shape and torch.equal(planes, want)


def test_model_accepts_device_feats_from_the_data_path():
    """a clip resized on the device goes through the unchanged model call and gives the logits of the CPU-resized clip."""
    from oracle import params as PR
    from oracle.gen_golden import small_cfg
    from util import build_pair, rel_max
    from vilco_b200 import data
    cfg = small_cfg()
    model, _ = build_pair(cfg)
    T = cfg.max_seq_len
    v = PR.synth_video_list(cfg, 1, seed=4, lens=[T], text_lens=[33], n_gt=[2])[0]
    raw = torch.randn(T // 2 + 5, cfg.input_dim, generator=torch.Generator().manual_seed(3))
    a = dict(v, feats=_ref(raw, T))
    b = dict(v, feats=data.resize_feats(raw, T))
    with torch.no_grad():
        la, oa, _ = model([a], is_training=False, get_emb=True)
        lb, ob, _ = model([b], is_training=False, get_emb=True)
    assert rel_max(torch.cat(lb, 1), torch.cat(la, 1)) < 1e-5
    assert rel_max(torch.cat(ob, 1), torch.cat(oa, 1)) < 1e-5


def test_eval_graph_takes_raw_clips():
    """EvalGraph.run / infer_stream with `feats_raw` (features as stored, resized on the device) == the same clips resized
    on the CPU like the dataset does and passed as `feats`; a batch mixing both forms is refused."""
    from oracle import params as PR
    from oracle.gen_golden import small_cfg
    from util import build_pair, rel_max
    cfg = small_cfg()
    model, _ = build_pair(cfg)
    T = cfg.max_seq_len
    vids = PR.synth_video_list(cfg, 2, seed=6, lens=[T, T], text_lens=[21, 50], n_gt=[2, 3])
    gen = torch.Generator().manual_seed(8)
    raws = [torch.randn(t, cfg.input_dim, generator=gen) for t in (T - 29, T // 3)]
    cpu = [dict(v, feats=_ref(r, T)) for v, r in zip(vids, raws)]
    raw = [dict({k: x for k, x in v.items() if k != "feats"}, feats_raw=r) for v, r in zip(vids, raws)]
    g = model.make_eval_graph(2, text_len=64)
    want = g.run(cpu)
    got = g.run(raw)
    streamed = [r for res in g.infer_stream([raw, raw, raw]) for r in res]
    assert len(streamed) == 6
    for res in (got, streamed[:2], streamed[4:]):
        for a, b in zip(res, want):
            assert a["video_id"] == b["video_id"] and a["segments"].shape == b["segments"].shape
            assert rel_max(a["scores"], b["scores"]) < 1e-5
            assert (a["labels"] == b["labels"]).float().mean().item() > 0.98
    with pytest.raises(ValueError, match="feats_raw"):
        g.run([raw[0], cpu[1]])


def test_validation_pass_equals_clip_by_clip_evaluation():
    """vilco_b200.utils.validate.valid_one_epoch (regrouped static batches, padded last batch, streaming graph, in-memory
    evaluator) gives the result table of the reference-style loop `for clip: model([clip])`."""
    import numpy as np
    import pandas as pd
    from oracle import params as PR
    from oracle.gen_golden import small_cfg
    from util import build_pair
    from vilco_b200.utils.metrics import ANETdetection
    from vilco_b200.utils.validate import results_table, valid_one_epoch
    cfg = small_cfg()
    model, _ = build_pair(cfg)
    T = cfg.max_seq_len
    vids = PR.synth_video_list(cfg, 5, seed=9, lens=[T, T - 20, 64, T, 90], text_lens=[21, 50, 33, 12, 60], n_gt=[2, 3, 1, 2, 2])
    for i, v in enumerate(vids):
        v["video_id"] = f"clip{i}"
    gt = pd.DataFrame({"video-id": [v["video_id"] for v in vids for _ in range(len(v["labels"]))],
                       "t-start": [float(s[0]) for v in vids for s in v["segments"]],
                       "t-end": [float(s[1]) for v in vids for s in v["segments"]],
                       "label": [int(l) for v in vids for l in v["labels"]]})
    index = {j: i for i, j in enumerate(sorted(gt["label"].unique()))}
    gt["label"] = gt["label"].map(index)
    ev = ANETdetection((gt, index), tiou_thresholds=np.linspace(0.1, 0.5, 5))
    mAP, avg, thr, rec = valid_one_epoch([[v] for v in vids], model, 0, evaluator=ev, batch_size=2, text_len=64)
    assert rec is None and mAP.shape == (5,) and 0.0 <= avg <= 1.0
    with torch.no_grad():
        want = results_table([model([v], is_training=False)[0] for v in vids])
    mAP2, avg2, _ = ANETdetection((gt, index), tiou_thresholds=np.linspace(0.1, 0.5, 5)).evaluate(want, verbose=False)
    # the batched graph and the single-clip call agree to the last bits of the scores; the mAP of both tables must agree
    assert abs(avg - avg2) < 1e-3, (avg, avg2)
