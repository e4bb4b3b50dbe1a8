"""GPU soft-NMS / hard-NMS / decode against the reference extension's known answers (tests/golden/nms.npz) and the
C oracle — kept segments identical, scores bit-exact (bar: identical segments, scores within 1e-5)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
CASES = ["n1", "n17", "n300", "n2000", "ties"]


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "nms.npz"))


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("sigma,ms", [(0.99, 1e-4), (0.5, 0.01)])
def test_softnms_single_class_bit_exact(g, name, sigma, ms):
    from vilco_b200.utils import batched_nms
    segs, scores = torch.from_numpy(g[name + "_segs"]), torch.from_numpy(g[name + "_scores"])
    dets = g[f"{name}_m2_s{sigma}_t{ms}_dets"]
    k = min(200, len(dets))
    s, sc, lb = batched_nms(segs, scores, torch.zeros(len(scores), dtype=torch.int64), 0.1, ms, 200, True, True, sigma)
    assert s.shape[0] == k
    assert (s.numpy() == dets[:k, :2]).all() and (sc.numpy() == dets[:k, 2]).all()


def test_batched_multiclass_bit_exact(g):
    from vilco_b200.utils import batched_nms
    s, sc, lb = batched_nms(torch.from_numpy(g["b_segs"]), torch.from_numpy(g["b_scores"]),
                            torch.from_numpy(g["b_labels"]), 0.1, 1e-4, 200, True, True, 0.99)
    assert (s.numpy() == g["b_out_segs"]).all() and (sc.numpy() == g["b_out_scores"]).all()
    assert (lb.numpy() == g["b_out_labels"]).all() and lb.dtype == torch.int64


def test_empty_input():
    from vilco_b200.utils import batched_nms
    s, sc, lb = batched_nms(torch.zeros(0, 2), torch.zeros(0), torch.zeros(0, dtype=torch.int64), 0.1, 1e-4, 200)
    assert s.shape == (0, 2) and sc.shape == (0,) and lb.shape == (0,) and lb.dtype == torch.int64


@pytest.mark.parametrize("n,K", [(20000, 22), (28000, 110), (6000, 1)])
def test_worst_case_sizes_vs_c_oracle(n, K):
    """untrained-weights candidate counts of the MQ config — whole batched_nms, bit-exact."""
    from oracle import mq_oracle as O
    from oracle import nms_c
    from vilco_b200.utils import batched_nms
    rs = np.random.RandomState(n + K)
    centre = rs.uniform(0, 1024, n).astype(np.float32)
    length = np.exp(rs.uniform(np.log(2.0), np.log(400.0), n)).astype(np.float32)
    segs = torch.from_numpy(np.stack([centre - length / 2, centre + length / 2], 1).astype(np.float32))
    scores = torch.from_numpy(rs.beta(0.5, 8, n).astype(np.float32))
    labels = torch.from_numpy(rs.randint(0, K, n).astype(np.int64))
    rs_, rsc, rl = O.batched_nms(segs, scores, labels, 0.1, 1e-4, 200, True, True, 0.99, softnms_fn=nms_c.softnms_1d)
    s, sc, lb = batched_nms(segs, scores, labels, 0.1, 1e-4, 200, True, True, 0.99)
    assert (s.numpy() == rs_.numpy()).all() and (sc.numpy() == rsc.numpy()).all() and (lb.numpy() == rl.numpy()).all()


def test_hard_nms_vs_oracle(g):
    from oracle import mq_oracle as O
    from vilco_b200.utils import batched_nms
    rs = np.random.RandomState(11)
    n, K = 4000, 5
    centre = rs.uniform(0, 1024, n).astype(np.float32)
    length = np.exp(rs.uniform(np.log(2.0), np.log(300.0), n)).astype(np.float32)
    segs = torch.from_numpy(np.stack([centre - length / 2, centre + length / 2], 1).astype(np.float32))
    scores = torch.from_numpy(rs.uniform(0, 1, n).astype(np.float32))
    labels = torch.from_numpy(rs.randint(0, K, n).astype(np.int64))
    a = O.batched_nms(segs, scores, labels, 0.4, 0.05, 100, use_soft_nms=False, multiclass=True)
    b = batched_nms(segs, scores, labels, 0.4, 0.05, 100, use_soft_nms=False, multiclass=True)
    for x, y in zip(a, b):
        assert (x.numpy() == y.numpy()).all()


def test_decode_vs_oracle():
    """decode kernel against oracle.decode_single_video on random head outputs (candidate sets and order)."""
    import ctypes as C
    from oracle import mq_oracle as O
    from vilco_b200 import engine as E, lib as L, ops
    cfg = O.ModelCfg()
    K, T = 22, 1024
    lens = [T >> l for l in range(10)]
    pyr = E.Pyramid(lens, "cuda")
    rs = np.random.RandomState(2)
    B = 2
    logits = torch.from_numpy(rs.normal(-4.0, 1.5, (B, pyr.P, K)).astype(np.float32)).cuda()
    offsets = torch.from_numpy(np.abs(rs.normal(1.0, 1.0, (B, pyr.P, 2))).astype(np.float32)).cuda()
    pmask = torch.zeros(B, pyr.P)
    valid = [1024, 611]
    for b in range(B):
        for l, (o, n) in enumerate(zip(pyr.off, pyr.lens)):
            pmask[b, o:o + n] = (torch.arange(n) * (1 << l) < valid[b]).float()
    pmask = pmask.cuda()
    nl, topk = 10, 5000
    cs = torch.empty(B, nl * topk, 2, device="cuda"); sc = torch.empty(B, nl * topk, device="cuda")
    lb = torch.empty(B, nl * topk, device="cuda", dtype=torch.int32); cnt = torch.zeros(B, nl, device="cuda", dtype=torch.int32)
    IntArr, FltArr = C.c_int * nl, C.c_float * nl
    L.check(L.lib().vilco_decode(ops._p(logits), ops._p(offsets), ops._p(pmask), B, pyr.P, K, nl, IntArr(*pyr.off),
                                 IntArr(*pyr.lens), FltArr(*[float(1 << l) for l in range(nl)]), C.c_float(0.001),
                                 C.c_float(0.01), topk, ops._p(cs), ops._p(sc), ops._p(lb), ops._p(cnt), L.stream_ptr()))
    pts = O.points(cfg, lens)
    for b in range(B):
        lv = [logits[b, o:o + n].cpu() for o, n in zip(pyr.off, pyr.lens)]
        ov = [offsets[b, o:o + n].cpu() for o, n in zip(pyr.off, pyr.lens)]
        mv = [pmask[b, o:o + n].cpu().bool() for o, n in zip(pyr.off, pyr.lens)]
        segs, scores, labels = O.decode_single_video(cfg, pts, mv, lv, ov)
        got_s = torch.cat([cs[b, l * topk:l * topk + int(cnt[b, l])] for l in range(nl)]).cpu()
        got_sc = torch.cat([sc[b, l * topk:l * topk + int(cnt[b, l])] for l in range(nl)]).cpu()
        got_l = torch.cat([lb[b, l * topk:l * topk + int(cnt[b, l])] for l in range(nl)]).cpu()
        assert got_s.shape == segs.shape
        # sigmoid differs in the last ulp between CPU and GPU: compare as sets keyed by (label, segment)
        ka = sorted(zip(labels.tolist(), [tuple(x) for x in segs.tolist()]))
        kb = sorted(zip(got_l.tolist(), [tuple(x) for x in got_s.tolist()]))
        same = sum(1 for x, y in zip(ka, kb) if x == y)
        assert same >= 0.999 * len(ka)
        assert abs(float(scores.sum()) - float(got_sc.sum())) < 1e-3 * float(scores.sum())


@pytest.mark.parametrize("name", ["a", "b"])
@pytest.mark.parametrize("soft", [True, False])
def test_class_agnostic_nms_with_segment_voting(name, soft):
    """multiclass=False + voting_thresh > 0 (nms.py:159-181): scores / labels / order identical to the reference, voted
    segments within fp32 rounding of its normalise-then-matmul formulation."""
    from vilco_b200.utils import batched_nms
    g = np.load(os.path.join(GOLDEN, "nms_voting.npz"))
    s, sc, lb = batched_nms(torch.from_numpy(g[name + "_in_segs"]), torch.from_numpy(g[name + "_in_scores"]),
                            torch.from_numpy(g[name + "_in_labels"]), 0.3, 1e-3, 100, use_soft_nms=soft, multiclass=False,
                            sigma=0.75, voting_thresh=0.75)
    tag = f"{name}_{'soft' if soft else 'hard'}"
    assert (sc.numpy() == g[tag + "_scores"]).all() and (lb.numpy() == g[tag + "_labels"]).all()
    assert np.abs(s.numpy() - g[tag + "_segs"]).max() <= 1e-5 * np.abs(g[tag + "_segs"]).max()


@pytest.mark.parametrize("B,K", [(32, 22), (40, 22), (3, 22)])
def test_many_clips_take_the_single_wave_cta_shapes(B, K):
    """the NMS launch picks its CTA shape from the grid (512 / 384 / 256 threads; shared-memory capacity 2048 / 1024 candidates
    per class, larger classes work in the global workspace): every shape must give what a clip evaluated alone gives — which the
    tests above pin bit-exactly to the reference extension.  One clip carries a class with more than 1024 candidates."""
    from vilco_b200.utils import batched_nms
    from vilco_b200.utils.nms import _run
    rs = np.random.RandomState(B)
    n = 6000
    segs = torch.zeros(B, n, 2)
    scores = torch.zeros(B, n)
    labels = torch.zeros(B, n, dtype=torch.int32)
    cnt = torch.zeros(B, 1, dtype=torch.int32)
    for b in range(B):
        m = n if b == 1 else int(rs.randint(200, 3000))
        c = rs.uniform(0, 1024, m).astype(np.float32)
        ln = np.exp(rs.uniform(np.log(2.0), np.log(300.0), m)).astype(np.float32)
        segs[b, :m] = torch.from_numpy(np.stack([c - ln / 2, c + ln / 2], 1))
        scores[b, :m] = torch.from_numpy(np.sort(rs.beta(0.5, 8, m).astype(np.float32))[::-1].copy())
        lab = rs.randint(0, K, m)
        if b == 1:
            lab[: m // 3] = 5                      # a 2000-candidate class: beyond the 1024-entry shared-memory shape
        labels[b, :m] = torch.from_numpy(lab.astype(np.int32))
        cnt[b, 0] = m
    out = _run(segs.cuda(), scores.cuda(), labels.cuda(), cnt.cuda(), B, 1, n, K, True, 2, 0.1, 0.99, 1e-4, 200)
    o_s, o_sc, o_lb, o_n = (t.cpu() for t in out)
    for b in range(B):
        m = int(cnt[b, 0])
        s, sc, lb = batched_nms(segs[b, :m], scores[b, :m], labels[b, :m].long(), 0.1, 1e-4, 200, True, True, 0.99)
        k = int(o_n[b])
        assert k == s.shape[0]
        assert (o_s[b, :k] == s).all() and (o_sc[b, :k] == sc).all() and (o_lb[b, :k] == lb).all(), b
