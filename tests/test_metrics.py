"""Evaluation tail: vilco_b200.utils.metrics / get_retrieval_performance against golden outputs of the
reference's own ANETdetection / Moment_Retrieval (tests/golden/metrics.json, made by oracle/gen_golden_metrics.py).
Host code only (the matcher is the C ABI's host entry vilco_ap_match), so everything here runs without a GPU."""
import json
import os
import pickle

import numpy as np
import pandas as pd
import pytest

from vilco_b200.utils import metrics as M
from vilco_b200.utils import get_retrieval_performance as R

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics.json")


@pytest.fixture(scope="module")
def golden():
    with open(GOLDEN) as f:
        return json.load(f)


def _preds(case):
    p = case["pred"]
    return {"video-id": [r[0] for r in p], "t-start": np.asarray([r[1] for r in p], np.float32),
            "t-end": np.asarray([r[2] for r in p], np.float32), "label": np.asarray([r[3] for r in p], np.int64),
            "score": np.asarray([r[4] for r in p], np.float32)}


def _annotation_json(case, path):
    db = {}
    for vid, t0, t1, lab in case["gt"]:
        e = db.setdefault(vid, dict(subset="val", clip_id=vid, annotations=[]))
        e["annotations"].append(dict(segment=[t0, t1], label_id=lab, label=f"class_{lab}"))
    with open(path, "w") as f:
        json.dump(dict(database=db), f)
    return db


def _cl_pickle(cases, path):
    tasks = []
    for case in cases:
        vids = {}
        for vid, t0, t1, lab in case["gt"]:
            e = vids.setdefault(vid, dict(id=vid, labels=[], segments=[]))
            e["labels"].append(lab)
            e["segments"].append([t0, t1])
        tasks.append(dict(dict_db=list(vids.values()), label_dict={f"class_{l}": l for l in case["label_ids"]}))
    with open(path, "wb") as f:
        pickle.dump(dict(val=tasks), f)


def _retrieval_json(case, path):
    res = {}
    for vid, t0, t1, lab, s in case["pred"]:
        res.setdefault(vid, []).append(dict(segment=[t0, t1], score=s, label=f"class_{lab}"))
    with open(path, "w") as f:
        json.dump(dict(version="1.0", external_data="", results=res), f)


def test_single_task_map_is_bit_identical_to_the_reference(golden, tmp_path):
    singles = [c for c in golden["cases"] if c["kind"] == "single"]
    assert len(singles) == 3
    for c in singles:
        ann = str(tmp_path / f"ann_{c['seed']}.json")
        _annotation_json(c["case"], ann)
        ev = M.ANETdetection(ann, "val", tiou_thresholds=np.asarray(golden["tious"]), num_workers=1)
        assert {str(k): int(v) for k, v in ev.activity_index.items()} == c["activity_index"]
        mAP, avg, thr = ev.evaluate(_preds(c["case"]), verbose=False)
        assert np.array_equal(ev.ap, np.asarray(c["ap"])), f"seed {c['seed']}: AP matrix differs"
        assert np.array_equal(mAP, np.asarray(c["mAP"])) and float(avg) == c["avg_mAP"]
        assert np.array_equal(thr, np.asarray(golden["tious"]))
        # the same through a DataFrame and through the per-label public function
        df = pd.DataFrame(_preds(c["case"]))
        df["t-start"], df["t-end"], df["score"] = (df[k].astype(np.float64) for k in ("t-start", "t-end", "score"))
        _, avg2, _ = ev.evaluate(df.copy(), verbose=False)
        assert float(avg2) == c["avg_mAP"]


def test_query_incremental_map_and_cumulative_ground_truth(golden, tmp_path):
    cl = [c for c in golden["cases"] if c["kind"] == "cl"]
    assert [c["task"] for c in cl] == [0, 1, 2]
    pk = str(tmp_path / "cl.pkl")
    _cl_pickle([c["own_case"] for c in cl], pk)
    ev = M.ANETdetection(pk, "val", tiou_thresholds=np.asarray(golden["tious"]), num_workers=1, use_cl=True)
    n_rows = 0
    for t, c in enumerate(cl):
        n_rows += len(c["own_case"]["gt"])
        assert len(ev.ground_truth[t]) == n_rows                      # cumulative over tasks 0..t (metrics.py:38-58)
        assert {str(k): int(v) for k, v in ev.activity_index[t].items()} == c["activity_index"]
        pred = sum((x["own_case"]["pred"] for x in cl[:t + 1]), [])
        mAP, avg, _ = ev.evaluate(_preds(dict(pred=pred)), current_task_id=t, verbose=False)
        assert np.array_equal(ev.ap, np.asarray(c["ap"])), f"task {t}: AP matrix differs"
        assert float(avg) == c["avg_mAP"]


def test_retrieval_recall_matches_the_reference(golden, tmp_path):
    for c in golden["cases"]:
        if c["kind"] == "single":
            db = _annotation_json(c["case"], str(tmp_path / "a.json"))
            flat = str(tmp_path / f"flat_{c['seed']}.json")
            with open(flat, "w") as f:
                json.dump(db, f)
            pj = str(tmp_path / f"pred_{c['seed']}.json")
            _retrieval_json(c["case"], pj)
            rec = R.evaluation_retrieval(gt=flat, pred=pj, subset="val", tiou=golden["tious"])
            assert np.array_equal(rec, np.asarray(c["recall"]))
            # without the json round trip: tables built from the arrays valid_one_epoch collects
            mr = R.Moment_Retrieval(flat, pj, subset="val")
            pred = R.predictions_from_results(_preds(c["case"]), idx_classes={l: f"class_{l}" for l in c["case"]["label_ids"]})
            rec2 = R.evaluation_retrieval(gt=mr.ground_truth, pred=pred, subset="val", tiou=golden["tious"])
            assert np.array_equal(rec2, np.asarray(c["recall"]))
    cl = [c for c in golden["cases"] if c["kind"] == "cl"]
    pk = str(tmp_path / "cl.pkl")
    _cl_pickle([c["own_case"] for c in cl], pk)
    for t, c in enumerate(cl):
        pj = str(tmp_path / f"pred_cl_{t}.json")
        _retrieval_json(c["own_case"], pj)
        rec = R.evaluation_retrieval(gt=pk, pred=pj, subset="val", tiou=golden["tious"], use_cl=True, current_task_id=t)
        assert np.array_equal(rec, np.asarray(c["recall"]))


def test_label_without_predictions_keeps_zero_ap(capsys):
    gt = pd.DataFrame({"video-id": ["a", "a", "b"], "t-start": [0., 20., 5.], "t-end": [10., 30., 9.], "label": [4, 9, 9]})
    index = {4: 0, 9: 1}
    gt["label"] = gt["label"].map(index)
    ev = M.ANETdetection((gt, index), tiou_thresholds=np.array([0.3, 0.5]))
    preds = {"video-id": ["a", "b", "zzz"], "t-start": np.float32([19., 5., 0.]), "t-end": np.float32([30., 9., 1.]),
             "label": np.int64([9, 9, 77]), "score": np.float32([0.9, 0.8, 0.7])}       # label 77: unknown to the ground truth
    mAP, avg, _ = ev.evaluate(preds, verbose=True)
    out = capsys.readouterr().out
    assert "No predictions of label '4'" in out and "Avearge mAP" in out
    assert ev.ap[:, 0].tolist() == [0.0, 0.0] and ev.ap[:, 1].tolist() == [1.0, 1.0]
    assert mAP.tolist() == [0.5, 0.5] and avg == 0.5


def _python_matcher(pred_seg, pred_vid, gt_seg, gt_start, thr):
    """the reference's loop (metrics.py:302-331) over plain arrays — independent of the C implementation."""
    tp = np.zeros((len(thr), len(pred_vid)), np.uint8)
    lock = -np.ones((len(thr), len(gt_seg)))
    for i in range(len(pred_vid)):
        v = pred_vid[i]
        if v < 0:
            continue
        g0, g1 = gt_start[v], gt_start[v + 1]
        tiou = M.segment_iou(pred_seg[i], gt_seg[g0:g1])
        order = np.argsort(tiou, kind="stable")[::-1]
        for t, th in enumerate(thr):
            for j in order:
                if tiou[j] < th:
                    break
                if lock[t, g0 + j] >= 0:
                    continue
                tp[t, i] = 1
                lock[t, g0 + j] = i
                break
    return tp


def test_matcher_edge_cases_and_random_agreement():
    thr = np.linspace(0.1, 0.5, 5)
    # no predictions / no ground truth
    assert M.ap_match(np.zeros((0, 2)), np.zeros(0, np.int64), np.zeros((0, 2)), np.zeros(1, np.int64), thr).shape == (5, 0)
    tp = M.ap_match(np.array([[0., 1.]]), np.array([-1]), np.zeros((0, 2)), np.zeros(1, np.int64), thr)
    assert tp.shape == (5, 1) and tp.sum() == 0
    # the second prediction of the same moment is a false positive; a looser threshold lets it take the neighbour
    gt = np.array([[0., 10.], [8., 20.]])
    pr = np.array([[0., 10.], [1., 10.], [30., 40.]])
    tp = M.ap_match(pr, np.array([0, 0, 0]), gt, np.array([0, 2]), np.array([0.1, 0.5]))
    assert tp.tolist() == [[1, 1, 0], [1, 0, 0]]
    # zero-length prediction and ground truth anywhere: tIoU = 0/0 = NaN, which the reference counts as a match
    tp = M.ap_match(np.array([[5., 5.]]), np.array([0]), np.array([[9., 9.]]), np.array([0, 1]), thr)
    assert tp.sum() == 5
    # tied tIoUs (the same moment annotated three times): each repeat of the prediction takes one of them, the fourth is a
    # false positive; among equal tIoUs the higher ground-truth row is taken first (argsort()[::-1] of <= 16 rows)
    gt3 = np.array([[0., 10.], [0., 10.], [0., 10.], [50., 60.]])
    pr4 = np.array([[0., 10.]] * 4)
    tp = M.ap_match(pr4, np.zeros(4, np.int64), gt3, np.array([0, 4]), np.array([0.5]))
    assert tp.tolist() == [[1, 1, 1, 0]]
    assert np.array_equal(tp, _python_matcher(pr4, np.zeros(4, np.int64), gt3, np.array([0, 4]), np.array([0.5])))
    # bad arguments are reported through the error code, not a crash
    with pytest.raises(Exception, match="out of range"):
        M.ap_match(np.array([[0., 1.]]), np.array([3]), np.array([[0., 1.]]), np.array([0, 1]), thr)
    rng = np.random.default_rng(0)
    for trial in range(20):
        n_vid, n_pred = int(rng.integers(1, 6)), int(rng.integers(1, 80))
        counts = rng.integers(0, 7, n_vid)
        gt_start = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        g0 = rng.uniform(0, 50, int(counts.sum()))
        gt = np.stack([g0, g0 + rng.uniform(1, 30, len(g0))], 1)
        p0 = rng.uniform(0, 50, n_pred)
        pr = np.stack([p0, p0 + rng.uniform(1, 30, n_pred)], 1)
        vid = rng.integers(-1, n_vid, n_pred).astype(np.int64)
        assert np.array_equal(M.ap_match(pr, vid, gt, gt_start, thr), _python_matcher(pr, vid, gt, gt_start, thr)), trial


def test_interpolated_ap_and_iou_helpers():
    rng = np.random.default_rng(1)
    for _ in range(10):
        n = int(rng.integers(1, 50))
        tp = (rng.random(n) < 0.4).astype(float)
        prec = np.cumsum(tp) / np.arange(1, n + 1)
        rec = np.cumsum(tp) / max(1.0, tp.sum() + 2)
        mprec = np.hstack([[0], prec, [0]])
        mrec = np.hstack([[0], rec, [1]])
        for i in range(len(mprec) - 1)[::-1]:                         # the reference's loop form (metrics.py:379-384)
            mprec[i] = max(mprec[i], mprec[i + 1])
        idx = np.where(mrec[1::] != mrec[0:-1])[0] + 1
        assert M.interpolated_prec_rec(prec, rec) == np.sum((mrec[idx] - mrec[idx - 1]) * mprec[idx])
    assert np.allclose(M.segment_iou(np.array([0., 10.]), np.array([[5., 15.], [20., 30.]])), [1 / 3, 0.0])
    assert np.allclose(R.iou([[0., 10.]], [[5., 15.], [20., 30.]]), [[1 / 3, 0.0]])
    ants = [dict(segment=[0., 1.], label_id=1), dict(segment=[0.0005, 1.0005], label_id=1), dict(segment=[0., 1.], label_id=2)]
    assert len(M.remove_duplicate_annotations(ants)) == 2
