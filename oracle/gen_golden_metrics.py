"""TEST INFRASTRUCTURE ONLY — golden vectors for the evaluation tail, produced by running the REFERENCE's
own `MQ/libs/utils/metrics.py` (ANETdetection) and `MQ/libs/utils/get_retrieval_performance.py` (Moment_Retrieval) on
seeded synthetic predictions / annotations.

Run in the authoring container only:   python -m oracle.gen_golden_metrics
Writes tests/golden/metrics.json (inputs + the reference's outputs).  The GPU box never runs this.

The two reference files are loaded by path (importing the `libs.utils` package would pull the whole training stack).
Two shims for the libraries of this container (the reference pins numpy 1.20.3 and an unpinned pandas of that era):
* `metrics.py:334` uses `np.float`, removed from numpy >= 1.24: aliased to `float` (same dtype, float64);
* in query-incremental mode `evaluate` calls `preds['label'].replace(self.activity_index)` with a LIST of dicts
  (metrics.py:257, activity_index is a list when use_cl=True).  pandas 1.x treats a list `to_replace` without `value` as
  "forward-fill the entries equal to a list element"; no label equals a dict, so it is a no-op there, while pandas >= 2
  raises.  `Series.replace` is wrapped to restore the 1.x behaviour for exactly that call shape.
"""
import importlib.util
import io
import json
import os
import pickle
import sys
import tempfile
from contextlib import redirect_stdout

import numpy as np

REF = "/root/reference/MQ/libs/utils"
GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _load(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_" + name] = mod          # joblib pickles functions by module name
    spec.loader.exec_module(mod)
    return mod


def make_case(seed, n_videos, n_labels, n_pred_per_video, label_ids=None):
    """One synthetic evaluation set.  Covers: several ground truths of a label in one video, exact duplicate ground
    truths (tIoU ties), labels nobody predicts, predictions in videos without ground truth of that label, a predicted
    video that has no annotation at all, zero-length predictions and ground truths (the 0/0 = NaN tIoU of segment_iou)."""
    rng = np.random.default_rng(seed)
    label_ids = list(label_ids) if label_ids is not None else list(range(0, 3 * n_labels, 3))[:n_labels]
    gt = []                                   # (video, t0, t1, label)
    for v in range(n_videos):
        vid = f"clip_{seed}_{v:03d}"
        dur = float(rng.uniform(60.0, 480.0))
        for _ in range(int(rng.integers(1, 9))):
            c = float(rng.uniform(0, dur))
            ln = float(np.exp(rng.uniform(np.log(1.0), np.log(120.0))))
            t0, t1 = max(0.0, c - ln / 2), min(dur, c + ln / 2)
            lab = label_ids[int(rng.integers(0, max(1, n_labels - 1)))]     # the last label never occurs in predictions' gt
            gt.append((vid, round(t0, 3), round(t1, 3), lab))
            if rng.random() < 0.15:           # the same moment annotated twice, a little apart (> 1e-3: not a duplicate)
                gt.append((vid, round(t0, 3) + 0.25, round(t1, 3) + 0.25, lab))
            if rng.random() < 0.05:           # exact duplicate rows (kept by the CL loader, dropped by the json loader)
                gt.append(gt[-1])
        if v == 1:
            gt.append((vid, 7.5, 7.5, label_ids[0]))                         # zero-length ground truth
    pred = []                                 # (video, t0, t1, label, score) in per-video descending score order
    scores_seen = set()
    for v in range(n_videos + 1):             # the extra video has predictions but no annotation
        vid = f"clip_{seed}_{v:03d}"
        rows = []
        g_here = [g for g in gt if g[0] == vid]
        for i in range(n_pred_per_video):
            if g_here and rng.random() < 0.5:  # jittered copy of a ground truth: produces matches at every threshold
                g = g_here[int(rng.integers(0, len(g_here)))]
                ln = g[2] - g[1]
                t0 = g[1] + float(rng.normal(0, 0.2)) * ln
                t1 = g[2] + float(rng.normal(0, 0.2)) * ln
                lab = g[3] if rng.random() < 0.8 else label_ids[int(rng.integers(0, n_labels))]
            else:
                t0 = float(rng.uniform(0, 400.0))
                t1 = t0 + float(rng.uniform(0.5, 60.0))
                lab = label_ids[int(rng.integers(0, n_labels))]
            if t1 < t0:
                t0, t1 = t1, t0
            if v == 1 and i < 3:
                t1 = t0                        # zero-length predictions
                lab = label_ids[0]
            while True:                        # distinct float32 scores: the sort order must not depend on numpy's tie order
                s = np.float32(rng.beta(0.5, 4.0))
                if float(s) not in scores_seen and s > 0:
                    scores_seen.add(float(s))
                    break
            rows.append((vid, float(np.float32(max(t0, 0.0))), float(np.float32(t1)), lab, float(s)))
        rows.sort(key=lambda r: -r[4])
        pred.extend(rows)
    return dict(label_ids=label_ids, gt=gt, pred=pred)


def write_annotation_json(case, path):
    """ActivityNet-style annotation file as read by load_gt_seg_from_json (metrics.py:60-115) and by
    Moment_Retrieval._import_ground_truth (get_retrieval_performance.py:72-91)."""
    db = {}
    for vid, t0, t1, lab in case["gt"]:
        e = db.setdefault(vid, dict(subset="val", clip_id=vid, annotations=[]))
        e["annotations"].append(dict(segment=[t0, t1], label_id=lab, label=f"class_{lab}"))
    with open(path, "w") as f:
        json.dump(dict(database=db), f)


def write_cl_pickle(cases, path):
    """query-incremental annotation pickle: data['val'][task] = {dict_db: [{id, labels, segments}], label_dict}."""
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


def preds_dict(case):
    p = case["pred"]
    return {"video-id": [r[0] for r in p],
            "t-start": np.asarray([r[1] for r in p], np.float32),
            "t-end": np.asarray([r[2] for r in p], np.float32),
            "label": np.asarray([r[3] for r in p], np.int64),
            "score": np.asarray([r[4] for r in p], np.float32)}


def write_retrieval_json(case, path):
    res = {}
    for vid, t0, t1, lab, s in case["pred"]:
        res.setdefault(vid, []).append(dict(segment=[t0, t1], score=s, label=f"class_{lab}"))
    with open(path, "w") as f:
        json.dump(dict(version="1.0", external_data="", results=res), f)


def main():
    np.float = float                          # metrics.py:334-335 (see module docstring)
    import pandas as pd
    _replace = pd.Series.replace

    def replace_1x(self, to_replace=None, *a, **k):
        if isinstance(to_replace, list) and not a and "value" not in k:
            return self.copy()
        return _replace(self, to_replace, *a, **k)
    pd.Series.replace = replace_1x
    M = _load("metrics")
    R = _load("get_retrieval_performance")
    tious = [0.1, 0.2, 0.3, 0.4, 0.5]
    out = dict(tious=tious, cases=[])
    tmp = tempfile.mkdtemp()

    # --- single-task evaluation (eval.py:103-110): json annotations
    for seed, nv, nl, npv in ((0, 12, 6, 40), (1, 30, 22, 200), (2, 3, 4, 5)):
        case = make_case(seed, nv, nl, npv)
        ann = os.path.join(tmp, f"ann_{seed}.json")
        write_annotation_json(case, ann)
        ev = M.ANETdetection(ann, "val", tiou_thresholds=np.asarray(tious), num_workers=1)
        with redirect_stdout(io.StringIO()):
            mAP, avg, _ = ev.evaluate(preds_dict(case), verbose=False)
        # the retrieval file lists every annotated video; drop the annotation-free extra video's absence problem by
        # construction (it has predictions, so `key_v in self.prediction` always holds)
        pj = os.path.join(tmp, f"pred_{seed}.json")
        write_retrieval_json(case, pj)
        # non-CL Moment_Retrieval reads a flat {video: {...}} json (no 'database' level)
        flat = os.path.join(tmp, f"flat_{seed}.json")
        with open(ann) as f:
            json.dump(json.load(f)["database"], open(flat, "w"))
        with redirect_stdout(io.StringIO()):
            rec = R.evaluation_retrieval(gt=flat, pred=pj, subset="val", tiou=tious)
        out["cases"].append(dict(kind="single", seed=seed, case=case, ap=ev.ap.tolist(), mAP=mAP.tolist(), avg_mAP=float(avg),
                                 activity_index={str(k): int(v) for k, v in ev.activity_index.items()},
                                 recall=np.asarray(rec).tolist()))
        print("single", seed, "avg mAP", avg, "R1@0.3", rec[2][0], "R5@0.5", rec[4][1])

    # --- query-incremental evaluation (train_cl.py:163-170, use_cl=True): one cumulative ground truth per task
    cl_cases = [make_case(10 + t, 8, 4, 30, label_ids=range(4 * t, 4 * t + 4)) for t in range(3)]
    pk = os.path.join(tmp, "cl.pkl")
    write_cl_pickle(cl_cases, pk)
    ev = M.ANETdetection(pk, "val", tiou_thresholds=np.asarray(tious), num_workers=1, use_cl=True)
    for t in range(3):
        # predictions of the videos of tasks 0..t (the ground truth of task t is cumulative, metrics.py:38-58)
        merged = dict(label_ids=sum((list(c["label_ids"]) for c in cl_cases[:t + 1]), []),
                      gt=sum((c["gt"] for c in cl_cases[:t + 1]), []), pred=sum((c["pred"] for c in cl_cases[:t + 1]), []))
        try:
            with redirect_stdout(io.StringIO()):
                mAP, avg, _ = ev.evaluate(preds_dict(merged), current_task_id=t, verbose=False)
            ap = ev.ap.tolist()
            err = None
        except Exception as e:                # recorded as is: the mirror must raise / behave the same way
            mAP, avg, ap, err = None, None, None, f"{type(e).__name__}: {e}"
        pj = os.path.join(tmp, f"pred_cl_{t}.json")
        write_retrieval_json(cl_cases[t], pj)
        with redirect_stdout(io.StringIO()):
            rec = R.evaluation_retrieval(gt=pk, pred=pj, subset="val", tiou=tious, use_cl=True, current_task_id=t)
        out["cases"].append(dict(kind="cl", task=t, own_case=cl_cases[t], ap=ap,
                                 mAP=None if mAP is None else mAP.tolist(), avg_mAP=None if avg is None else float(avg),
                                 error=err, recall=np.asarray(rec).tolist(),
                                 activity_index={str(k): int(v) for k, v in ev.activity_index[t].items()}))
        print("cl task", t, "avg mAP", avg, "err", err, "R1@0.3", rec[2][0])

    with open(os.path.join(GOLDEN, "metrics.json"), "w") as f:
        json.dump(out, f)
    print("wrote", os.path.join(GOLDEN, "metrics.json"), os.path.getsize(os.path.join(GOLDEN, "metrics.json")), "bytes")


if __name__ == "__main__":
    main()
