"""Detection mAP evaluator — host-side mirror of `MQ/libs/utils/metrics.py` (the evaluation tail).

Same names, constructor arguments and return values as the reference (`ANETdetection(ant_file, split, tiou_thresholds,
label, label_offset, num_workers, dataset_name, debug_video_id, use_cl)`, `.evaluate(preds, current_task_id, verbose)` →
`(mAP, average_mAP, tiou_thresholds)`, `.ap`, `.ground_truth`, `.activity_index`), so `eval.py:103-110`, `train_cl.py:163-170`
and `valid_one_epoch*` (`train_utils.py:1148`) use it unchanged.  What differs is how the work is done:

* the reference walks every prediction as a pandas row (`iterrows`, one `groupby.get_group` + `reset_index` per prediction)
  inside joblib workers, one per label (`metrics.py:216-222, 302-331`);
* here predictions and ground truth are plain arrays, grouped once per label, and the greedy matching runs in the C ABI
  entry `vilco_ap_match` (`include/vilco_b200.h`, `csrc/eval_host.cu`; host code, float64, the reference's operation
  order).  Sorting and the precision/recall integration stay in numpy with the reference's own calls
  (`argsort()[::-1]`, `cumsum`, `np.sum`), so the AP values are bit-identical (`tests/test_metrics.py`).

Two library-version quirks of the reference are pinned to its own environment (numpy 1.20 / pandas 1.x): `np.float` is
float64, and in query-incremental mode `preds['label'].replace(<list of dicts>)` (`metrics.py:257`) is a no-op, i.e.
predicted labels are compared *as given* with the dense ground-truth label index of the task.
"""
import ctypes as C
import json
import os
import pickle as pkl
from typing import Dict

import numpy as np
import pandas as pd

from .. import lib as L


def remove_duplicate_annotations(ants, tol=1e-3):
    """metrics.py:14-30 — keep the first of annotations that agree in label and (within tol) in both end points."""
    kept = []
    for ev in ants:
        s, e, lab = ev['segment'][0], ev['segment'][1], ev['label_id']
        if not any(abs(s - k['segment'][0]) <= tol and abs(e - k['segment'][1]) <= tol and lab == k['label_id']
                   for k in kept):
            kept.append(ev)
    return kept


def _label_id(value, label_offset):
    if isinstance(value, (tuple, list)):
        # metrics.py:93-97 (sic: `label_offset**i + int(x)`, kept as the reference computes it)
        return sum(label_offset ** i + int(x) for i, x in enumerate(value[::-1]))
    return int(value)


def _dense_index(frame):
    """sorted unique labels -> 0..n-1, applied in place (metrics.py:55-56, 111-112)."""
    index = {j: i for i, j in enumerate(sorted(frame['label'].unique()))}
    frame['label'] = frame['label'].map(index).astype(np.int64) if len(frame) else frame['label']
    return index


def load_gt_seg_from_json(json_file, split=None, label='label_id', label_offset=0, debug_video_id=None, use_cl=False):
    """metrics.py:33-115.  use_cl: `json_file` is the query-incremental pickle, one CUMULATIVE ground truth per task (the
    reference never resets its row lists between tasks, :41-54) with that task's own dense label index."""
    if use_cl:
        with open(json_file, 'rb') as f:
            tasks = pkl.load(f)['val']
        vids, t0, t1, labs = [], [], [], []
        ground_truth, activity_index = [], []
        for task in tasks:
            for video in task['dict_db']:
                for i, lab in enumerate(video['labels']):
                    vids.append(video['id'])
                    t0.append(float(video['segments'][i][0]))
                    t1.append(float(video['segments'][i][1]))
                    labs.append(lab)
            frame = pd.DataFrame({'video-id': list(vids), 't-start': list(t0), 't-end': list(t1), 'label': list(labs)})
            activity_index.append(_dense_index(frame))
            ground_truth.append(frame)
        return ground_truth, activity_index

    with open(json_file, 'r', encoding='utf8') as f:
        db = json.load(f)
    db = db.get('database', db)
    vids, t0, t1, labs = [], [], [], []
    for key, v in db.items():
        if debug_video_id is not None and v['clip_id'] != debug_video_id[-1]:
            continue
        if split is not None and v['subset'].lower() != split:
            continue
        for ev in remove_duplicate_annotations(v['annotations']):
            vids.append(key)
            t0.append(float(ev['segment'][0]))
            t1.append(float(ev['segment'][1]))
            labs.append(_label_id(ev[label], label_offset))
    frame = pd.DataFrame({'video-id': vids, 't-start': t0, 't-end': t1, 'label': labs})
    return frame, _dense_index(frame)


def load_pred_seg_from_json(json_file, label='label_id', label_offset=0):
    """metrics.py:115-148."""
    with open(json_file, 'r', encoding='utf8') as f:
        db = json.load(f)['database']
    vids, t0, t1, labs, scores = [], [], [], [], []
    for key, events in db.items():
        for ev in events:
            vids.append(key)
            t0.append(float(ev['segment'][0]))
            t1.append(float(ev['segment'][1]))
            labs.append(_label_id(ev[label], label_offset))
            scores.append(float(ev['scores']))
    return pd.DataFrame({'video-id': vids, 't-start': t0, 't-end': t1, 'label': labs, 'score': scores})


def segment_iou(target_segment, candidate_segments):
    """metrics.py:348-372 (numpy form, float64; 0/0 stays NaN)."""
    target_segment = np.asarray(target_segment, np.float64)
    candidate_segments = np.asarray(candidate_segments, np.float64)
    lo = np.maximum(target_segment[0], candidate_segments[:, 0])
    hi = np.minimum(target_segment[1], candidate_segments[:, 1])
    inter = (hi - lo).clip(0)
    union = (candidate_segments[:, 1] - candidate_segments[:, 0]) + (target_segment[1] - target_segment[0]) - inter
    with np.errstate(invalid='ignore', divide='ignore'):
        return inter.astype(float) / union


def interpolated_prec_rec(prec, rec):
    """metrics.py:375-384 — VOC-2011 interpolated AP; the running maximum from the right is one accumulate."""
    mprec = np.hstack([[0], prec, [0]])
    mrec = np.hstack([[0], rec, [1]])
    mprec = np.maximum.accumulate(mprec[::-1])[::-1]
    idx = np.where(mrec[1::] != mrec[0:-1])[0] + 1
    return np.sum((mrec[idx] - mrec[idx - 1]) * mprec[idx])


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def ap_match(pred_seg, pred_vid, gt_seg, gt_start, tiou_thresholds):
    """tp (n_thr, n_pred) uint8 through the C ABI (`vilco_ap_match`).  All arguments are host arrays."""
    pred_seg = np.ascontiguousarray(pred_seg, np.float64).reshape(-1, 2)
    pred_vid = np.ascontiguousarray(pred_vid, np.int64)
    gt_seg = np.ascontiguousarray(gt_seg, np.float64).reshape(-1, 2)
    gt_start = np.ascontiguousarray(gt_start, np.int64)
    thr = np.ascontiguousarray(tiou_thresholds, np.float64)
    tp = np.zeros((len(thr), len(pred_vid)), np.uint8)
    rc = L.lib().vilco_ap_match(_ptr(pred_seg, C.c_double), _ptr(pred_vid, C.c_int64), C.c_int64(len(pred_vid)),
                                _ptr(gt_seg, C.c_double), _ptr(gt_start, C.c_int64), C.c_int64(len(gt_start) - 1),
                                _ptr(thr, C.c_double), C.c_int(len(thr)), _ptr(tp, C.c_uint8))
    L.check(rc, "vilco_ap_match")
    return tp


def _average_precision(gt_vid, gt_t0, gt_t1, pr_vid, pr_t0, pr_t1, pr_score, tiou_thresholds):
    """AP per tIoU threshold for the ground truth / predictions of one label (array form of metrics.py:277-346)."""
    ap = np.zeros(len(tiou_thresholds))
    if len(pr_score) == 0:
        return ap
    npos = float(len(gt_vid))
    order = np.asarray(pr_score).argsort()[::-1]                     # :297 (same call: same order on ties)
    # ground truth grouped by video, original row order inside a video (:303, 314)
    codes, uniques = pd.factorize(np.asarray(gt_vid, dtype=object))
    by_video = np.argsort(codes, kind='stable')
    gt_start = np.zeros(len(uniques) + 1, np.int64)
    np.cumsum(np.bincount(codes, minlength=len(uniques)), out=gt_start[1:])
    gt_seg = np.stack([np.asarray(gt_t0, np.float64)[by_video], np.asarray(gt_t1, np.float64)[by_video]], 1)
    pred_vid = pd.Index(uniques).get_indexer(np.asarray(pr_vid, dtype=object)[order]) if len(uniques) else \
        np.full(len(order), -1, np.int64)
    pred_seg = np.stack([np.asarray(pr_t0, np.float64)[order], np.asarray(pr_t1, np.float64)[order]], 1)
    tp = ap_match(pred_seg, pred_vid, gt_seg, gt_start, tiou_thresholds).astype(np.float64)
    fp = 1.0 - tp                                                    # every prediction is exactly one of the two (:328-331)
    tp_cumsum = np.cumsum(tp, axis=1)
    fp_cumsum = np.cumsum(fp, axis=1)
    recall_cumsum = tp_cumsum / npos
    precision_cumsum = tp_cumsum / (tp_cumsum + fp_cumsum)
    for t in range(len(tiou_thresholds)):
        ap[t] = interpolated_prec_rec(precision_cumsum[t, :], recall_cumsum[t, :])
    return ap


def compute_average_precision_detection(ground_truth, prediction, tiou_thresholds=np.linspace(0.1, 0.5, 5)):
    """metrics.py:277-346 with the reference's DataFrame arguments (`video-id`, `t-start`, `t-end` [, `score`])."""
    if prediction.empty:
        return np.zeros(len(tiou_thresholds))
    return _average_precision(ground_truth['video-id'].values, ground_truth['t-start'].values, ground_truth['t-end'].values,
                              prediction['video-id'].values, prediction['t-start'].values, prediction['t-end'].values,
                              prediction['score'].values, tiou_thresholds)


class ANETdetection(object):
    """Drop-in for `libs.utils.ANETdetection` (metrics.py:151-275).  `ant_file` may also be an already loaded
    `(ground_truth, activity_index)` pair (what `load_gt_seg_from_json` returns), so a caller that holds the annotations in
    memory skips the file."""

    def __init__(self, ant_file, split=None, tiou_thresholds=np.linspace(0.1, 0.5, 5), label='label_id', label_offset=0,
                 num_workers=8, dataset_name=None, debug_video_id=None, use_cl=False):
        self.tiou_thresholds = tiou_thresholds
        self.ap = None
        self.num_workers = num_workers            # kept for the signature; the matcher needs no worker pool
        self.use_cl = use_cl
        self.split = split
        if isinstance(ant_file, (tuple, list)):
            self.dataset_name = dataset_name if dataset_name is not None else 'in-memory'
            ground_truth, activity_index = ant_file
        else:
            self.dataset_name = dataset_name if dataset_name is not None else \
                os.path.basename(ant_file).replace('.json', '')
            ground_truth, activity_index = load_gt_seg_from_json(
                ant_file, split=self.split, label=label, label_offset=label_offset, debug_video_id=debug_video_id,
                use_cl=self.use_cl)
        self.ground_truth = ground_truth
        self.activity_index = activity_index

    def wrapper_compute_average_precision(self, preds, current_task_id=None):
        """AP (n_thr, n_labels) — metrics.py:200-228.  One pass groups both tables by label; labels without predictions
        keep AP 0 (the reference prints a warning for them, :190-198)."""
        if self.use_cl:
            ground_truth = self.ground_truth[current_task_id]
            activity_index = self.activity_index[current_task_id]
        else:
            ground_truth, activity_index = self.ground_truth, self.activity_index
        ap = np.zeros((len(self.tiou_thresholds), len(activity_index)))
        gt_rows = ground_truth.groupby('label').indices
        pr_rows = preds.groupby('label').indices if len(preds) else {}
        g_vid, g_t0, g_t1 = (ground_truth[c].values for c in ('video-id', 't-start', 't-end'))
        p_vid, p_t0, p_t1, p_sc = (preds[c].values for c in ('video-id', 't-start', 't-end', 'score'))
        for label_name, cidx in activity_index.items():
            g = gt_rows[cidx]                     # KeyError like the reference's get_group when the label has no rows
            p = pr_rows.get(cidx)
            if p is None:
                print('Warning: No predictions of label \'%s\' were provdied.' % label_name)
                continue
            ap[:, cidx] = _average_precision(g_vid[g], g_t0[g], g_t1[g], p_vid[p], p_t0[p], p_t1[p], p_sc[p],
                                             self.tiou_thresholds)
        return ap

    def evaluate(self, preds, current_task_id=None, verbose=True):
        """metrics.py:230-275.  preds: DataFrame, path of a prediction json, or the dict of arrays / tensors built by
        `valid_one_epoch*` (`video-id`, `t-start`, `t-end`, `label`, `score`)."""
        if isinstance(preds, pd.DataFrame):
            assert 'label' in preds
        elif isinstance(preds, str) and os.path.isfile(preds):
            preds = load_pred_seg_from_json(preds)
        elif isinstance(preds, Dict):
            preds = pd.DataFrame({
                'video-id': preds['video-id'],
                't-start': np.asarray(preds['t-start']).astype(np.float64),     # == .tolist(): float32 widened exactly
                't-end': np.asarray(preds['t-end']).astype(np.float64),
                'label': np.asarray(preds['label']),
                'score': np.asarray(preds['score']).astype(np.float64)})
        self.ap = None
        if isinstance(self.activity_index, dict):
            # original label ids -> dense ids; ids the ground truth does not know stay as they are (Series.replace)
            lab = preds['label']
            preds['label'] = lab.map(self.activity_index).where(lab.isin(list(self.activity_index)), lab).astype(lab.dtype)
        # else (use_cl): see the module docstring — the reference's replace(list) leaves the labels untouched
        self.ap = self.wrapper_compute_average_precision(preds, current_task_id)
        mAP = self.ap.mean(axis=1)
        average_mAP = mAP.mean()
        if verbose:
            print('[RESULTS] Action detection results on {:s}.'.format(self.dataset_name))
            print(''.join('\n|tIoU = {:.2f}: mAP = {:.2f} (%)'.format(t, m * 100)
                          for t, m in zip(self.tiou_thresholds, mAP)))
            print('Avearge mAP: {:.2f} (%)'.format(average_mAP * 100))
        return mAP, average_mAP, self.tiou_thresholds
