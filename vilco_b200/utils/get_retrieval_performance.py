"""Recall@k×(number of ground truths) of the moment-retrieval evaluation — host-side mirror of
`MQ/libs/utils/get_retrieval_performance.py`, called by `valid_one_epoch_cl_single_gpu`
(`train_utils.py:1117-1135`) after every validation pass.

Same class / function names and arguments as the reference.  In addition the ground truth and the predictions may be
passed as already loaded Python objects (what `_import_ground_truth` / `_import_prediction` return), and
`predictions_from_results` builds the prediction table straight from the arrays `valid_one_epoch*` collects — the
reference writes them to `retrieval_json/results_*.json` one `.item()` at a time and parses the file again
(`train_utils.py:1109-1128`).
"""
import json
import pickle as pkl

import numpy as np

TIOUS = (0.1, 0.2, 0.3, 0.4, 0.5)        # hard-coded in the reference's evaluate() (:118-119); the ctor value is not used
RECALLS = (1, 5)


def iou(pred, gt):
    """get_retrieval_performance.py:164-184 — intersection over the HULL of the two segments (max right − min left), for
    all (prediction, ground truth) pairs.  Lists of [t0, t1(, score)]; a flat segment is treated as one row."""
    assert isinstance(pred, list) and isinstance(gt, list)
    pred_is_list, gt_is_list = isinstance(pred[0], list), isinstance(gt[0], list)
    p = np.array(pred if pred_is_list else [pred])
    g = np.array(gt if gt_is_list else [gt])
    inter = np.maximum(0.0, np.minimum(p[:, 1, None], g[None, :, 1]) - np.maximum(p[:, 0, None], g[None, :, 0]))
    hull = np.maximum(0.0, np.maximum(p[:, 1, None], g[None, :, 1]) - np.minimum(p[:, 0, None], g[None, :, 0]))
    with np.errstate(invalid='ignore', divide='ignore'):
        overlap = 1.0 * inter / hull
    if not gt_is_list:
        overlap = overlap[:, 0]
    if not pred_is_list:
        overlap = overlap[0]
    return overlap


def predictions_from_results(results, idx_classes=None):
    """{video: {label: [[t0, t1, score], ...]}} from the `results` dict of `valid_one_epoch*` (`video-id` list and
    `t-start` / `t-end` / `label` / `score` arrays, rows of a video in the model's output order).  `idx_classes` maps the
    integer label to the ground truth's label key (the reference's `idx_classes`, train_utils.py:1100-1106)."""
    t0 = np.asarray(results['t-start'], np.float32).astype(np.float64)      # == tensor.item() of a float32
    t1 = np.asarray(results['t-end'], np.float32).astype(np.float64)
    sc = np.asarray(results['score'], np.float32).astype(np.float64)
    lab = np.asarray(results['label'])
    out = {}
    for i, vid in enumerate(results['video-id']):
        key = int(lab[i]) if idx_classes is None else idx_classes[int(lab[i])]
        out.setdefault(vid, {}).setdefault(key, []).append([t0[i], t1[i], sc[i]])
    return out


class Moment_Retrieval(object):
    GROUND_TRUTH_FIELDS = ['database']
    PREDICTION_FIELDS = ['results', 'version', 'external_data']

    def __init__(self, ground_truth_filename=None, prediction_filename=None, ground_truth_fields=GROUND_TRUTH_FIELDS,
                 prediction_fields=PREDICTION_FIELDS, tiou_thresholds=np.linspace(0.5, 0.95, 10), subset='test',
                 verbose=False, check_status=False, use_cl=False):
        if ground_truth_filename is None or (isinstance(ground_truth_filename, str) and not ground_truth_filename):
            raise IOError('Please input a valid ground truth file.')
        if prediction_filename is None or (isinstance(prediction_filename, str) and not prediction_filename):
            raise IOError('Please input a valid prediction file.')
        self.subset = subset
        self.tiou_thresholds = tiou_thresholds
        self.verbose = verbose
        self.gt_fields = ground_truth_fields
        self.pred_fields = prediction_fields
        self.ap = None
        self.check_status = check_status
        self.use_cl = use_cl
        self.ground_truth = ground_truth_filename if not isinstance(ground_truth_filename, str) else \
            self._import_ground_truth(ground_truth_filename)
        self.prediction = prediction_filename if not isinstance(prediction_filename, str) else \
            self._import_prediction(prediction_filename)
        if self.verbose:
            n_gt = sum(len(g) for g in self.ground_truth) if self.use_cl else len(self.ground_truth)
            print('[INIT] Loaded annotations from {} subset.'.format(subset))
            print('\tNumber of ground truth instances: {}'.format(n_gt))
            print('\tNumber of predictions: {}'.format(len(self.prediction)))
            print('\tFixed threshold for tiou score: {}'.format(self.tiou_thresholds))

    def _import_ground_truth(self, ground_truth_filename):
        """:46-91 → {video: {label: [[t0, t1], ...]}} (use_cl: one such dict per task, labels by name)."""
        if self.use_cl:
            with open(ground_truth_filename, 'rb') as f:
                tasks = pkl.load(f)['val']
            out = []
            for task in tasks:
                name_of = {v: k for k, v in task['label_dict'].items()}
                per_video = {}
                for video in task['dict_db']:
                    ann = {}
                    for i, lab in enumerate(video['labels']):
                        ann.setdefault(name_of[lab], []).append([video['segments'][i][0], video['segments'][i][1]])
                    per_video[video['id']] = ann
                out.append(per_video)
            return out
        with open(ground_truth_filename, 'r') as f:
            data = json.load(f)
        out = {}
        for _, v in data.items():
            if not v['subset'] in self.subset:
                continue
            ann = {}
            for a in v['annotations']:
                ann.setdefault(a['label'], []).append([a['segment'][0], a['segment'][1]])
            out[v['clip_id']] = ann
        return out

    def _import_prediction(self, prediction_filename):
        """:93-114."""
        with open(prediction_filename, 'r') as f:
            data = json.load(f)
        if not all(field in data.keys() for field in self.pred_fields):
            raise IOError('Please input a valid prediction file.')
        out = {}
        for vid, props in data['results'].items():
            per_label = {}
            for p in props:
                per_label.setdefault(p['label'], []).append([p['segment'][0], p['segment'][1], p['score']])
            out[vid] = per_label
        return out

    def evaluate(self, current_task_id=None):
        """:116-160 → array (5 tIoU, 2 recall levels): fraction of ground-truth moments that one of the first r·n
        predictions of their (video, label) overlaps by more than t (n = ground truths of that video and label).
        All thresholds and both recall levels are decided from one overlap matrix per (video, label)."""
        ground_truth = self.ground_truth[current_task_id] if self.use_cl else self.ground_truth
        thr = np.asarray(TIOUS)[:, None, None]
        hits = np.zeros((len(TIOUS), len(RECALLS)), np.int64)
        total = 0
        for vid, gt_v in ground_truth.items():
            if vid not in self.prediction:
                raise KeyError(f'no predictions for ground-truth video {vid!r} (the reference stops in pdb here, :134-136)')
            pred_v = self.prediction[vid]
            for label, gt_v_c in gt_v.items():
                n = len(gt_v_c)
                total += n
                if label not in pred_v:
                    continue
                over = iou(pred_v[label], gt_v_c) > thr               # (5, n_pred, n); NaN > t is False
                for j, r in enumerate(RECALLS):
                    hits[:, j] += over[:, :r * n].any(axis=1).sum(axis=1)
        return hits / float(total) if total else np.full(hits.shape, np.nan)


def evaluation_retrieval(gt, pred, subset, tiou, use_cl=False, current_task_id=None):
    """:186-195; gt / pred: file names as in the reference, or loaded objects (see Moment_Retrieval)."""
    mr = Moment_Retrieval(ground_truth_filename=gt, prediction_filename=pred, subset=subset, tiou_thresholds=tiou,
                          verbose=False, check_status=False, use_cl=use_cl)
    return mr.evaluate(current_task_id=current_task_id)
