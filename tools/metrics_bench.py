"""Evaluation-tail timing: vilco_b200.utils.metrics.ANETdetection + get_retrieval_performance on a synthetic
validation set of N videos x 200 detections, 22 labels; with --reference (authoring container only, needs /root/reference)
the reference's own evaluator is timed on the same tables (one process) and the AP matrices compared.

    python tools/metrics_bench.py [--videos 200] [--reference]
"""
import argparse
import io
import os
import sys
import tempfile
import time
from contextlib import redirect_stdout

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gen_golden_metrics as G          # noqa: E402  (case generator only; bench tooling, not the product)
from vilco_b200.utils import metrics as M           # noqa: E402
from vilco_b200.utils import get_retrieval_performance as R  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--videos", type=int, default=200)
    ap.add_argument("--reference", action="store_true")
    a = ap.parse_args()
    case = G.make_case(7, a.videos, 22, 200)
    tmp = tempfile.mkdtemp()
    ann = os.path.join(tmp, "ann.json")
    G.write_annotation_json(case, ann)
    preds = G.preds_dict(case)
    tious = np.linspace(0.1, 0.5, 5)
    ev = M.ANETdetection(ann, "val", tiou_thresholds=tious)
    t0 = time.perf_counter()
    _, avg, _ = ev.evaluate(dict(preds), verbose=False)
    t1 = time.perf_counter()
    pred_tab = R.predictions_from_results(preds)
    t2 = time.perf_counter()
    print(f"vilco_b200: mAP over {len(preds['score'])} detections / {len(case['gt'])} ground truths: {t1 - t0:.3f} s "
          f"(avg mAP {avg:.4f}); prediction table for the recall metric: {t2 - t1:.3f} s")
    if a.reference:
        np.float = float
        ref = G._load("metrics")
        rev = ref.ANETdetection(ann, "val", tiou_thresholds=tious, num_workers=1)   # in-process: loky workers would not see the np.float shim
        t0 = time.perf_counter()
        with redirect_stdout(io.StringIO()):
            _, ravg, _ = rev.evaluate(dict(preds), verbose=False)
        t1 = time.perf_counter()
        print(f"reference : {t1 - t0:.3f} s in one process (the reference's 8 joblib workers divide this by at most 8) (avg mAP {ravg:.4f}); "
              f"AP matrices equal: {np.array_equal(rev.ap, ev.ap)}")


if __name__ == "__main__":
    main()
