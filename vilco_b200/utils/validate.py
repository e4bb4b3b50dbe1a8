"""Validation pass — the caller on the far side of the hot path (`valid_one_epoch`, MQ/libs/utils/train_utils.py:658-808;
SURVEY.md §8e/§8f): clips → model → soft-NMS → result table → detection mAP + retrieval recall.

The reference runs one clip per `model(...)` call, appends per-clip tensors, writes every detection to a json file with
`.item()` and reads it back for the recall metric, then walks the table row by row in the evaluator.  Here the same pass is

* batched and pipelined: clips are regrouped into `batch_size` and fed to `EvalGraph.infer_stream` (one CUDA graph per
  batch, the upload of batch i+1 overlapping batch i; clips may carry `feats_raw`, see vilco_b200/data.py),
* sharded by clip over the ranks of an initialised process group with no data-path collective (results are gathered
  host-side, `vilco_b200.dist.gather_results`, and restored to loader order),
* evaluated in memory by `vilco_b200.utils.metrics.ANETdetection` / `get_retrieval_performance` (bit-identical numbers).

Return value and logging follow the reference: `(mAP, avg_mAP, tiou_thresholds, eval_result)`.
"""
import logging
import pickle
import time

import numpy as np
import torch

from .. import dist as D
from . import get_retrieval_performance as R

_log = logging.getLogger("vilco_b200.validate")


def _batches(clips, batch_size):
    """groups of exactly `batch_size` clips; the last group is filled up by repeating its last clip (static graph shape) and
    the number of real clips is yielded next to it."""
    buf = []
    for c in clips:
        buf.append(c)
        if len(buf) == batch_size:
            yield buf, batch_size
            buf = []
    if buf:
        n = len(buf)
        yield buf + [buf[-1]] * (batch_size - n), n


def _own_clips(val_loader, rank, world, sharded_loader):
    """clips of this rank in loader order.  A loader that is already sharded (DistributedSampler) is taken as is; otherwise
    clip i of the flattened loader belongs to rank i % world (vilco_b200.dist.shard_indices)."""
    i = 0
    for video_list in val_loader:
        for v in (video_list if isinstance(video_list, (list, tuple)) else [video_list]):
            if sharded_loader or i % world == rank:
                yield v
            i += 1


def results_table(outputs):
    """list of per-clip result dicts -> the evaluator's table (train_utils.py:697-709, 753-756): clips without detections
    contribute no rows."""
    vids, t0, t1, lab, sc = [], [], [], [], []
    for o in outputs:
        n = o['segments'].shape[0]
        if n > 0:
            vids.extend([o['video_id']] * n)
            t0.append(o['segments'][:, 0])
            t1.append(o['segments'][:, 1])
            lab.append(o['labels'])
            sc.append(o['scores'])
    cat = lambda xs, dt: torch.cat(xs).numpy() if xs else np.zeros(0, dt)      # noqa: E731
    return {'video-id': vids, 't-start': cat(t0, np.float32), 't-end': cat(t1, np.float32), 'label': cat(lab, np.int64),
            'score': cat(sc, np.float32)}


class _EagerStream:
    """`infer_stream` for models whose evaluation step cannot be captured in a CUDA graph — the L2P prompt pool of
    `mq_vilco.yaml` selects its prompts on the host (`EvalGraph` refuses such a model): the same batches go through
    `model(batch, is_training=False, task_id=...)`, i.e. the same kernels launched eagerly."""

    def __init__(self, model, task_id=-1):
        self.model, self.task_id = model, task_id

    def infer_stream(self, batches):
        for batch in batches:
            yield self.model(batch, task_id=self.task_id, is_training=False)


def valid_one_epoch(val_loader, model, curr_epoch, ext_score_file=None, evaluator=None, output_file=None, tb_writer=None,
                    print_freq=20, logger=None, dataset_name=None, *, batch_size=32, text_len=128, graph=None,
                    sharded_loader=False, current_task_id=None, retrieval_gt=None, idx_classes=None, use_cl=False,
                    task_id=-1):
    """Drop-in for `valid_one_epoch(val_loader, model, curr_epoch, ...)`; keyword-only extensions after `*`:
    batch_size / text_len of the captured graph, `graph` (an existing `EvalGraph`), `sharded_loader`, `current_task_id` +
    `use_cl` for the query-incremental evaluator, `retrieval_gt` (annotation file or loaded object) + `idx_classes`
    (integer label → ground-truth label key) to get the recall table without the json round trip, `task_id` for models
    with a prompt pool (evaluated eagerly, see `_EagerStream`)."""
    assert (evaluator is not None) or (output_file is not None)
    if ext_score_file is not None:
        raise NotImplementedError("valid_one_epoch: external classification scores (postprocess_results) are not part of "
                                  "the Moment-Query path")
    logger = logger or _log
    model.eval()
    rank, world = 0, 1
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        rank, world = torch.distributed.get_rank(), torch.distributed.get_world_size()
    if graph is not None:
        g = graph
    elif hasattr(model, "prompt"):
        g = _EagerStream(model, task_id)
    else:
        g = model.make_eval_graph(batch_size, text_len=text_len)
    real = []
    start = time.time()

    def feed():
        for k, (batch, n) in enumerate(_batches(_own_clips(val_loader, rank, world, sharded_loader), batch_size)):
            real.append(n)
            yield batch

    local, k = [], 0
    with torch.no_grad():
        for out in g.infer_stream(feed()):
            local.extend(out[:real[k]])
            k += 1
            if k % print_freq == 0:
                logger.info('Test: [{0:05d}]\tTime {1:.2f} s per batch of {2}'.format(k, (time.time() - start) / k, batch_size))
    outputs = D.gather_results(local) if (world > 1 and not sharded_loader) else \
        (_concat_ranks(local, world) if world > 1 else local)
    results = results_table(outputs)

    eval_result = None
    if retrieval_gt is not None:
        pred = R.predictions_from_results(results, idx_classes)
        for o in outputs:                       # a clip without detections retrieves nothing (the reference's evaluate()
            pred.setdefault(o['video_id'], {})  # stops in pdb when a ground-truth clip has no prediction entry, :133-136)
        eval_result = R.evaluation_retrieval(gt=retrieval_gt, pred=pred, subset="val", tiou=list(R.TIOUS), use_cl=use_cl,
                                             current_task_id=current_task_id)
        for i, t in enumerate(R.TIOUS):
            for j, r in enumerate(R.RECALLS):
                logger.info(f'Rank {r}x @ tIoU {t} is {eval_result[i, j]}')
    if evaluator is None:
        if rank == 0:
            with open(output_file, "wb") as f:
                pickle.dump(results, f)
        return None, 0.0, None, eval_result
    mAP, avg_mAP, tiou_thresholds = evaluator.evaluate(results, current_task_id=current_task_id, verbose=False) \
        if current_task_id is not None else evaluator.evaluate(results, verbose=False)
    for tiou, tiou_mAP in zip(tiou_thresholds, mAP):
        logger.info(f'tIoU = {tiou:.1f}: mAP = {tiou_mAP * 100:.2f} %')
    logger.info(f'Average Map is :{avg_mAP * 100: .2f} %')
    return mAP, avg_mAP, tiou_thresholds, eval_result


def _concat_ranks(local, world):
    """loader already sharded: concatenate the ranks' lists in rank order (no interleaving to undo)."""
    buckets = [None] * world
    torch.distributed.all_gather_object(buckets, list(local))
    return [o for b in buckets for o in b]
