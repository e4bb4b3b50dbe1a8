"""Host logic of the validation pass (vilco_b200/utils/validate.py) on CPU: regrouping into static batches with a padded last
batch, sharding by clip over a world_size-2 gloo group with the loader order restored, result table, in-memory evaluator
and recall.  The inference engine is injected (an object with EvalGraph's `infer_stream`), so no GPU is needed."""
import os

import numpy as np
import pandas as pd
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class FakeGraph:
    """deterministic detections per clip, derived from the clip's own fields (so any mix-up of clips shows)."""

    def __init__(self, batch_size):
        self.B, self.calls = batch_size, 0

    def infer_stream(self, batches):
        for vl in batches:
            assert len(vl) == self.B
            self.calls += 1
            out = []
            for v in vl:
                k = int(v["video_id"][1:])
                n = k % 4                                   # clip 0, 4, ...: no detections at all
                seg = torch.tensor([[10.0 * k + j, 10.0 * k + j + 5.0] for j in range(n)], dtype=torch.float32).reshape(n, 2)
                out.append({"video_id": v["video_id"], "segments": seg, "labels": torch.full((n,), k % 3, dtype=torch.int64),
                            "scores": torch.linspace(0.9, 0.5, n) if n else torch.zeros(0)})
            yield out


def _loader(n):
    """the reference's validation loader yields lists of one clip; a second shape (lists of 3) must give the same pass."""
    return [[{"video_id": f"v{i}"}] for i in range(n)]


def _ground_truth(n):
    rows = [(f"v{k}", 10.0 * k + j, 10.0 * k + j + 5.0, k % 3) for k in range(n) for j in range(max(1, k % 4))]
    gt = pd.DataFrame(rows, columns=["video-id", "t-start", "t-end", "label"])
    return gt, {0: 0, 1: 1, 2: 2}


def _run(n, batch_size, loader=None):
    from vilco_b200.utils.metrics import ANETdetection
    from vilco_b200.utils.validate import valid_one_epoch

    class M(torch.nn.Module):
        pass
    g = FakeGraph(batch_size)
    ev = ANETdetection(_ground_truth(n), tiou_thresholds=np.linspace(0.1, 0.5, 5))
    # recall table over every clip: clips 0, 4, ... have a ground truth but no detections -> counted as not retrieved
    ret_gt = {f"v{k}": {k % 3: [[10.0 * k + j, 10.0 * k + j + 5.0] for j in range(max(1, k % 4))]} for k in range(n)}
    res = valid_one_epoch(loader if loader is not None else _loader(n), M(), 0, evaluator=ev, graph=g, batch_size=batch_size,
                          retrieval_gt=ret_gt, print_freq=2)
    return res, g, ev


def test_single_process_pass_pads_the_last_batch_and_evaluates():
    (mAP, avg, thr, rec), g, ev = _run(7, 3)
    assert g.calls == 3                                     # 7 clips -> 3 + 3 + (1 real + 2 repeats)
    assert np.allclose(thr, np.linspace(0.1, 0.5, 5))
    # every detection coincides with a ground truth of its clip and label; clips 0 and 4 have no detections but one
    # ground truth each -> recall of their labels is below 1, precision stays 1
    assert avg > 0.5 and np.all(mAP <= 1.0)
    n_gt = sum(max(1, k % 4) for k in range(7))
    assert rec.shape == (5, 2) and np.allclose(rec, (n_gt - 2) / n_gt)     # all but the two detection-less clips' moments
    # same pass with another batch size and another loader grouping gives identical numbers
    (mAP2, avg2, _, rec2), g2, _ = _run(7, 32, loader=[[{"video_id": f"v{i}"} for i in range(j, min(7, j + 3))] for j in (0, 3, 6)])
    assert g2.calls == 1 and np.array_equal(mAP, mAP2) and avg == avg2 and np.array_equal(rec, rec2)


def test_results_table_layout():
    from vilco_b200.utils.validate import results_table
    outs = next(FakeGraph(4).infer_stream([[{"video_id": f"v{i}"} for i in range(4)]]))
    t = results_table(outs)
    assert t["video-id"] == ["v1", "v2", "v2", "v3", "v3", "v3"]
    assert t["t-start"].dtype == np.float32 and t["label"].dtype == np.int64 and len(t["score"]) == 6
    empty = results_table([outs[0]])
    assert empty["video-id"] == [] and empty["t-start"].shape == (0,)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    (mAP, avg, _, rec), g, ev = _run(9, 2)
    q.put((rank, g.calls, mAP.tolist(), float(avg), rec.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_by_clip_and_agree_with_one_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (mAP, avg, _, rec), g, _ = _run(9, 2)
    assert g.calls == 5
    assert out[0][1] == 3 and out[1][1] == 2                # rank 0: clips 0,2,4,6,8 -> 3 batches; rank 1: 1,3,5,7 -> 2
    for _, _, m, a, r in out:
        assert m == mAP.tolist() and a == float(avg) and r == rec.tolist()


def test_prompt_models_are_evaluated_eagerly_in_the_same_batches():
    """a model with an L2P prompt pool cannot be graph-captured: the pass calls model(batch, is_training=False, task_id=...)
    on the regrouped batches instead and gives the same table."""
    from vilco_b200.utils.metrics import ANETdetection
    from vilco_b200.utils.validate import valid_one_epoch
    calls = []

    class PromptModel(torch.nn.Module):
        prompt = object()

        def forward(self, video_list, task_id=-1, is_training=True):
            assert not is_training
            calls.append((len(video_list), task_id))
            return next(FakeGraph(len(video_list)).infer_stream([video_list]))

        def make_eval_graph(self, *a, **k):
            raise AssertionError("a prompt model must not be graph-captured")

    ev = ANETdetection(_ground_truth(7), tiou_thresholds=np.linspace(0.1, 0.5, 5))
    mAP, avg, _, rec = valid_one_epoch(_loader(7), PromptModel(), 0, evaluator=ev, batch_size=3, task_id=2)
    assert calls == [(3, 2), (3, 2), (3, 2)] and rec is None
    (mAP2, avg2, _, _), _, _ = _run(7, 3)
    assert np.array_equal(mAP, mAP2) and avg == avg2
