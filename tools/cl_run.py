"""BASELINE.json configs[1] on synthetic data: query-incremental continual-learning training of the Moment-Query model —
5 sub-tasks, replay memory 1010, classifier grown 22 -> 110 — data-parallel over the GPUs of one box, following the body of the
reference's MQ/train_cl.py:206-388 with the vilco_b200 components in place of libs.modeling / libs.utils:

  per task j:  validate -> [train_one_epoch over the task's clips + the replay memory] x epochs -> validate ->
               add_samples_to_mem(m = memory_size // n_outputs) -> n_known -> augment_classification(+22) -> new optimizer

    python tools/cl_run.py                                   # one GPU
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/cl_run.py --out gpurun_out/r2_cl_8gpu.json

What it records per task: classes, clips (new + replayed), optimizer steps, train videos/s (CUDA events, max over ranks, whole
job), losses of the first / last step, validation videos/s and mAP of the captured evaluation graph re-built for the grown
classifier.  Synthetic features carry no signal, so the accuracy numbers only show that the loop runs end to end.
"""
import argparse
import json
import os
import random
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def task_clips(task, n, classes_per_task, seed, pin=True):
    """n synthetic clips whose ground-truth labels lie in the class range of sub-task `task`"""
    vids = bench.synth_videos(n, seed=seed, K=classes_per_task, pin=pin)
    for i, v in enumerate(vids):
        v["labels"] = v["labels"] + task * classes_per_task
        v["video_id"] = f"task{task}_clip{i}"
    return vids


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tasks", type=int, default=5)
    ap.add_argument("--classes-per-task", type=int, default=22)
    ap.add_argument("--clips-per-task", type=int, default=192)
    ap.add_argument("--val-clips", type=int, default=64)
    ap.add_argument("--memory-size", type=int, default=1010)
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--batch", type=int, default=2, help="clips per GPU per step (train_cfg.batch_size of the MQ configs: 2)")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    random.seed(0)                     # the replay-memory sampling (random.shuffle) must agree on every rank
    from vilco_b200.dist import max_over_ranks, shard_indices
    from vilco_b200.trainer import Trainer, broadcast_parameters, make_optimizer
    from vilco_b200.utils.metrics import ANETdetection
    from vilco_b200.utils.validate import valid_one_epoch
    import pandas as pd

    K0 = a.classes_per_task
    model = bench.build_model(K=K0).cuda()
    model.cl_name = "replay"
    broadcast_parameters(model)
    oc = {"type": "AdamW", "learning_rate": 1e-4, "weight_decay": 0.05}
    opt = make_optimizer(model, oc, flat=True)
    tr = Trainer(model, opt, clip_grad_l2norm=1.0)
    report = {"config": "mq_no_cl.yaml model, query-incremental split: %d sub-tasks x %d classes, %d clips per task, replay memory %d, "
                        "%d clip(s) per GPU per step, %d epoch(s) per task, %d GPU(s)" %
                        (a.tasks, K0, a.clips_per_task, a.memory_size, a.batch, a.epochs, world), "tasks": []}
    val_sets = []

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for j in range(a.tasks):
        new = task_clips(j, a.clips_per_task, K0, seed=1000 * j + 7)
        replay = [v for vs in model.memory.values() for v in vs]
        clips = new + replay
        val_sets += task_clips(j, a.val_clips // a.tasks + 1, K0, seed=1000 * j + 99, pin=False)
        model.train()
        n_steps, first, last = 0, None, None
        t_ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        timed_clips = 0
        for epoch in range(a.epochs):
            order = list(range(len(clips)))
            random.Random(100 * j + epoch).shuffle(order)                  # DistributedSampler.set_epoch(epoch)
            mine = [clips[i] for i in (order[k] for k in shard_indices(len(order), rank, world))]
            n_batches = len(order) // (a.batch * world)                       # drop_last
            for b in range(n_batches):
                batch = mine[b * a.batch:(b + 1) * a.batch]
                if n_steps == 2:
                    barrier()
                    t_ev[0].record()
                losses = tr.step(batch, task_id=j)
                if n_steps >= 2:
                    timed_clips += a.batch * world
                n_steps += 1
                if first is None:
                    first = float(losses["final_loss"].detach())
            last = float(losses["final_loss"].detach())
        t_ev[1].record()
        barrier()
        train_s = max_over_ranks([t_ev[0].elapsed_time(t_ev[1]) * 1e-3], device="cuda")[0] if n_steps > 2 else float("nan")
        # ---- validation on every class seen so far (the captured graph is re-built for the grown classifier) ----
        model.eval()
        gv, g0, g1, gl = [], [], [], []
        for v in val_sets:
            for s, lb in zip(v["segments"].tolist(), v["labels"].tolist()):
                gv.append(v["video_id"]); g0.append(s[0] * v["feat_stride"] / v["fps"]); g1.append(s[1] * v["feat_stride"] / v["fps"]); gl.append(lb)
        gt = pd.DataFrame({"video-id": gv, "t-start": g0, "t-end": g1, "label": gl})
        index = {c: i for i, c in enumerate(sorted(gt["label"].unique()))}
        gt["label"] = gt["label"].map(index)
        evaluator = ANETdetection((gt, index), tiou_thresholds=np.linspace(0.1, 0.5, 5))
        barrier()
        t0 = time.perf_counter()
        _, avg_map, _, _ = valid_one_epoch(val_sets, model, j, evaluator=evaluator, batch_size=16, text_len=128)
        barrier()
        val_s = time.perf_counter() - t0
        # ---- replay memory + growth (train_cl.py:343-378) ----
        n_out = model.cls_head.cls_head.conv.out_channels
        m = a.memory_size // n_out
        data = {}
        for v in new:
            for c in set(v["labels"].tolist()):
                data.setdefault(int(c), []).append(v)
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            model.add_samples_to_mem(None, data, m)
        model.n_known = len(model.memory)
        rec = {"task": j, "classes": n_out, "clips": len(clips), "replayed": len(replay), "steps": n_steps,
               "train_videos_per_s": timed_clips / train_s if train_s == train_s else None,
               "train_ms_per_step": 1e3 * train_s / max(n_steps - 2, 1) if train_s == train_s else None,
               "loss_first": first, "loss_last": last, "val_clips": len(val_sets), "val_videos_per_s": len(val_sets) / val_s,
               "val_avg_mAP": float(avg_map), "memory_classes": len(model.memory), "memory_clips_per_class": m}
        report["tasks"].append(rec)
        if rank == 0:
            print(json.dumps(rec), flush=True)
        if j + 1 < a.tasks:
            model.augment_classification(K0, model.device)
            opt = make_optimizer(model, oc, flat=True)           # new optimizer for the grown model, like the reference
            tr = Trainer(model, opt, clip_grad_l2norm=1.0)
    if rank == 0:
        report["n_gpus"] = world
        print(json.dumps(report), flush=True)
        if a.out:
            with open(a.out, "w") as f:
                json.dump(report, f, indent=1)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
