#!/usr/bin/env python
"""Benchmark of the Moment-Query hot path (BASELINE.json metric: MQ train videos/s & infer videos/s incl. soft-NMS).
The headline `value` is inference; the training half of the metric is the `train` object of the same JSON line.

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's OWN modules on the host cores (baseline/_ref)
    python bench.py --metric train ...                       # same run, the training half reported as the headline value

A step = one batch of `--batch` synthetic clips (features N(0,1) of shape (4096, 1024), CLIP-token text (768, L),
mq_no_cl.yaml model, K=22 classes, random-init weights -> worst-case NMS load) through
pack -> backbone -> neck -> heads -> decode -> soft-NMS.
  value : videos/s with the inputs already resident in HBM (CUDA-graph replay, CUDA-event timing, max over ranks)
  e2e   : videos/s through the public streaming API `EvalGraph.infer_stream(batches)` with pinned HOST inputs, including
          every step's H2D copies (double-buffered against the previous step's compute) and the D2H read of the detections.
  train : one iteration of the reference's train_one_epoch loop (zero_grad, forward with dropout / drop-path, loss,
          backward, NCCL gradient all-reduce for N > 1, clip_grad_norm 1.0, AdamW) on `--train-batch` clips per GPU.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "mq_no_cl.yaml inference + soft-NMS, T=1024, input 4096-d, C=1024, H=16, arch [2,2,9], K=22, text L~U[20,128]"


def synth_videos(n, seed, T=1024, Cin=4096, Ct=768, K=22, pin=False):
    rs = np.random.RandomState(1000 + seed)
    out = []
    for i in range(n):
        feats = torch.from_numpy(rs.standard_normal((Cin, T)).astype(np.float32))
        text = torch.from_numpy(rs.standard_normal((Ct, int(rs.randint(20, 129)))).astype(np.float32))
        ng = int(rs.randint(1, 9))
        c = rs.uniform(0, T, ng)
        ln = np.exp(rs.uniform(np.log(4.0), np.log(512.0), ng))
        s0 = np.clip(c - ln / 2, 0, T).astype(np.float32)
        s1 = np.maximum(np.clip(c + ln / 2, 0, T), s0 + 1).astype(np.float32)
        if pin:
            feats, text = feats.pin_memory(), text.pin_memory()
        out.append({"video_id": f"syn_{seed}_{i}", "feats": feats, "prompt_feature": text,
                    "segments": torch.from_numpy(np.stack([s0, s1], 1)), "labels": torch.from_numpy(rs.randint(0, K, ng).astype(np.int64)),
                    "fps": 30.0, "duration": 480.0, "feat_stride": 480.0 * 30.0 / T, "feat_num_frames": 480.0 * 30.0 / T,
                    "segmentation_labels": torch.zeros(T, K)})
    return out


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, gpu, enabled=True, period=0.1):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag, self.enabled, self.period = gpu, [], False, enabled, period

    def run(self):
        if not self.enabled:     # only rank 0 polls: N pollers per node perturb the ranks they are supposed to observe
            return
        try:                     # NVML in-process (nvidia_ml_py): far lighter than spawning nvidia-smi several times a second
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            bits = [("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                    ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap)]
            while not self.stop_flag:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.rows.append([str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(mx)] +
                                 ["Active" if r & bit else "Not Active" for _, bit in bits])
                time.sleep(self.period)
            return
        except Exception:
            pass
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            time.sleep(0.3)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(len(r) > 2 + j and r[2 + j].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return j["hbm_gbs"], j["bf16_tflops"], j["bf16_tflops_sustained"], "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def build_model(K=22):
    from vilco_b200.config import mq_model_kwargs
    from vilco_b200.modeling import make_meta_arch
    torch.manual_seed(0)
    return make_meta_arch("LocPointTransformer", **mq_model_kwargs(num_classes=K))


# ----------------------------------------------------------------------------------------------------------------
# Baselines: the REFERENCE's own modules (un-modified MQ tree under baseline/_ref or /root/reference, imported through
# oracle/ref_shim.py, soft-NMS = its own nms_cpu.cpp compiled into oracle/_ref) — "kind": "reference"; when that tree is not
# there, the oracle port of the same path — "kind": "port".  Never on the product path.
# ----------------------------------------------------------------------------------------------------------------
def reference_model(device="cpu"):
    """(callable infer(video) -> result, callable train_step(videos) -> None, kind, description)"""
    try:
        from oracle import ref_shim
        if ref_shim.available():
            torch.manual_seed(0)
            model, _ = ref_shim.build_model(None, "mq_no_cl.yaml")
            model = model.to(device).eval()
            opt = [None]

            def infer(v):
                with torch.no_grad():
                    return model([v], is_training=False)[0]

            def train_step(vs):
                if opt[0] is None:
                    opt[0] = torch.optim.AdamW([q for q in model.parameters() if q.requires_grad], lr=1e-4, weight_decay=0.05)
                model.train()
                opt[0].zero_grad()
                model(vs, is_training=True)["final_loss"].backward()
                torch.nn.utils.clip_grad_norm_([q for q in model.parameters() if q.grad is not None], 1.0)
                opt[0].step()
                model.eval()
            return infer, train_step, "reference", "the reference's PtTransformer (MQ/libs/modeling, mq_no_cl.yaml) + its nms_cpu.cpp"
    except Exception as e:   # a baseline must never take the bench down
        print(f"[bench] reference modules unavailable ({e!r}); using the oracle port", file=sys.stderr)
    from oracle import mq_oracle as O
    from oracle import nms_c
    cfg = O.ModelCfg()
    sd = build_model().state_dict()
    P = {k: v.detach().float().to(device).clone() for k, v in sd.items()
         if torch.is_floating_point(v) and not k.startswith("backbone.xlnet.word_embedding")}
    opt = [None]
    if device != "cpu":
        orig_pe = O.sinusoid_pe
        O.sinusoid_pe = lambda *a: orig_pe(*a).to(device)

    def infer(v):
        with torch.no_grad(), torch.device(device):
            x, mask, text, tmask = O.preprocess(cfg, [v], False)
            logits, offs, masks, _ = O.forward_heads(P, cfg, x.to(device), mask.to(device), text.to(device), tmask.to(device), training=False)
            pts = [q.to(device) for q in O.points(cfg, [m.shape[1] for m in masks])]
            segs, scores, labels = O.decode_single_video(cfg, pts, [m[0] for m in masks], [l[0] for l in logits], [o[0] for o in offs])
        return O.postprocess(cfg, segs.cpu(), scores.cpu(), labels.cpu(), v["fps"], v["duration"], v["feat_stride"],
                             v["feat_num_frames"], nms_c.softnms_1d)

    def train_step(vs):
        if opt[0] is None:
            for q in P.values():
                q.requires_grad_(True)
            opt[0] = torch.optim.AdamW(list(P.values()), lr=1e-4, weight_decay=0.05)
        opt[0].zero_grad()
        with torch.device(device):
            vv = [{**v, "feats": v["feats"].to(device), "prompt_feature": v["prompt_feature"].to(device),
                   "segments": v["segments"].to(device), "labels": v["labels"].to(device)} for v in vs]
            lo, _ = O.model_train_losses(P, cfg, vv)
        lo["final_loss"].backward()
        torch.nn.utils.clip_grad_norm_([q for q in P.values() if q.grad is not None], 1.0)
        opt[0].step()
    return infer, train_step, "port", "oracle/mq_oracle.py (torch restatement of the reference path) + oracle/softnms.c"


def cpu_baseline_leg(n_videos, seed=7):
    """bounded sample on the box's host cores: n_videos clips, one at a time like the reference's evaluation loop"""
    torch.set_num_threads(os.cpu_count() or 1)
    infer, _, kind, what = reference_model("cpu")
    vids = synth_videos(n_videos + 1, seed)
    infer(vids[0])                                  # warm-up
    t0 = time.perf_counter()
    for v in vids[1:]:
        infer(v)
    dt = time.perf_counter() - t0
    cores = os.cpu_count() or 1
    return {"value": n_videos / dt, "unit": "videos/s", "cores": cores, "kind": kind,
            "sample": f"{n_videos} clips, one per call, through {what}; torch CPU fp32, {cores} threads, incl. soft-NMS; {dt:.1f} s"}


def eager_gpu_rates(tf32, n_inf=6, n_train=2):
    """The reference's own modules run as eager PyTorch fp32 ON THE B200 (cuDNN / cuBLAS kernels behind torch ops, batch-1
    evaluation, soft-NMS on the host through its nms_cpu.cpp): the practical bar a GPU user of the reference has today."""
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = tf32
    try:
        infer, train_step, kind, _ = reference_model("cuda")
        vids = synth_videos(n_inf + 1, 21)
        infer(vids[0])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for v in vids[1:]:
            infer(v)
        torch.cuda.synchronize()
        inf_rate = n_inf / (time.perf_counter() - t0)
        tv = synth_videos(n_train, 22)
        train_step(tv)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            train_step(tv)
        torch.cuda.synchronize()
        tr_rate = 3 * n_train / (time.perf_counter() - t0)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    return inf_rate, tr_rate, kind


def next_rows():
    """The rows either side of the hot path timed beside it, each next to the reference's CPU way of doing the
    same thing: the dataset's linear time-resize of a stored clip (vilco_resize_feats vs F.interpolate on the host) and the
    evaluation tail (ANETdetection mirror on 200 clips x 200 detections; the reference's pandas walk takes ~55 s for that
    table, tools/metrics_bench.py --reference).  Reported extras: a failure here never takes the bench line down."""
    out = {}
    try:
        import torch.nn.functional as F
        from vilco_b200 import lib as L
        from vilco_b200 import ops
        Bc, T_in, C, T = 32, 512, 4096, 1024
        x = torch.randn(Bc * T_in, C, device="cuda")
        row_start = torch.arange(0, (Bc + 1) * T_in, T_in, dtype=torch.int64, device="cuda")
        o16 = ops.empty16(Bc, T, C)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for i in range(13):                      # 3 warm-ups; input 268 MB + output 537 MB per launch exceed the L2
            if i == 3:
                ev[0].record()
            L.check(L.lib().vilco_resize_feats(ops._p(x), ops._p(row_start), Bc, C, T, None, ops._p(o16), ops._i64(ops.lo(o16)),
                                               L.stream_ptr()), "vilco_resize_feats")
        ev[1].record()
        torch.cuda.synchronize()
        sec = ev[0].elapsed_time(ev[1]) / 10 * 1e-3
        nbytes = x.numel() * 4 + o16.numel() * 2
        xc = x[:T_in].cpu()
        t0 = time.perf_counter()
        for _ in range(3):
            F.interpolate(xc.permute(1, 0).unsqueeze(0), size=T, mode="linear", align_corners=False)
        cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
        hbm = peaks()[0]
        out["data_path"] = {"kernel": "vilco::resize_feats_kernel", "workload": "32 stored clips 512x4096 fp32 -> 1024 rows as the operand planes of the shipped mode (%d x 16 bit)" % o16.shape[0],
                            "us_per_launch": sec * 1e6, "bound": "hbm", "achieved": nbytes / sec / 1e9, "peak": hbm, "unit": "GB/s",
                            "frac": nbytes / sec / 1e9 / hbm, "clips_per_s": Bc / sec, "cpu_interpolate_ms_per_clip": cpu_ms}
        del x, o16
    except Exception as e:
        out["data_path"] = {"unavailable": repr(e)[:200]}
    try:
        import pandas as pd
        from vilco_b200.utils.metrics import ANETdetection
        rs = np.random.RandomState(5)
        gv, g0, g1, gl, pv, p0, p1, pl, ps = [], [], [], [], [], [], [], [], []
        for v in range(200):
            n = int(rs.randint(1, 9))
            c, ln = rs.uniform(0, 480, n), np.exp(rs.uniform(0, np.log(120), n))
            lab = rs.randint(0, 22, n)
            gv += [f"clip{v}"] * n
            g0 += list(np.maximum(c - ln / 2, 0)); g1 += list(c + ln / 2); gl += list(lab)
            k = rs.randint(0, n, 200)
            j0 = np.maximum(c[k] - ln[k] / 2 + rs.normal(0, 0.2, 200) * ln[k], 0)
            pv += [f"clip{v}"] * 200
            p0 += list(j0); p1 += list(j0 + ln[k] * np.exp(rs.normal(0, 0.2, 200)))
            pl += list(np.where(rs.rand(200) < 0.7, lab[k], rs.randint(0, 22, 200))); ps += list(rs.beta(0.5, 4.0, 200))
        gt = pd.DataFrame({"video-id": gv, "t-start": g0, "t-end": g1, "label": gl})
        index = {j: i for i, j in enumerate(sorted(gt["label"].unique()))}
        gt["label"] = gt["label"].map(index)
        evl = ANETdetection((gt, index), tiou_thresholds=np.linspace(0.1, 0.5, 5))
        preds = {"video-id": pv, "t-start": np.float32(p0), "t-end": np.float32(p1), "label": np.int64(pl), "score": np.float32(ps)}
        t0 = time.perf_counter()
        _, avg, _ = evl.evaluate(preds, verbose=False)
        dt = time.perf_counter() - t0
        out["evaluation_tail"] = {"workload": "detection mAP@0.1:0.5 of 200 clips x 200 detections, 22 labels (host code, vilco_ap_match)",
                                  "seconds": dt, "detections_per_s": len(ps) / dt, "avg_mAP": float(avg),
                                  "reference": "the reference's evaluator on a table of this size: ~55 s in one process (tools/metrics_bench.py --reference)"}
    except Exception as e:
        out["evaluation_tail"] = {"unavailable": repr(e)[:200]}
    try:      # §8f-1: the NLQ model (evaluation) on the same kernels, full size (T = 2560, C = 384, 4 heads of 96, window 9, 7 levels)
        from vilco_b200 import lib as L
        from vilco_b200.modeling import make_meta_arch
        nlq = make_meta_arch("NlqLocPointTransformer", regression_range=[[0, 4], [2, 8], [4, 16], [8, 32], [16, 64], [32, 128], [64, 10000]],
                             test_cfg=dict(voting_thresh=0.9, pre_nms_topk=2000, max_seg_num=5, min_score=0.001, nms_sigma=0.75,
                                           duration_thresh=0.001)).cuda().eval()
        g = torch.Generator().manual_seed(1)
        Bq = 16
        clips = [{"video_id": f"v{i}", "query_id": f"q{i}", "feats": torch.randn(256, 2560 - 17 * i, generator=g),
                  "query_feats": torch.randn(512, 12, generator=g), "fps": 30.0, "duration": 1400.0, "feat_stride": 16.043,
                  "feat_num_frames": 16.043} for i in range(Bq)]
        for _ in range(3):
            nlq(clips, is_training=False)
        torch.cuda.synchronize()
        n0 = L.launch_count()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(5):
            res = nlq(clips, is_training=False)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 5
        from vilco_b200.modeling.nlq import NlqEvalGraph
        gq = NlqEvalGraph(nlq, Bq, 12)
        for _ in range(3):
            gq.run(clips)
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(5):
            res_g = gq.run(clips)
        ev[1].record()
        torch.cuda.synchronize()
        ms_g = ev[0].elapsed_time(ev[1]) / 5
        out["nlq_model"] = {"workload": "ego4d_nlq_v2_egovlp_1e-4.yaml evaluation, 16 queries per step, T = 2560 (host batching, upload, "
                                        "decode + soft-NMS, result download inside the timed region); captured CUDA graph, eager beside it",
                            "ms_per_step": ms_g, "queries_per_s": Bq / ms_g * 1e3, "gpu_launches_per_step": gq.launches,
                            "eager_ms_per_step": ms, "eager_queries_per_s": Bq / ms * 1e3,
                            "graph_equals_eager": bool(all(torch.equal(a["scores"], b["scores"]) for a, b in zip(res, res_g))),
                            "operand_mode": nlq.operand_mode, "segments_per_query": int(res[0]["segments"].shape[0])}
        del gq
        del nlq
    except Exception as e:
        out["nlq_model"] = {"unavailable": repr(e)[:200]}
    return out


def run_reference_arm(args):
    """`--impl reference`: the reference's own CPU implementation of the path on the box's host cores, all threads.  EXACTLY
    `--warmup` untimed + `--steps` timed steps are run; a step = a bounded sample (`--ref-videos` clips, evaluated one per call
    as the reference's loop does) of our arm's step, so that the whole run ends within a few minutes.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    cores = os.cpu_count() or 1
    n = max(1, args.ref_videos)
    infer, train_step, kind, what = reference_model("cpu")
    vids = synth_videos(n * 2, seed=7)
    for i in range(args.warmup):
        for v in vids[:n]:
            infer(v)
    t0 = time.perf_counter()
    for i in range(args.steps):
        for v in vids[(i % 2) * n:(i % 2) * n + n]:
            infer(v)
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    t1 = time.perf_counter()
    train_step(synth_videos(2, seed=11))
    tr_dt = time.perf_counter() - t1
    sample = (f"{n} clip(s) per step, one per call, through {what}; torch CPU fp32, {cores} threads, incl. soft-NMS; "
              f"{args.warmup} warm-up + {args.steps} timed steps, {dt:.1f} s timed")
    line = {"impl": "reference", "metric": "mq_infer_videos_per_s", "value": v, "unit": "videos/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args, reference_sample=n),
            "cpu_baseline": {"value": v, "unit": "videos/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "videos/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "train_value": 2 / tr_dt,
            "train": {"metric": "mq_train_videos_per_s", "value": 2 / tr_dt, "unit": "videos/s", "videos_per_step": 2,
                      "sample": f"1 un-warmed step (forward, loss, autograd backward, clip, AdamW) of 2 clips, {tr_dt:.1f} s"}}
    print(json.dumps(line), flush=True)


def bench_config(args, reference_sample=None):
    """the `config` object: identical in both arms"""
    from vilco_b200 import ops
    modes = {"mixed": "fp16 operand planes, fp32 accumulate; one plane (1 tcgen05.mma per k-step) except the input projection, "
                      "embedding convs and channel-attention qkv / core (split operands, 3 MMAs): logits / offsets within 1e-3",
             "fp16x3": "fp16 hi+lo planes everywhere (3 MMAs per k-step, ~1e-5)", "bf16x3": "bf16 hi+lo planes everywhere (3 MMAs, ~1e-5)",
             "fp16": "fp16 single planes everywhere", "bf16": "bf16 single planes everywhere"}
    c = {"workload": WORKLOAD, "videos_per_step_per_gpu": args.batch, "precision": ops.precision() + ": " + modes[ops.precision()],
         "l2": "working set per step (packed weights 0.9 GB + activations) exceeds the 126 MB L2; no explicit flush",
         "train_videos_per_step_per_gpu": args.train_batch}
    if reference_sample is not None:
        c["reference_sample_videos_per_step"] = reference_sample
    return c


def verify_against_oracle(model, video):
    """clip 0 of the bench batch, outside the timed region: head outputs vs the fp32 oracle, and the decode + soft-NMS kernels
    vs the reference algorithm on the CUDA path's own head outputs (tests/util.kernel_parity_on_own_outputs)"""
    from oracle import mq_oracle as O
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import kernel_parity_on_own_outputs, rel_max
    cfg = O.ModelCfg()
    P = {k: v.detach().float().cpu() for k, v in model.state_dict().items() if torch.is_floating_point(v)}
    cls_l, off_l, _ = model([video], is_training=False, get_emb=True)
    with torch.no_grad():
        x, mask, text, tmask = O.preprocess(cfg, [video], False)
        lg, of, _, _ = O.forward_heads(P, cfg, x, mask, text, tmask, training=False)
    kp = kernel_parity_on_own_outputs(cfg, model, video)
    return {"logits_rel_err": rel_max(torch.cat(cls_l, 1)[0].cpu(), torch.cat(lg, 1)[0]),
            "offsets_rel_err": rel_max(torch.cat(off_l, 1)[0].cpu(), torch.cat(of, 1)[0]),
            "decode_nms_vs_reference_algorithm": {k: kp[k] for k in ("cand_count", "cand_count_oracle", "score", "rank_swaps", "seg", "orphans")},
            "against": "oracle/mq_oracle.py fp32 on the host (pinned to the reference's outputs by tests/golden/model_full.npz), "
                       "random-init weights, clip 0 of the timed batch; bar 1e-3"}


# ----------------------------------------------------------------------------------------------------------------
def gemm_roofline(graph_runner, model, B):
    """Time every tcgen05 GEMM launch of one step with CUDA events on the launching stream and relate the algorithmic
    FLOPs (2*M*N*K*taps per launch, summed) to the measured bf16 peak."""
    from vilco_b200 import lib as L
    rec = []
    orig = L.gemm

    def timed(A, Bm, D, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig(A, Bm, D, **kw)
        e1.record()
        Z = kw.get("Z", (1, 1))
        rec.append((e0, e1, 2.0 * kw["M"] * kw["N"] * kw["K"] * kw.get("taps", 1) * Z[0] * Z[1],
                    (kw["M"], kw["N"], kw["K"], kw.get("taps", 1), Z[0] * Z[1], str(D.dtype)[6:])))
        return r

    L.gemm = timed
    try:
        for _ in range(2):
            rec.clear()
            graph_runner._step()
            torch.cuda.synchronize()
    finally:
        L.gemm = orig
    t = sum(a.elapsed_time(b) for a, b, _, _ in rec) * 1e-3
    fl = sum(f for _, _, f, _ in rec)
    agg = {}
    for a, b, f, shape in rec:
        d = agg.setdefault(shape, [0.0, 0.0, 0])
        d[0] += a.elapsed_time(b) * 1e3; d[1] += f; d[2] += 1
    if os.environ.get("VILCO_GEMM_TABLE"):
        for shape, (us, f, n) in sorted(agg.items(), key=lambda x: -x[1][0]):
            print(f"  gemm M{shape[0]:6d} N{shape[1]:5d} K{shape[2]:5d} taps{shape[3]} Z{shape[4]:4d} {shape[5]:9s} x{n:3d}  {us:9.1f} us  {f / us / 1e6:7.1f} TFLOP/s", file=sys.stderr)
    top_shape, (top_us, top_f, top_n) = max(agg.items(), key=lambda x: x[1][0])
    gemm_roofline.top = {"shape": dict(zip(("M", "N", "K", "taps", "Z", "out"), top_shape)), "launches": top_n,
                         "us_per_launch": top_us / top_n, "tflops": top_f / top_us / 1e6, "flop_per_launch": top_f / top_n}
    return fl, t, len(rec)


def run_train_leg(args, model, rank, world, dist, barrier, local):
    """Training half of the metric: K timed iterations of trainer.Trainer.step (device-resident inputs), then K more with
    pinned host inputs and a host read of the loss every step (e2e)."""
    from vilco_b200 import lib as L
    from vilco_b200.dist import max_over_ranks
    from vilco_b200.trainer import Trainer, broadcast_parameters, make_optimizer
    Bt = args.train_batch
    torch.cuda.empty_cache()
    model.train()
    broadcast_parameters(model)
    opt = make_optimizer(model, {"type": "AdamW", "learning_rate": 1e-4, "weight_decay": 0.05}, flat=True)
    tr = Trainer(model, opt, clip_grad_l2norm=1.0, overlap=os.environ.get("VILCO_OVERLAP", "1") == "1")
    host_sets = [synth_videos(Bt, seed=500 + rank * 10 + i, pin=True) for i in range(2)]
    dev_set = [dict(v) for v in host_sets[0]]
    for v in dev_set:
        v["feats"] = v["feats"].cuda()
    for _ in range(args.warmup):
        tr.step(dev_set)
    barrier()
    # (NVML queries contend with the ~2500 kernel launches of a training step for the driver: poll slowly here)
    sampler = ClockSampler(local, enabled=rank == 0, period=0.4)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = L.launch_count()
    e0.record()
    for _ in range(args.steps):
        tr.step(dev_set)
    e1.record()
    barrier()
    t_dev = e0.elapsed_time(e1) * 1e-3
    launches = (L.launch_count() - n0) // args.steps
    # end to end: pinned host batches; the upload of batch i+1 is started (Trainer.stage, copy stream) before step i runs
    nxt = tr.stage(host_sets[0])
    for i in range(args.warmup):
        cur, nxt = nxt, tr.stage(host_sets[(i + 1) % 2])
        float(tr.step(cur)["final_loss"].detach())
    barrier()
    t0 = time.perf_counter()
    last = 0.0
    for i in range(args.steps):
        cur, nxt = nxt, tr.stage(host_sets[(i + 1) % 2])
        last = float(tr.step(cur)["final_loss"].detach())      # D2H read of the loss every step
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    # the reference's own batch size (train_cfg.batch_size 2 per GPU): launch-bound, reported for context
    small = dev_set[:2]
    for _ in range(2):
        tr.step(small)
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(5):
        tr.step(small)
    s1.record()
    barrier()
    t_small = s0.elapsed_time(s1) * 1e-3
    t_dev, t_e2e, t_small = max_over_ranks([t_dev, t_e2e, t_small], device="cuda")
    mem = torch.cuda.max_memory_allocated() / 2 ** 30
    model.eval()
    h2d = sum(v["feats"].numel() * 4 + v["prompt_feature"].numel() * 4 for v in host_sets[0])
    return {"metric": "mq_train_videos_per_s", "value": world * Bt * args.steps / t_dev, "unit": "videos/s",
            "ms_per_step": 1e3 * t_dev / args.steps, "videos_per_step_per_gpu": Bt, "gpu_launches_per_step": int(launches),
            "e2e": {"value": world * Bt * args.steps / t_e2e, "unit": "videos/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "step": "zero_grad, forward (dropout 0.1, drop-path 0.1, XLNet dropout 0.1), focal + DIoU + label-involved loss, "
                    "hand-written backward, " + (f"NCCL all-reduce of the flat gradient buffer ({tr.grad_comm} on the wire, fp32 master), " if world > 1 else "") +
                    "clip_grad_norm 1.0, fused flat AdamW (lr 1e-4, wd 0.05)",
            "grad_bytes_allreduced_per_step": int((2 if tr.grad_comm == "bf16" else 4) * sum(b - a for a, b in opt.live_ranges())) if world > 1 else 0,
            "allreduce": ("bucketed (128 MB), launched as the backward completes each bucket" if tr.overlap else "one call after the backward") if world > 1 else None,
            "live_parameters": int(sum(b - a for a, b in opt.live_ranges())), "all_parameters": int(opt.n), "last_loss": last, "peak_mem_gib": mem,
            "batch2": {"value": world * 2 * 5 / t_small, "unit": "videos/s", "ms_per_step": 1e3 * t_small / 5,
                       "note": "same step at the reference's 2 clips per GPU: bound by the ~1500 kernel launches issued from Python"},
            "clocks": sampler.summary()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64,
                    help="clips per step per GPU (64: 1.9 k videos/s; 32, the round-1 setting: 1.8 k; batch-1 latency is reported beside it)")
    ap.add_argument("--train-batch", type=int, default=64,
                    help="clips per training step per GPU (0 = skip the training leg); 64 clips keep 112 of the 180 GB busy and run 16 %% "
                         "faster per clip than 32 (launch gaps and the short pyramid levels amortise); the reference's 2 is reported beside it")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--metric", default="infer", choices=["infer", "train"], help="which half of the metric is the line's `value`")
    ap.add_argument("--precision", default=None, choices=[None, "mixed", "fp16x3", "fp16", "bf16x3", "bf16"])
    ap.add_argument("--ref-videos", type=int, default=None,
                    help="clips per step of the reference arm (default 2) / of the cpu_baseline sample of our arm (default 12)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        args.ref_videos = args.ref_videos or 2
        return run_reference_arm(args)
    args.ref_videos = args.ref_videos or 12

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from vilco_b200 import lib as L
    from vilco_b200 import ops
    if args.precision:
        ops.set_precision(args.precision)
    model = build_model().cuda().eval()
    B = args.batch
    g = model.make_eval_graph(B)
    vids = synth_videos(B, seed=rank, pin=True)
    g.load_inputs(vids)
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (graph replay) ----
    for _ in range(args.warmup):
        g.replay()
    barrier()
    sampler = ClockSampler(local, enabled=rank == 0)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        g.replay()
    e1.record()
    barrier()
    t_dev = e0.elapsed_time(e1) * 1e-3
    # ---- end to end through the public API: pinned host inputs, H2D, D2H of the detections ----
    sets = [synth_videos(B, seed=100 + rank * 10 + i, pin=True) for i in range(3)]
    for i in range(args.warmup):
        g.run(sets[i % 3])
    barrier()
    for _ in g.infer_stream([sets[i % 3] for i in range(args.warmup)]):
        pass
    barrier()
    t0 = time.perf_counter()
    n_res = 0
    for res in g.infer_stream([sets[i % 3] for i in range(args.steps)]):   # public streaming API: H2D(i+1) overlaps step i
        n_res += len(res)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    assert n_res == B * args.steps
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    from vilco_b200.dist import max_over_ranks
    t_dev, t_e2e = max_over_ranks([t_dev, t_e2e], device="cuda")
    h2d = sum(v["feats"].numel() * 4 + v["prompt_feature"].numel() * 4 for v in sets[0]) + B * (1024 + 128 + 1) * 4
    d2h = B * (200 * (2 + 1) * 4 + 200 * 8 + 4)
    hbm, tf_burst, tf_sus, how = peaks()
    fl, tg, nl = gemm_roofline(g, model, B)
    top = gemm_roofline.top
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")   # dram bytes per launch from the committed ncu --set full capture
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        key = "x".join(str(top["shape"][k]) for k in ("M", "N", "K", "taps", "Z")) + ":" + ops.precision()
        traffic = tj.get(key)
    roof = {"bound": "tensor", "kernel": "vilco::gemm_tc_kernel<256,2> (CTA-pair tcgen05 GEMM): the shape with the largest share of the step",
            "shape": top["shape"], "launches_per_step": top["launches"], "us_per_launch": top["us_per_launch"],
            "achieved": top["tflops"], "peak": tf_sus, "unit": "TFLOP/s", "frac": top["tflops"] / tf_sus,
            "traffic": traffic, "peak_source": f"bf16_tflops_sustained of {how} (cuBLAS bf16 = the same tensor rate as fp16)",
            "algorithmic_flop_per_launch": top["flop_per_launch"],
            "all_gemm_launches": {"launches": nl, "achieved": fl / tg / 1e12, "share_of_step": tg / (t_dev / args.steps)}}
    # latency of the reference's own evaluation mode (one clip per call)
    g1 = model.make_eval_graph(1)
    g1.load_inputs(vids[:1])
    for _ in range(3):
        g1.replay()
    torch.cuda.synchronize()
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record()
    for _ in range(10):
        g1.replay()
    l1.record()
    torch.cuda.synchronize()
    lat_b1 = l0.elapsed_time(l1) / 10
    launches_inf = int(g.launches * args.steps)
    verify = None
    if rank == 0 and not args.no_verify:
        try:
            verify = verify_against_oracle(model, vids[0])
        except Exception as e:
            verify = {"failed": repr(e)[:300]}
    del g, g1          # release the CUDA-graph memory pools before the training leg
    import gc
    gc.collect()
    train = run_train_leg(args, model, rank, world, dist, barrier, local) if args.train_batch > 0 else None
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    infer_value, infer_ms = world * B * args.steps / t_dev, 1e3 * t_dev / args.steps
    e2e_inf = {"value": world * B * args.steps / t_e2e, "unit": "videos/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}
    head = {"metric": "mq_infer_videos_per_s", "value": infer_value, "unit": "videos/s", "ms_per_step": infer_ms, "e2e": e2e_inf,
            "gpu_launches": launches_inf}
    if args.metric == "train" and train is not None:
        head = {"metric": "mq_train_videos_per_s", "value": train["value"], "unit": "videos/s", "ms_per_step": train["ms_per_step"],
                "e2e": train["e2e"], "gpu_launches": int(train["gpu_launches_per_step"] * args.steps)}
    line = {
        "metric": head["metric"], "value": head["value"], "unit": head["unit"], "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f16" if ops.ACT_DTYPE == torch.float16 else "bf16", "data": "synthetic",
        # both halves of BASELINE.json's metric, compact and first (the verbose objects follow)
        "infer_value": infer_value, "infer_ms_per_step": infer_ms, "infer_e2e_value": e2e_inf["value"],
        "train_value": train["value"] if train else None, "train_ms_per_step": train["ms_per_step"] if train else None,
        "train_e2e_value": train["e2e"]["value"] if train else None,
        "train_batch2_value": train["batch2"]["value"] if train else None,
        "train_batch2_ms_per_step": train["batch2"]["ms_per_step"] if train else None,
        "e2e": head["e2e"], "gpu_launches": head["gpu_launches"],
        "clocks": sampler.summary(),
        "config": bench_config(args),
        "roofline": roof,
        "verify": verify,
        "latency_b1_ms": lat_b1,
    }
    extras = {"train": train}
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_leg(args.ref_videos)
        try:
            eg = {}
            for tf32 in (False, True):
                i_r, t_r, kind = eager_gpu_rates(tf32)
                eg["tf32" if tf32 else "fp32"] = {"infer_videos_per_s": i_r, "train_videos_per_s": t_r}
            eg["kind"] = kind
            eg["what"] = ("the reference's own modules as eager PyTorch on this B200 (torch ops -> cuDNN / cuBLAS), evaluation one "
                          "clip per call + its host soft-NMS, training batch 2 (forward, backward, clip, AdamW)")
            line["eager_gpu_baseline"] = eg
        except Exception as e:  # a baseline must never take the bench down
            line["eager_gpu_baseline"] = {"unavailable": repr(e)[:300]}
        torch.cuda.empty_cache()
        if world == 1:
            try:
                extras["next_rows"] = next_rows()
            except Exception as e:  # extras must never take the bench line down
                extras["next_rows"] = {"unavailable": repr(e)[:200]}
    line["train"] = train
    if "next_rows" in extras:
        line["next_rows"] = extras["next_rows"]
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
