"""GPU probe: time one data-parallel training iteration of the full MQ config.  usage: python tools/train_bench.py [B] [steps]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from vilco_b200 import lib as L  # noqa: E402
from vilco_b200.trainer import Trainer, make_optimizer  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    prof = os.environ.get("TRAIN_PROF")
    if os.environ.get("VILCO_CFG") == "vilco":      # mq_vilco.yaml: L2P prompts + temporal adapters (+EMA) + narration SSL
        import numpy as np
        from vilco_b200.config import mq_model_kwargs
        from vilco_b200.modeling import make_meta_arch
        kw = mq_model_kwargs(num_classes=22)
        kw["cl_cfg"].update(name="l2p", memory_size=1010, prompt_pool=True, pool_size=10, topk=4, length=20, embed_dim=768,
                            narration_ssl=True, narration_dim=512, ssl_factor=0.01, use_adapt=True, adapt_blocks=[0, 1, 2, 3, 4])
        torch.manual_seed(0)
        model = make_meta_arch("LocPointTransformer", **kw).cuda().train()
    else:
        model = bench.build_model().cuda().train()
    opt = make_optimizer(model, {"type": "AdamW", "learning_rate": 1e-4, "weight_decay": 0.05}, flat=os.environ.get("FLAT", "1") == "1")
    tr = Trainer(model, opt, clip_grad_l2norm=1.0)
    vids = bench.synth_videos(B, seed=0)
    for i, v in enumerate(vids):
        v["feats"] = v["feats"].cuda()
        if os.environ.get("VILCO_CFG") == "vilco":
            rs = np.random.RandomState(i)
            v["narration_feats"] = torch.from_numpy(rs.standard_normal((512, int(rs.randint(1, 17)))).astype(np.float32))
            v["narration_mask"] = float(rs.rand() < 0.7)
    for i in range(2):
        lo = tr.step(vids)
    torch.cuda.synchronize()
    print("losses", {k: float(v) for k, v in lo.items()}, "mem GB", torch.cuda.max_memory_allocated() / 2**30)
    if os.environ.get("GEMM_TABLE"):
        rec, orig = [], L.gemm

        def timed(A, Bm, D, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = orig(A, Bm, D, **kw); e1.record()
            Z = kw.get("Z", (1, 1))
            rec.append((e0, e1, 2.0 * kw["M"] * kw["N"] * kw["K"] * kw.get("taps", 1) * Z[0] * Z[1],
                        (kw["M"], kw["N"], kw["K"], kw.get("taps", 1), Z[0] * Z[1], f"{kw.get('a_major', 0)}{kw.get('b_major', 0)}p{A.shape[0]}{Bm.shape[0]}", str(D.dtype)[6:])))
            return r
        L.gemm = timed
        tr.step(vids)
        torch.cuda.synchronize()
        L.gemm = orig
        agg = {}
        for a, b, f, shape in rec:
            d = agg.setdefault(shape, [0.0, 0.0, 0])
            d[0] += a.elapsed_time(b) * 1e3; d[1] += f; d[2] += 1
        tot = sum(v[0] for v in agg.values())
        print(f"GEMM total {tot / 1e3:.1f} ms, {sum(v[1] for v in agg.values()) / tot / 1e6:.1f} TFLOP/s algorithmic, {len(rec)} launches")
        for shape, (us, f, n) in sorted(agg.items(), key=lambda x: -x[1][0])[:28]:
            print(f"  M{shape[0]:6d} N{shape[1]:5d} K{shape[2]:6d} taps{shape[3]} Z{shape[4]:4d} bmaj{shape[5]} {shape[6]:9s} x{n:3d}  {us:9.1f} us  {f / us / 1e6:7.1f} TFLOP/s")
        return
    if os.environ.get("CPROF"):
        import cProfile
        import pstats
        pr = cProfile.Profile()
        pr.enable()
        tr.step(vids)
        torch.cuda.synchronize()
        pr.disable()
        pstats.Stats(pr).sort_stats(os.environ.get("CPROF_SORT", "cumulative")).print_stats(45)
        return
    n0 = L.launch_count()
    t0 = time.perf_counter()
    if prof:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as p:
            tr.step(vids)
            torch.cuda.synchronize()
        print(p.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
        return
    cpu = 0.0
    for _ in range(steps):
        c0 = time.perf_counter()
        lo = tr.step(vids)
        cpu += time.perf_counter() - c0          # host time to ISSUE the step (no synchronisation inside tr.step except the loss sums)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    print(f"host issue time {cpu / steps * 1e3:.1f} ms/step")
    print(f"B={B}: {dt * 1e3:.1f} ms/step, {B / dt:.1f} train videos/s, our launches/step {(L.launch_count() - n0) / steps:.0f}")


if __name__ == "__main__":
    main()
