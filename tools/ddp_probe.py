"""2-GPU probe (torchrun): bucketed / overlapped gradient all-reduce == single all-reduce, replicas stay identical.
torchrun --nproc-per-node 2 tools/ddp_probe.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import gen_golden as GG  # noqa: E402
from oracle import params as PR  # noqa: E402
import util  # noqa: E402
from vilco_b200.dist import shard_indices  # noqa: E402
from vilco_b200.trainer import Trainer, broadcast_parameters, make_optimizer  # noqa: E402


def run(overlap, rank, world, cfg, videos, steps=3):
    model, P = util.build_pair(cfg, 0)
    model.eval()
    model.loss_normalizer = cfg.init_loss_norm
    broadcast_parameters(model)
    opt = make_optimizer(model, {"type": "AdamW", "learning_rate": 0.0, "weight_decay": 0.0}, flat=True)   # lr 0: same weights
    tr = Trainer(model, opt, clip_grad_l2norm=1.0, overlap=overlap)                                         # every step
    tr.BUCKET = 1 << 20
    tr.keep_grad = True
    mine = [videos[i] for i in shard_indices(len(videos), rank, world)]
    losses = [float(tr.step(mine)["final_loss"].detach()) for _ in range(steps)]
    names = {id(p): k for k, p in model.named_parameters()}
    seeded = torch.cat([tr.last_grad[o:o + k] for g in opt.param_groups for p in g["params"]
                        for (o, k) in [opt.slots[id(p)]] if names[id(p)] in P])      # dead weights are random per build
    return seeded, losses, tr.plan, opt.flat_p.clone()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    cfg = GG.small_cfg()
    videos = PR.synth_video_list(cfg, 4, seed=3, lens=[128, 100, 90, 128], text_lens=[40, 57, 33, 64], n_gt=[3, 2, 4, 1])
    g0, l0, _, _ = run(False, rank, world, cfg, videos)
    g1, l1, plan, p1 = run(True, rank, world, cfg, videos)
    n_early = sum(len(v) for v in plan[1].values())
    same_mode = float((g0 - g1).abs().max() / g0.abs().max())
    other = [torch.empty_like(g1) for _ in range(world)]
    dist.all_gather(other, g1)
    across = max(float((o - g1).abs().max()) for o in other)
    if rank == 0:
        print(f"losses single {l0}\nlosses overlap {l1}")
        print(f"buckets launched during backward: {n_early}, at the end: {len(plan[2])}, nodes {plan[0]}")
        print(f"averaged gradient, single all-reduce vs bucketed / overlapped: rel. max diff {same_mode:.3e}; "
              f"max difference across ranks = {across:.3e}")
        print("OK" if same_mode < 1e-4 and across == 0.0 and n_early > len(plan[2]) else "BAD")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
