"""2-GPU probe (torchrun): the reference's own loop — model wrapped in torch DistributedDataParallel(find_unused_parameters=True)
like MQ/train_cl.py, optimizer.zero_grad(set_to_none=True), final_loss.backward(), clip_grad_norm_, torch AdamW — on the CUDA
training path.  Checks that DDP's reducer saw the hand-written backward's gradients: after steps on DIFFERENT data the
replicas are still identical, and the averaged gradient equals the mean of the two ranks' local gradients.
torchrun --nproc-per-node 2 tools/ddp_native_probe.py"""
import os
import sys

import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import gen_golden as GG  # noqa: E402
from oracle import params as PR  # noqa: E402
import util  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = GG.small_cfg()
    model, P = util.build_pair(cfg, 0)
    model.eval()                                   # deterministic (no dropout), is_training=True selects the loss path
    model.loss_normalizer_momentum = 1.0
    videos = PR.synth_video_list(cfg, 4, seed=3, lens=[128, 100, 90, 128], text_lens=[40, 57, 33, 64], n_gt=[3, 2, 4, 1])
    mine = videos[rank::world]
    # local (un-synchronised) gradient of this rank, for the check below
    model.zero_grad(set_to_none=True)
    model.loss_normalizer = cfg.init_loss_norm
    model(mine)["final_loss"].backward()
    keys = [k for k, p in model.named_parameters() if k in P and p.grad is not None]
    local_g = torch.cat([dict(model.named_parameters())[k].grad.reshape(-1) for k in keys]).clone()
    ddp = DDP(model, device_ids=[local], find_unused_parameters=True)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    for it in range(3):
        opt.zero_grad(set_to_none=True)
        model.loss_normalizer = cfg.init_loss_norm
        losses = ddp(mine)
        losses["final_loss"].backward()
        if it == 0:
            avg_g = torch.cat([dict(model.named_parameters())[k].grad.reshape(-1) for k in keys]).clone()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
    both = [torch.empty_like(local_g) for _ in range(world)]
    dist.all_gather(both, local_g)
    want = sum(both) / world
    err = float((avg_g - want).abs().max() / want.abs().max())
    flat = torch.cat([p.detach().reshape(-1) for k, p in model.named_parameters() if k in P])
    others = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(others, flat)
    drift = max(float((o - flat).abs().max()) for o in others)
    if rank == 0:
        print(f"DDP-averaged gradient vs mean of local gradients: rel. max diff {err:.3e}; replica drift after 3 steps {drift:.3e}")
        print("OK" if err < 1e-4 and drift == 0.0 else "BAD")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
