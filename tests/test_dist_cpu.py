"""world_size-2 gloo test (CPU) of the N>1 host logic: shard by video, no data-path collective, max-over-ranks timing,
ordered gather of results."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vilco_b200.dist import gather_results, max_over_ranks, shard_indices
    mine = shard_indices(7, rank, world)
    res = [{"video_id": f"v{i}", "scores": torch.full((2,), float(i))} for i in mine]
    allr = gather_results(res)
    tmax = max_over_ranks([1.0 + rank, 5.0 - rank])
    q.put((rank, mine, [r["video_id"] for r in allr], tmax))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_gather_and_max_over_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out[0][1] == [0, 2, 4, 6] and out[1][1] == [1, 3, 5]
    for _, _, ids, tmax in out:
        assert ids == [f"v{i}" for i in range(7)]      # original order restored on every rank
        assert tmax == [2.0, 5.0]


def test_single_process_is_identity():
    from vilco_b200.dist import gather_results, max_over_ranks, shard_indices
    assert shard_indices(5, 0, 1) == [0, 1, 2, 3, 4]
    assert max_over_ranks([3.0]) == [3.0]
    assert gather_results([1, 2]) == [1, 2]


class _Toy(torch.nn.Module):
    """Stand-in with the model's training interface (forward(video_list, ...) -> {'final_loss': ...}) so that the
    data-parallel step of vilco_b200.trainer can be exercised on CPU / gloo."""
    use_adapt = False

    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.a = torch.nn.Linear(4, 3)
        self.b = torch.nn.Parameter(torch.ones(3))

    def forward(self, video_list, task_id=0, prev_out_cls_logits=None):
        x = torch.stack([v["feats"] for v in video_list])
        return {"final_loss": ((self.a(x) * self.b) ** 2).mean()}


def _train_worker(rank, world, port, q, comm="fp32"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vilco_b200.dist import shard_indices
    from vilco_b200.trainer import Trainer, broadcast_parameters
    model = _Toy()
    with torch.no_grad():
        model.b.add_(rank)            # deliberately different before the broadcast
    broadcast_parameters(model)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    tr = Trainer(model, opt, clip_grad_l2norm=1.0, grad_comm=comm)
    g = torch.Generator().manual_seed(5)
    videos = [{"feats": torch.randn(4, generator=g)} for _ in range(6)]
    for _ in range(3):
        tr.step([videos[i] for i in shard_indices(len(videos), rank, world)])
    q.put((rank, [p.detach().reshape(-1).tolist() for p in model.parameters()], tr.grads.attached()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("comm,atol", [("fp32", 1e-6), ("bf16", 2e-3)])
def test_data_parallel_training_step_matches_single_process(comm, atol):
    """2 ranks x 3 videos with gradient all-reduce == 1 process x 6 videos (equal shard sizes, mean loss): exactly with fp32
    buckets, to bf16 rounding of the exchanged gradients with the default bf16-on-the-wire buckets."""
    from vilco_b200.trainer import Trainer
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q, comm)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted((q.get(timeout=120) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    model = _Toy()
    tr = Trainer(model, torch.optim.SGD(model.parameters(), lr=0.1), clip_grad_l2norm=1.0)
    g = torch.Generator().manual_seed(5)
    videos = [{"feats": torch.randn(4, generator=g)} for _ in range(6)]
    for _ in range(3):
        tr.step(videos)
    for rank, params, attached in out:
        assert attached
        for a, b in zip(params, model.parameters()):
            assert torch.allclose(torch.tensor(a), b.detach().reshape(-1), atol=atol), rank
