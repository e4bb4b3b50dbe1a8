"""world_size-2 gloo test (CPU) of the N>1 host logic: shard by video, no data-path collective, max-over-ranks timing,
ordered gather of results."""
import os

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
