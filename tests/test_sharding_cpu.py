"""CPU, world_size 2, gloo: the N>1 host logic of bench.py / multi-GPU inference — images are sharded by rank with
no collective on the data path; only the barrier and the max-over-ranks of the timing use the process group."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def shard(global_batch: int, rank: int, world: int):
    """Contiguous split of the global batch."""
    per = global_batch // world
    extra = global_batch % world
    start = rank * per + min(rank, extra)
    return start, start + per + (1 if rank < extra else 0)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(9)                       # one global draw, identical on every rank
    noise = torch.randn(10, 3, 4, 4, generator=g)
    lo, hi = shard(10, rank, world)
    mine = noise[lo:hi]
    # a rank's "result" depends only on its own shard (stand-in for enhance): no data-path collective needed
    result = mine * 2 + 1
    ms = torch.tensor([5.0 + rank])                            # per-rank device time
    dist.barrier()
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, result))
    if rank == 0:
        full = torch.cat([r for _, _, r in sorted(gathered, key=lambda t: t[0])])
        q.put((ms.item(), torch.equal(full, noise * 2 + 1), [(a, b) for a, b, _ in gathered]))
    dist.destroy_process_group()


def test_shard_covers_batch_exactly():
    for gb in (1, 7, 64, 256):
        for world in (1, 2, 4, 8):
            spans = [shard(gb, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))


def test_two_rank_gloo_sharding():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ms, same, spans = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ms == 6.0                      # max over ranks
    assert same                           # sharded result == single-process result on the same global draw
    assert spans == [(0, 5), (5, 10)]
