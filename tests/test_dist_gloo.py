"""CPU, world_size 2, gloo: the N>1 host logic (sharding, gradient all-reduce, result gather, max-over-ranks)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from nglod_b200 import dist as nd
    r, w, _ = nd.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    # rays of a 20x6 x-major frame, cut on column boundaries
    n, h = 120, 6
    full = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3)
    s, e = nd.shard_range(n, rank, world, align=h)
    assert s % h == 0
    local = full[s:e] * 2.0                         # "trace" the shard
    out = nd.gather_shards(local, n, rank, world, dst=0, align=h)
    if rank == 0:
        assert torch.equal(out, full * 2.0)
    else:
        assert out is None
    # uneven shards (7 items over 2 ranks)
    s, e = nd.shard_range(7, rank, world)
    out = nd.gather_shards(torch.arange(s, e).float().unsqueeze(1), 7, rank, world)
    if rank == 0:
        assert out[:, 0].tolist() == [0, 1, 2, 3, 4, 5, 6]
    # a frame rendered in interleaved column strips, gathered back into image order on rank 0
    ncols, h2 = 37, 5
    strips = [nd.interleaved_strips(ncols, r, world, strips_per_rank=3) for r in range(world)]
    assert sorted(c for s_ in strips for c0, c1 in s_ for c in range(c0, c1)) == list(range(ncols))
    frame = torch.arange(ncols * h2, dtype=torch.float32).reshape(-1, 1)
    mine = torch.cat([frame[c0 * h2:c1 * h2] for c0, c1 in strips[rank]]) + 0.5
    got = nd.gather_strips(mine, strips, h2, rank, world, dst=0)
    if rank == 0:
        assert torch.equal(got, frame + 0.5)
    else:
        assert got is None
    # ... and with equal strips (the single-buffer path): 40 columns = 2 ranks x 4 strips x 5
    ncols = 40
    strips = [nd.interleaved_strips(ncols, r, world, strips_per_rank=4) for r in range(world)]
    frame = torch.arange(ncols * h2 * 3, dtype=torch.float32).reshape(-1, 3)
    mine = torch.cat([frame[c0 * h2:c1 * h2] for c0, c1 in strips[rank]])
    got = nd.gather_strips(mine, strips, h2, rank, world, dst=0)
    if rank == 0:
        assert torch.equal(got, frame)
    else:
        assert got is None
    # data-parallel gradient: each rank contributes its slice's gradient, scaled by the GLOBAL batch
    g = torch.full((10,), float(rank + 1))
    nd.allreduce_sum_(g)
    assert g.tolist() == [3.0] * 10
    assert nd.max_over_ranks(1.0 + rank, "cpu") == 2.0
    dist.barrier()
    dist.destroy_process_group()
    q.put(rank)


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=5) for _ in range(2)) == [0, 1]
