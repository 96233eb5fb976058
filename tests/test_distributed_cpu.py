"""world_size-2 gloo checks of the multi-GPU host logic (env sharding by global id, the per-rollout stats all-reduce).
The step path itself has no collective."""
import os
import socket

import torch
import torch.multiprocessing as mp

from quadsim_b200 import distributed as D


def test_shard_range_partitions_exactly():
    for total in (10, 1 << 20, 12345):
        for world in (1, 2, 3, 8):
            parts = [D.shard_range(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, lr, w = D.init_from_env(backend="gloo")
    lo, hi = D.shard_range(1000, r, w)
    stats = torch.tensor([1.5 * (r + 1), 10.0 * (r + 1), float(r + 1), 0.0, float(hi - lo)], dtype=torch.float64)
    red = D.all_reduce_stats(stats)
    mx = D.max_over_ranks(0.1 * (r + 1))
    q.put((r, lo, hi, red.tolist(), mx))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_stats_allreduce_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert [o[1:3] for o in out] == [(0, 500), (500, 1000)]
    for o in out:
        assert o[3] == [4.5, 30.0, 3.0, 0.0, 1000.0] and abs(o[4] - 0.2) < 1e-12
