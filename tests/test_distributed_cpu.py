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
    # unequal shards (5 envs over 2 ranks = 3 / 2): the default global id of a rank's env 0 is the exclusive prefix sum of the
    # ranks' env counts -- rank * num_envs would give rank 1 id0 = 2 and overlap rank 0's env 2
    ulo, uhi = D.shard_range(5, r, w)
    assert D.default_env_id0(uhi - ulo) == ulo == (0, 3)[r]
    stats = torch.tensor([1.5 * (r + 1), 10.0 * (r + 1), float(r + 1), 0.0, float(hi - lo)], dtype=torch.float64)
    red = D.all_reduce_stats(stats)
    mx = D.max_over_ranks(0.1 * (r + 1))
    # gradient averaging of the PPO2 caller: one flat bucket, parameters without a gradient are skipped
    torch.manual_seed(0)
    lin, unused = torch.nn.Linear(3, 2), torch.nn.Linear(2, 2)
    lin.weight.grad = torch.full((2, 3), float(r + 1))
    lin.bias.grad = torch.tensor([10.0 * (r + 1), -1.0])
    D.all_reduce_mean_grads(list(lin.parameters()) + list(unused.parameters()))
    assert unused.weight.grad is None
    assert torch.equal(lin.weight.grad, torch.full((2, 3), 1.5)) and torch.equal(lin.bias.grad, torch.tensor([15.0, -1.0]))
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


def test_numa_binding_helpers(tmp_path, monkeypatch):
    """bind_to_gpu_numa_node reads the GPU's node and that node's cpulist from sysfs; exercised on a fake sysfs tree."""
    import os
    import types

    import torch
    from quadsim_b200 import distributed as D
    assert D._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11} and D._parse_cpulist("") == set()
    props = types.SimpleNamespace(pci_domain_id=0, pci_bus_id=0x1b, pci_device_id=0)
    monkeypatch.setattr(torch.cuda, "get_device_properties", lambda i: props)
    dev = tmp_path / "bus/pci/devices/0000:1b:00.0"
    dev.mkdir(parents=True)
    (dev / "numa_node").write_text("1\n")
    node = tmp_path / "devices/system/node/node1"
    node.mkdir(parents=True)
    allowed = sorted(os.sched_getaffinity(0))
    (node / "cpulist").write_text("%d\n" % allowed[0])
    assert D.gpu_numa_node(0, sysfs=str(tmp_path)) == 1
    before = os.sched_getaffinity(0)
    try:
        assert D.bind_to_gpu_numa_node(0, sysfs=str(tmp_path)) == 1 and os.sched_getaffinity(0) == {allowed[0]}
    finally:
        os.sched_setaffinity(0, before)
    (dev / "numa_node").write_text("-1\n")  # VMs report -1: nothing to bind to
    assert D.gpu_numa_node(0, sysfs=str(tmp_path)) is None and D.bind_to_gpu_numa_node(0, sysfs=str(tmp_path)) is None
    monkeypatch.setenv("QUADSIM_B200_NO_NUMA_BIND", "1")
    (dev / "numa_node").write_text("1\n")
    assert D.bind_to_gpu_numa_node(0, sysfs=str(tmp_path)) is None
