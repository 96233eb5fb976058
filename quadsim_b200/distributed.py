"""Multi-GPU plumbing: envs shard as independent index ranges, one rank per GPU, with NO collective on the step path
(SURVEY.md section 8e).  The only exchange is one all-reduce of the 5-element episode-statistics vector per rollout."""
import os

import torch
import torch.distributed as dist


_orig_affinity = None


def restore_affinity():
    """Undo bind_to_gpu_numa_node (e.g. before a CPU-side job that should use every core)."""
    if _orig_affinity is not None:
        os.sched_setaffinity(0, _orig_affinity)


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(device_index, sysfs="/sys"):
    """NUMA node of a CUDA device from its PCI address (sysfs), or None when the box does not say (single node, VM)."""
    try:
        p = torch.cuda.get_device_properties(device_index)
        addr = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open(os.path.join(sysfs, "bus/pci/devices", addr, "numa_node")) as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_numa_node(device_index, sysfs="/sys"):
    """Pin this process to the CPUs of its GPU's NUMA node so that the pinned staging buffers of the host path
    (QuadVecEnv.step_host) are first-touched on the socket the GPU's PCIe root hangs off.  With one rank per GPU and
    69 B per env-step crossing PCIe in each rank, un-bound ranks all stage through whichever node they happen to run on
    and the host path stops scaling at ~2 GPUs (measured: 0.86e9 env-steps/s at 1 GPU, 1.6e9 at 2, 4 and 8).
    Returns the node bound to, or None (nothing known / QUADSIM_B200_NO_NUMA_BIND=1 / not Linux)."""
    if os.environ.get("QUADSIM_B200_NO_NUMA_BIND") == "1" or not hasattr(os, "sched_setaffinity"):
        return None
    node = gpu_numa_node(device_index, sysfs)
    if node is None:
        return None
    try:
        with open(os.path.join(sysfs, "devices/system/node/node%d/cpulist" % node)) as f:
            cpus = _parse_cpulist(f.read())
        global _orig_affinity
        allowed = os.sched_getaffinity(0)
        cpus &= allowed  # stay inside the cgroup / taskset we were given
        if not cpus:
            return None
        if _orig_affinity is None:
            _orig_affinity = allowed
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's RANK/LOCAL_RANK/WORLD_SIZE/MASTER_* (no-op for a single process)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
        bind_to_gpu_numa_node(local_rank)
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def shard_range(total_envs, rank=None, world=None):
    """Global env ids [lo, hi) owned by `rank`: contiguous, near-equal blocks; Philox uses the global id, so a
    1-GPU run and an 8-GPU run reproduce the same per-env trajectories."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    base, rem = divmod(int(total_envs), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def default_env_id0(num_envs, device=None):
    """Global id of this rank's env 0 when the caller gives none: the exclusive prefix sum of every rank's env count.
    With equal shards this is rank * num_envs; with unequal ones (shard_range of a non-divisible total: 10 envs over 3 ranks
    = 4/3/3) the ranks' Philox counter ranges still tile [0, total) without overlap or gap.  One tiny all_gather at
    construction."""
    if not dist.is_initialized():
        return 0
    dev = device if dist.get_backend() == "nccl" else "cpu"
    mine = torch.tensor([int(num_envs)], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(counts, mine)
    return int(sum(int(c.item()) for c in counts[:dist.get_rank()]))


def all_reduce_stats(vals, device=None):
    """SUM over ranks of the stats vector (float64[5]).  NCCL needs a device tensor; gloo reduces on the host."""
    if not dist.is_initialized():
        return vals
    if dist.get_backend() == "nccl":
        t = vals.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu()
    t = vals.clone()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def all_reduce_mean_grads(params):
    """Average the gradients of `params` over ranks with ONE all-reduce: the words travel as a single flat bucket (the PPO2
    policy has 11 parameter tensors, ~35 k words; 11 small collectives per minibatch cost more than the gradient kernel).
    Each p.grad is replaced by a view into the reduced bucket.  No-op without an initialised process group."""
    if not dist.is_initialized():
        return
    with_grad = [p for p in params if p.grad is not None]
    if not with_grad:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in with_grad])
    dist.all_reduce(flat)
    flat /= dist.get_world_size()
    off = 0
    for p in with_grad:
        p.grad = flat[off:off + p.numel()].view_as(p)
        off += p.numel()


def max_over_ranks(x, device=None):
    """MAX-reduce a python float over ranks (bench timing)."""
    if not dist.is_initialized():
        return float(x)
    dev = device if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
