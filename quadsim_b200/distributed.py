"""Multi-GPU plumbing: envs shard as independent index ranges, one rank per GPU, with NO collective on the step path
(SURVEY.md section 8e).  The only exchange is one all-reduce of the 5-element episode-statistics vector per rollout."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's RANK/LOCAL_RANK/WORLD_SIZE/MASTER_* (no-op for a single process)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
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


def max_over_ranks(x, device=None):
    """MAX-reduce a python float over ranks (bench timing)."""
    if not dist.is_initialized():
        return float(x)
    dev = device if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
