#!/usr/bin/env python
"""Multi-GPU check of the PPO2 caller (SURVEY.md section 8 rows e / f1), one process per GPU over NCCL:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dist_ppo_check.py

Asserts, on every rank: (1) the ranks' default global env ids partition [0, total) (distributed.default_env_id0 over NCCL),
so their Philox streams differ -- the first observations of rank 0 and rank 1 are not equal; (2) after `--iters` learn
iterations with the hand-written update (gradient kernels -> ONE in-place all-reduce of the gradient bucket -> qs_adam_step)
every rank holds bit-identical weights, and they differ from the initial ones; (3) the episode statistics each rank logs
are the all-reduced, job-wide ones.  Rank 0 prints one JSON line."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quadsim_b200 as qs
from quadsim_b200 import distributed as D, ppo

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 16, help="envs per rank (rank r gets n + r: ragged shards)")
ap.add_argument("--iters", type=int, default=8)
args = ap.parse_args()
rank, local_rank, world = D.init_from_env(backend="nccl")
dev = torch.device("cuda", local_rank)
n = args.n + rank
env = qs.make_vec("docking-v2", n, seed=0, randomize="episode", want_info=False, reuse_buffers=True, device=dev)
counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
torch.distributed.all_gather(counts, torch.tensor([n], dtype=torch.int64, device=dev))
id0s = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
torch.distributed.all_gather(id0s, torch.tensor([env.env_id0], dtype=torch.int64, device=dev))
counts, id0s = [int(c) for c in counts], [int(i) for i in id0s]
assert id0s == [sum(counts[:r]) for r in range(world)], (counts, id0s)
obs0 = env.reset()[:1024].clone()
peers = [torch.zeros_like(obs0) for _ in range(world)]
torch.distributed.all_gather(peers, obs0)
assert all(not torch.equal(peers[0], peers[r]) for r in range(1, world)), "ranks drew the same initial states"

algo = ppo.PPO2(env, n_steps=16, nminibatches=4, noptepochs=2, fused_policy=True, fused_update=True, seed=0)
w0 = torch.cat([p.detach().reshape(-1) for p in algo.policy.parameters()]).clone()
hist = algo.learn(args.iters * n * 16)
w1 = torch.cat([p.detach().reshape(-1) for p in algo.policy.parameters()])
allw = [torch.zeros_like(w1) for _ in range(world)]
torch.distributed.all_gather(allw, w1)
assert all(torch.equal(allw[0], a) for a in allw), "ranks' weights diverged"
assert torch.isfinite(w1).all() and not torch.equal(w0, w1)
row = {k: hist[-1][k] for k in ("timesteps", "mean_step_reward", "ep_mean_return", "ep_mean_length", "ep_episodes") if k in hist[-1]}
rows = [None] * world
torch.distributed.all_gather_object(rows, row)
same = lambda a, b: a == b or (a != a and b != b)  # NaN (no episode finished yet) on every rank alike
assert all(same(r[k], rows[0][k]) for r in rows for k in ("ep_mean_return", "ep_mean_length", "ep_episodes")), rows  # job-wide statistics
if rank == 0:
    print(json.dumps({"dist_ppo_check": "ok", "world": world, "envs": counts, "env_id0": id0s, "iters": len(hist),
                      "weights_identical": True, "weight_delta_l2": float((w1 - w0).norm()), "last": rows[0]}), flush=True)
torch.distributed.barrier()
torch.distributed.destroy_process_group()
