#!/usr/bin/env python
"""BASELINE config 5: PPO2 docking-v2 rollout-in-the-loop with device-resident obs/reward, 8 Mi envs per GPU, one rank per
GPU (torchrun), gradients all-reduced over NCCL.  Per PPO iteration: a T-step rollout (qs_rollout: tcgen05 policy + fused env
step, in place) + GAE + `--epochs` optimisation epochs of 10 minibatches (qs_policy_ppo_grad incl. the weight-gradient contraction + qs_adam_step).
Prints one JSON line on rank 0; times are CUDA-event, max over ranks.

    python tools/config5_bench.py [--envs-per-gpu 8388608] [--n-steps 16] [--iters 3] [--epochs 1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/config5_bench.py
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quadsim_b200 as qs
from quadsim_b200 import distributed as D
from quadsim_b200 import ppo

ap = argparse.ArgumentParser()
ap.add_argument("--envs-per-gpu", type=int, default=1 << 23)
ap.add_argument("--n-steps", type=int, default=16)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--epochs", type=int, default=1)
args = ap.parse_args()
rank, local_rank, world = D.init_from_env()
dev = torch.device("cuda", local_rank)
n, T = args.envs_per_gpu, args.n_steps
env = qs.make_vec("docking-v2", n, device=dev, seed=0, env_id0=rank * n, randomize="episode", want_info=False, reuse_buffers=True)
algo = ppo.PPO2(env, n_steps=T, nminibatches=10, noptepochs=args.epochs, fused_policy=True, seed=0)
for _ in range(2):
    batch = algo.runner.run()
    algo.update(batch)  # warm-up: workspaces, cuBLAS heuristics, NCCL communicators
env.stats()
torch.cuda.synchronize(dev)
if world > 1:
    torch.distributed.barrier()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.iters + 1)]
ev[0].record()
for it in range(args.iters):
    batch = algo.runner.run()
    ev[2 * it + 1].record()
    log = algo.update(batch)
    ev[2 * it + 2].record()
torch.cuda.synchronize(dev)
roll_ms = sum(ev[2 * it].elapsed_time(ev[2 * it + 1]) for it in range(args.iters)) / args.iters
upd_ms = sum(ev[2 * it + 1].elapsed_time(ev[2 * it + 2]) for it in range(args.iters)) / args.iters
roll_ms, upd_ms = D.max_over_ranks(roll_ms, dev), D.max_over_ranks(upd_ms, dev)
stats = env.stats(all_reduce=world > 1)
if rank == 0:
    steps = world * n * T
    print(json.dumps({
        "config": "PPO2 docking-v2 rollout-in-the-loop (BASELINE config 5)", "n_gpus": world, "envs_total": world * n, "envs_per_gpu": n,
        "n_steps": T, "epochs": args.epochs, "minibatches": 10, "iters_timed": args.iters,
        "rollout_ms": roll_ms, "rollout_env_steps_per_s": steps / (roll_ms * 1e-3),
        "update_ms": upd_ms, "update_samples_per_s": steps * args.epochs / (upd_ms * 1e-3),
        "iteration_ms": roll_ms + upd_ms, "env_steps_per_s_incl_update": steps / ((roll_ms + upd_ms) * 1e-3),
        "gpu_mem_GB": torch.cuda.max_memory_allocated(dev) / 1e9, "log": log,
        "episodes": stats["episodes"], "mean_return": stats["mean_return"], "mean_length": stats["mean_length"],
        "collectives": "one gradient all-reduce (35 k floats, flat bucket) per minibatch + 5 doubles of episode statistics per iteration"}))
if world > 1:
    torch.distributed.destroy_process_group()
