#!/usr/bin/env python
"""The reference's training script run_moving_docking_ppo2.py:60-110 on the device-resident stack: the same PPO2 call
(net_arch [128, dict(pi=[128], vf=[128])], relu, lam .95, gamma .99, ent_coef 0, lr 3e-4, vf_coef .5, max_grad_norm .5,
nminibatches 10, noptepochs 10, cliprange .2) with `SubprocVecEnv x 10` replaced by N envs in one CUDA handle and the 600-step
rollouts by short ones.  Checkpoints like CheckpointCallback(save_freq, save_path='./logs/').

    python examples/run_moving_docking_ppo2.py [--n 262144] [--n-steps 16] [--total-timesteps 2e9]
    torchrun --nproc-per-node 8 examples/run_moving_docking_ppo2.py        # one rank per GPU, gradients all-reduced"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quadsim_b200
from quadsim_b200 import distributed as D
from quadsim_b200 import ppo

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 18)
ap.add_argument("--n-steps", type=int, default=16)
ap.add_argument("--total-timesteps", type=float, default=2e9)
ap.add_argument("--save-path", default="./logs/")
args = ap.parse_args()
rank, local_rank, world = D.init_from_env()
env = quadsim_b200.make_vec("gym_docking:docking-v2", args.n, device=torch.device("cuda", local_rank), seed=0, randomize="episode",
                            want_info=False, reuse_buffers=True)     # env ids are global: rank r owns [r n, (r+1) n)
model = ppo.PPO2(env, n_steps=args.n_steps, lam=0.95, gamma=0.99, ent_coef=0.0, learning_rate=3e-4, vf_coef=0.5, max_grad_norm=0.5,
                 nminibatches=10, noptepochs=10, cliprange=0.2, fused_policy=True)
log = (lambda row: print(json.dumps({k: (round(v, 5) if isinstance(v, float) else v) for k, v in row.items()}), flush=True)) if rank == 0 else None
model.learn(total_timesteps=int(args.total_timesteps) // world, log_fn=log, save_freq=int(5e7), save_path=args.save_path)
if rank == 0:
    os.makedirs(args.save_path, exist_ok=True)
    model.save(os.path.join(args.save_path, "ppo2_docking_v2_final.zip"))
