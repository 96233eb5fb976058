#!/usr/bin/env python
"""End-to-end check that the device-resident PPO2 (tcgen05 policy rollout + fused update) LEARNS docking-v2: runs
`PPO2.learn` with the reference's hyper-parameters (run_moving_docking_ppo2.py:88-104) except n_steps (16 instead of 600,
SURVEY H6) and prints one JSON line per iteration (episode statistics come from the env's device-side counters).

    python tools/train_demo.py [--n 262144] [--iters 60] [--epochs 4] [--save model.zip]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quadsim_b200 as qs
from quadsim_b200 import ppo

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 18)
ap.add_argument("--n-steps", type=int, default=16)
ap.add_argument("--iters", type=int, default=60)
ap.add_argument("--epochs", type=int, default=4)
ap.add_argument("--autograd", action="store_true", help="torch autograd update instead of qs_policy_ppo_grad (A/B of learning curves)")
ap.add_argument("--graph", action="store_true", help="replay the minibatch step from a CUDA graph (PPO2(graph_update=True))")
ap.add_argument("--save", default=None)
args = ap.parse_args()
env = qs.make_vec("docking-v2", args.n, seed=0, randomize="episode", want_info=False, reuse_buffers=True)
algo = ppo.PPO2(env, n_steps=args.n_steps, nminibatches=10, noptepochs=args.epochs, fused_policy=True, fused_update=not args.autograd, graph_update=args.graph, seed=0)
t0 = time.perf_counter()


def log_fn(log):
    torch.cuda.synchronize()
    row = {k: (round(v, 5) if isinstance(v, float) else v) for k, v in log.items()}
    row["wall_s"] = round(time.perf_counter() - t0, 2)
    print(json.dumps(row), flush=True)


hist = algo.learn(args.iters * args.n * args.n_steps, log_fn=log_fn)
first, last = hist[: max(1, len(hist) // 10)], hist[-max(1, len(hist) // 10):]
m = lambda rows, k: sum(r[k] for r in rows) / len(rows)
print(json.dumps({"summary": True, "envs": args.n, "iters": len(hist), "env_steps": hist[-1]["timesteps"], "wall_s": round(time.perf_counter() - t0, 2),
                  "fused_update": not args.autograd, "graph_update": args.graph, "mean_step_reward_first": m(first, "mean_step_reward"), "mean_step_reward_last": m(last, "mean_step_reward"),
                  "ep_mean_return_first": m(first, "ep_mean_return"), "ep_mean_return_last": m(last, "ep_mean_return"),
                  "ep_mean_length_first": m(first, "ep_mean_length"), "ep_mean_length_last": m(last, "ep_mean_length")}))
if args.save:
    algo.save(args.save)
