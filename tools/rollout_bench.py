#!/usr/bin/env python
"""BASELINE config 5 measurement: PPO2 rollout-in-the-loop on device-resident tensors (policy forward + env step + GAE) and
one optimisation pass, at N envs per GPU.  Prints one JSON line.

    python tools/rollout_bench.py [--n 1048576] [--n-steps 16] [--rollouts 5] [--amp]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quadsim_b200 as qs
from quadsim_b200 import ppo

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 20)
ap.add_argument("--n-steps", type=int, default=16)
ap.add_argument("--rollouts", type=int, default=5)
ap.add_argument("--tf32", action="store_true")
ap.add_argument("--fused", action="store_true", help="tcgen05 fused policy kernel (qs_policy_step)")
ap.add_argument("--slices", type=int, default=2, help="qs_rollout env slices (2 = overlap policy and env kernels)")
ap.add_argument("--autograd-update", action="store_true", help="with --fused: keep the torch autograd update (A/B)")
ap.add_argument("--epochs", type=int, default=1)
ap.add_argument("--bf16", action="store_true", help="bf16 autocast for the rollout forward (cuBLAS tensor cores)")
args = ap.parse_args()
torch.backends.cuda.matmul.allow_tf32 = args.tf32
env = qs.make_vec("docking-v2", args.n, seed=0, randomize="episode", want_info=False, reuse_buffers=True)
algo = ppo.PPO2(env, n_steps=args.n_steps, nminibatches=10, noptepochs=args.epochs, fused_policy=args.fused,
                fused_update=args.fused and not args.autograd_update)
algo.runner.n_slices = args.slices
for _ in range(3):  # warm-up; keep the batch alive like the timed loop does, so the caching allocator reaches steady state
    batch = algo.runner.run()
torch.cuda.synchronize()
e0, e1, e2, e3 = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
e0.record()
with torch.autocast("cuda", dtype=torch.bfloat16, enabled=args.bf16):
    for _ in range(args.rollouts):
        batch = algo.runner.run()
e1.record()
algo.update(batch)  # warm-up: workspace allocation, cuBLAS heuristics
torch.cuda.synchronize()
e2.record()
log = algo.update(batch)
e3.record()
torch.cuda.synchronize()
# policy forward alone (the kernel the rollout is bound by)
fwd = (lambda: algo.fused.step(algo.runner.obs)) if args.fused else (lambda: algo.policy.step(algo.runner.obs))
for _ in range(3):
    fwd()
f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
f0.record()
for _ in range(50):
    fwd()
f1.record()
torch.cuda.synchronize()
fwd_ms = f0.elapsed_time(f1) / 50
flops = 2.0 * args.n * (12 * 128 + 2 * 128 * 128 + 128 * 5)
steps = args.n * args.n_steps
t_roll = e0.elapsed_time(e1) / args.rollouts
t_upd = e2.elapsed_time(e3)
print(json.dumps({"config": "PPO2 docking-v2 rollout-in-the-loop", "envs": args.n, "n_steps": args.n_steps, "tf32": args.tf32, "bf16_rollout": args.bf16, "fused_tcgen05": args.fused, "in_place_rollout": bool(algo.runner.in_place), "fused_update": bool(algo.fused_update), "epochs": args.epochs, "slices": args.slices, "policy_fwd_ms": fwd_ms,
                  "policy_fwd_TFLOPs": flops / (fwd_ms * 1e-3) / 1e12,
                  "rollout_ms": t_roll, "rollout_env_steps_per_s": steps / (t_roll * 1e-3),
                  "ms_per_env_step_launch": t_roll / args.n_steps, "update_ms": t_upd, "update_ms_per_epoch": t_upd / args.epochs,
                  "update_samples_per_s": steps * args.epochs / (t_upd * 1e-3), "log": log,
                  "gpu_mem_GB": torch.cuda.max_memory_allocated() / 1e9}))
