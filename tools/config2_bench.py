#!/usr/bin/env python
"""BASELINE config 2: hovering-v0 with random actions, 4096 batched envs on one B200 (the dynamics-only kernel).  At this
batch size a step is latency-bound, not bandwidth-bound: reports microseconds per step for one launch per step (the VecEnv
protocol) and for qs_step_n (T steps per launch, state in registers across steps).  One JSON line."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quadsim_b200 as qs

ap = argparse.ArgumentParser()
ap.add_argument("--env", default="hovering-v0")
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--steps", type=int, default=1000)
ap.add_argument("--dtype", default="f32")
args = ap.parse_args()
dt = torch.float32 if args.dtype == "f32" else torch.float64
env = qs.make_vec(args.env, args.n, dtype=dt, seed=0, want_info=False, reuse_buffers=True)
g = torch.Generator(device="cuda").manual_seed(0)
acts = torch.rand((args.steps, args.n, 4), device="cuda", generator=g, dtype=torch.float32).to(dt)  # U[0,1)^4, seed 0
if args.env != "hovering-v0":
    acts = acts * 2 - 1
env.reset()
for t in range(20):
    env.step(acts[t])
env.step_many(acts[:20])
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
e[0].record()
for t in range(args.steps):
    env.step(acts[t])
e[1].record()
torch.cuda.synchronize()
best = 1e30
for _ in range(5):
    e[2].record()
    env.step_many(acts)
    e[3].record()
    torch.cuda.synchronize()
    best = min(best, e[2].elapsed_time(e[3]))
per_launch, many = e[0].elapsed_time(e[1]) / args.steps * 1e3, best / args.steps * 1e3
print(json.dumps({"config": "%s, %d envs, %d steps, random actions (BASELINE config 2)" % (args.env, args.n, args.steps), "dtype": args.dtype,
                  "us_per_step_one_launch_per_step": per_launch, "env_steps_per_s_one_launch_per_step": args.n / per_launch * 1e6,
                  "us_per_step_step_n": many, "env_steps_per_s_step_n": args.n / many * 1e6, "stats": env.stats()}))
