#!/usr/bin/env python
"""Where does a PPO2 rollout step go?  Kernel-time table (torch.profiler / CUPTI) of DeviceRunner.run() plus the host-side
wall time of the same loop, so launch-bound and GPU-bound rollouts can be told apart.  Prints a table and one JSON line.

    python tools/rollout_profile.py [--n 1048576] [--n-steps 16] [--fused]
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
ap.add_argument("--n", type=int, default=1 << 20)
ap.add_argument("--n-steps", type=int, default=16)
ap.add_argument("--fused", action="store_true")
ap.add_argument("--tf32", action="store_true")
ap.add_argument("--slices", type=int, default=2)
args = ap.parse_args()
torch.backends.cuda.matmul.allow_tf32 = args.tf32
env = qs.make_vec("docking-v2", args.n, seed=0, randomize="episode", want_info=False, reuse_buffers=True)
algo = ppo.PPO2(env, n_steps=args.n_steps, nminibatches=10, noptepochs=1, fused_policy=args.fused)
algo.runner.n_slices = args.slices
for _ in range(2):
    algo.runner.run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
algo.runner.run()
e1.record()
t_host = time.perf_counter() - t0  # host time to ENQUEUE the rollout
torch.cuda.synchronize()
t_wall = time.perf_counter() - t0
t_gpu = e0.elapsed_time(e1)
from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    algo.runner.run()
    torch.cuda.synchronize()
rows = [(k.key, k.device_time_total, k.count) for k in prof.key_averages() if k.device_time_total > 0 and k.device_type.name == "CUDA"]
if not rows:
    rows = [(k.key, k.device_time_total, k.count) for k in prof.key_averages() if k.device_time_total > 0]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print("%-90s %10s %6s %8s" % ("kernel", "us total", "calls", "share"))
for name, us, cnt in rows[:25]:
    print("%-90s %10.1f %6d %7.1f%%" % (name[:90], us, cnt, 100.0 * us / tot))
print(json.dumps({"envs": args.n, "n_steps": args.n_steps, "fused": args.fused, "slices": args.slices, "tf32": args.tf32, "rollout_gpu_ms": t_gpu,
                  "rollout_host_enqueue_ms": t_host * 1e3, "rollout_wall_ms": t_wall * 1e3, "kernel_time_sum_ms": tot / 1e3,
                  "ms_per_step": t_gpu / args.n_steps}))
