#!/usr/bin/env python
"""Kernel-time table of one PPO2 optimisation epoch (torch.profiler), fused (qs_policy_ppo_grad) or autograd.

    python tools/update_profile.py [--n 1048576] [--n-steps 16] [--autograd]
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
ap.add_argument("--autograd", action="store_true")
args = ap.parse_args()
env = qs.make_vec("docking-v2", args.n, seed=0, randomize="episode", want_info=False, reuse_buffers=True)
algo = ppo.PPO2(env, n_steps=args.n_steps, nminibatches=10, noptepochs=1, fused_policy=True, fused_update=not args.autograd)
batch = algo.runner.run()
algo.update(batch)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
algo.update(batch)
e1.record()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    algo.update(batch)
    torch.cuda.synchronize()
rows = [(k.key, k.device_time_total, k.count) for k in prof.key_averages() if k.device_time_total > 0 and k.device_type.name == "CUDA"]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print("%-100s %10s %6s %8s" % ("kernel", "us total", "calls", "share"))
for name, us, cnt in rows[:22]:
    print("%-100s %10.1f %6d %7.1f%%" % (name[:100], us, cnt, 100.0 * us / tot))
print(json.dumps({"envs": args.n, "n_steps": args.n_steps, "samples": args.n * args.n_steps, "fused_update": not args.autograd,
                  "epoch_ms": e0.elapsed_time(e1), "kernel_time_sum_ms": tot / 1e3,
                  "samples_per_s": args.n * args.n_steps / (e0.elapsed_time(e1) * 1e-3)}))
