#!/usr/bin/env python
"""Do the policy forward kernel and the env step kernel overlap when they run on two streams?  Times, on one GPU:
P alone (policy forward on half of the envs), E alone (env step of the other half, qs_step_range), and P || E issued on two
streams in the same iteration.  `both` ~ max(P, E): they overlap; `both` ~ P + E: they serialise / contend.

    python tools/overlap_probe.py [--n 1048576]
"""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quadsim_b200 as qs
from quadsim_b200 import _capi, ppo

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 20)
ap.add_argument("--iters", type=int, default=200)
args = ap.parse_args()
n, half = args.n, args.n // 2
env = qs.make_vec("docking-v2", n, seed=0, randomize="episode", want_info=False, reuse_buffers=True)
algo = ppo.PPO2(env, n_steps=4, nminibatches=2, noptepochs=1, fused_policy=True, seed=0)
fused, lib = algo.fused, _capi.load()
obs = env.reset().clone()
act = torch.rand((n, 4), device="cuda") * 2 - 1
obs_out, rew, done = torch.empty_like(obs), torch.empty(n, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()
p = lambda t: C.c_void_p(t.data_ptr())


def P():
    with torch.cuda.stream(s0):
        fused.step(obs[:half])


def E():
    _capi.check(lib.qs_step_range(env._h, half, n - half, p(act), p(obs_out), p(rew), p(done), C.c_void_p(s1.cuda_stream)))


def timed(fns):
    for _ in range(20):
        for f in fns:
            f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s0.wait_event(e0); s1.wait_event(e0)
    for _ in range(args.iters):
        for f in fns:
            f()
    d0, d1 = torch.cuda.Event(), torch.cuda.Event()
    d0.record(s0); d1.record(s1)
    torch.cuda.current_stream().wait_event(d0); torch.cuda.current_stream().wait_event(d1)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.iters * 1e3


res = {"envs": n, "P_us (policy fwd, half)": timed([P]), "E_us (env step, other half)": timed([E]), "both_us (two streams)": timed([P, E])}
res["overlap"] = (res["P_us (policy fwd, half)"] + res["E_us (env step, other half)"] - res["both_us (two streams)"]) / min(res["P_us (policy fwd, half)"], res["E_us (env step, other half)"])
print(json.dumps(res))
