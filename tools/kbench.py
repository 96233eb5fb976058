#!/usr/bin/env python
"""Kernel micro-benchmark for build variants: times the fused step kernel in steady state (episodes de-synchronised,
~1.35 % of envs resetting per step) for each library given.  Used to pick compile flags / kernel designs; the numbers
judged are bench.py's.

    python tools/kbench.py [--env docking-v2] [--n 1048576] [--dtype f32] lib1.so [lib2.so ...]
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(args):
    sys.path.insert(0, ROOT)
    import torch
    import quadsim_b200 as qs
    tdt = torch.float32 if args.dtype == "f32" else torch.float64
    env = qs.make_vec(args.env, args.n, dtype=tdt, seed=0, randomize="episode", want_info=False, reuse_buffers=True)
    gen = torch.Generator(device="cuda").manual_seed(1)
    lo = 0.0 if args.env.startswith("hover") else -1.0
    acts = [(torch.rand((args.n, 4), generator=gen, device="cuda") * (1 - lo) + lo).to(tdt) for _ in range(8)]
    env.reset()
    for k in range(args.warmup):
        env.step(acts[k % 8])
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(args.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(args.steps):
            env.step(acts[k % 8])
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / args.steps)
    algo = {"f32": {"hov": 225, "doc": 405}, "f64": {"hov": 425, "doc": 793}}[args.dtype][args.env[:3]]
    st = env.stats()
    knobs = {k: v for k, v in os.environ.items() if k.startswith("QUADSIM_B200_")}
    print(json.dumps({"lib": knobs or "default", "env": args.env, "dtype": args.dtype, "n": args.n,
                      "us_per_step": best * 1e3, "env_steps_per_s": args.n / (best * 1e-3),
                      "algo_GBps": algo * args.n / (best * 1e-3) / 1e9, "episodes": st["episodes"], "nonfinite": st["nonfinite_steps"]}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="*")
    ap.add_argument("--env", default="docking-v2")
    ap.add_argument("--n", type=int, default=1 << 20)
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--warmup", type=int, default=300)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--child", action="store_true")
    args = ap.parse_args()
    if args.child:
        return child(args)
    for lib in args.libs or [""]:
        env = dict(os.environ)
        if "=" in lib:  # "KEY=VAL,KEY2=VAL2": environment knobs on the default library
            for kv in lib.split(","):
                k, v = kv.split("=", 1)
                env[k] = v
        elif lib:
            env["QUADSIM_B200_LIB"] = os.path.abspath(lib)
        cmd = [sys.executable, os.path.abspath(__file__), "--child", "--env", args.env, "--n", str(args.n), "--dtype", args.dtype,
               "--warmup", str(args.warmup), "--steps", str(args.steps), "--reps", str(args.reps)]
        r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "no output", flush=True)


if __name__ == "__main__":
    main()
