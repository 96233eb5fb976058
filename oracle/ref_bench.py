"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- times the UNMODIFIED reference (staged under oracle/_ref/ by make_ref.py) on
this machine's host cores.  Used by bench.py's cpu_baseline leg; never imported by the product package.

    python oracle/ref_bench.py --env docking-v2 --seconds 10 --seed 3      # ONE worker: prints one JSON line
    python oracle/ref_bench.py --config1                                   # run_sim_PID.py:43-54, 2000 steps, one process

What one worker does is what one of the reference's own `SubprocVecEnv` workers does
(run_moving_docking_ppo2.py:65-67, and stable-baselines' worker loop): construct the env with `gym.make`'s entry point,
then `obs, r, done, info = env.step(a); if done: obs = env.reset()` for a fixed wall-clock window, with the benchmark's
action distribution U(-1,1)^4 (U(0,1)^4 for hovering-v0).  bench.py starts one such process per host core, all at once
(`run_workers`), and adds the step counts up -- BASELINE.md section 4.
"""
import argparse
import json
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _env_class(ns, env_id):
    return {"docking-v0": ns.DockingEnv, "docking-v2": ns.MovingDockingEnv, "hovering-v0": ns.HoveringEnv}[env_id]


def worker(env_id, seconds, seed):
    import numpy as np
    sys.path.insert(0, ROOT)
    from oracle import ref_shim
    ref_shim.use_staged()
    ns = ref_shim.load()
    env = _env_class(ns, env_id)()
    rng = np.random.RandomState(seed)
    lo = 0.0 if env_id.startswith("hover") else -1.0
    acts = rng.uniform(lo, 1.0, (4096, 4))
    env.reset()
    for k in range(20):  # warm-up: imports, first-call allocations
        _, _, done, _ = env.step(acts[k])
        if done:
            env.reset()
    steps = episodes = 0
    t0 = time.perf_counter()
    t_end = t0 + seconds
    while True:
        _, _, done, _ = env.step(acts[steps & 4095])
        steps += 1
        if done:
            env.reset()
            episodes += 1
        if (steps & 15) == 0 and time.perf_counter() >= t_end:
            break
    el = time.perf_counter() - t0
    return {"env": env_id, "steps": steps, "episodes": episodes, "seconds": el, "steps_per_s": steps / el}


def config1():
    """BASELINE config 1: the loop of run_sim_PID.py:43-54 (start at the origin, pos_des = [-0.2, 0.2, 0.2], 2000 steps of
    state_now = quad.get_state(); u = control.PID(state_des, state_now); quad.step(u)) on the unmodified reference."""
    import numpy as np
    sys.path.insert(0, ROOT)
    from oracle import ref_shim
    ref_shim.use_staged()
    ns = ref_shim.load()
    quad = ns.Drone()
    control = ns.controller(quad.arm_length, quad.mass)
    ini = np.zeros(13)
    ini[6] = 1.0
    quad.reset(ini)
    des = np.zeros(13)
    des[0:3] = [-0.2, 0.2, 0.2]
    des[6] = 1.0
    n = 2000
    t0 = time.perf_counter()
    for _ in range(n):
        now = quad.get_state()
        u = control.PID(des, now)
        quad.step(u)
    el = time.perf_counter() - t0
    return {"steps": n, "seconds": el, "steps_per_s": n / el, "final_pos": [float(x) for x in quad.get_state()[:3]]}


def run_workers(env_id, seconds, procs):
    """Start `procs` single-env worker processes at once (fresh interpreters: nothing of the caller's CUDA / torch state is
    inherited), wait for all, and return the aggregate."""
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
    cmd = [sys.executable, os.path.abspath(__file__), "--env", env_id, "--seconds", str(seconds)]
    ps = [subprocess.Popen(cmd + ["--seed", str(k)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env) for k in range(procs)]
    rows, errs = [], []
    for p in ps:
        out, err = p.communicate()
        if p.returncode == 0:
            rows.append(json.loads(out.decode().strip().splitlines()[-1]))
        else:
            errs.append(err.decode()[-300:])
    if not rows:
        raise RuntimeError("no reference worker finished: " + (errs[0] if errs else "?"))
    total = sum(r["steps_per_s"] for r in rows)
    return {"value": total, "per_core": total / len(rows), "procs": len(rows), "failed": len(errs), "seconds": max(r["seconds"] for r in rows),
            "steps": sum(r["steps"] for r in rows), "episodes": sum(r["episodes"] for r in rows)}


def run_config1():
    out = subprocess.check_output([sys.executable, os.path.abspath(__file__), "--config1"],
                                  env=dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1"))
    return json.loads(out.decode().strip().splitlines()[-1])


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", default="docking-v2")
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--config1", action="store_true")
    a = ap.parse_args()
    print(json.dumps(config1() if a.config1 else worker(a.env, a.seconds, a.seed)))
