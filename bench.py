#!/usr/bin/env python
"""Headline benchmark: env-steps/s of the fused docking-v2 step (BASELINE.json `metric`), 2^20 envs per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the CPU port of the reference (oracle/), rank 0 only
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N   # one rank per GPU; envs shard, no collective per step

One "step" = one VecEnv.step of every env on every GPU (chaser command, target controller, 2x Drone.step, dock ports,
state2rel, reward/done, Philox auto-reset).  Prints ONE JSON line (rank 0).
  value     whole-job env-steps/s, inputs (pre-generated action tensors) resident in HBM, CUDA-event timed, max over ranks
  e2e       same metric through the public host-buffer API (numpy in/out): every step copies the actions host->device
            from pinned memory and obs/reward/done device->host inside the timed region
  roofline  HBM: algorithmic bytes per env-step (405 B fp32) x envs / kernel time vs MEASURED_PEAKS.json
  cpu_baseline  the fp64 C port of the reference (oracle/, "port") on this box's host cores, bounded sample
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES = {"f32": 405, "f64": 793}  # SURVEY.md section 8d / DESIGN.md "Algorithmic bytes"
LAYOUT_BYTES = {"f32": 389, "f64": 777}  # what the 40-word layout really moves per env-step (state R+W, actions, obs, reward, done)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--env", default="docking-v2")
    ap.add_argument("--envs-per-gpu", type=int, default=1 << 20)
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--e2e-steps", type=int, default=40)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-callers", action="store_true", help="skip the short PPO2-rollout measurement reported under 'callers'")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ CPU leg (oracle port)
def cpu_port_run(env_id, n_envs, steps, warmup, seconds=None, threads=None):
    """Times the fp64 C port of the reference (oracle/quadsim_oracle.c) with OpenMP over envs.  Returns a dict."""
    import numpy as np
    from oracle import oracle as O
    threads = threads or os.cpu_count() or 1
    env = O.OracleVecEnv(env_id, n_envs, seed=0, randomize=O.RAND_EPISODE, nthreads=threads)
    env.reset()
    rng = np.random.RandomState(0)
    acts = [rng.uniform(-1, 1, (n_envs, 4)).astype(np.float32).astype(np.float64) for _ in range(4)]
    for k in range(warmup):
        env.step(acts[k % 4])
    t0 = time.perf_counter()
    done_steps = 0
    while True:
        env.step(acts[done_steps % 4])
        done_steps += 1
        el = time.perf_counter() - t0
        if (steps is not None and done_steps >= steps) or (seconds is not None and el >= seconds):
            break
    el = time.perf_counter() - t0
    env.close()
    return {"value": n_envs * done_steps / el, "seconds": el, "steps": done_steps, "envs": n_envs, "cores": threads}


def run_reference(args):
    """--impl reference: the reference's algorithm on the host CPU.  The reference itself is pure Python that cannot
    travel to the GPU box (no /root/reference there), so this is its C port (oracle/), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # torchrun exports OMP_NUM_THREADS=1 to every rank; this arm is meant to use all host cores (libgomp reads the
    # variable when it is first loaded, which happens below)
    os.environ["OMP_NUM_THREADS"] = str(threads)
    os.environ.setdefault("OMP_PROC_BIND", "false")
    n = 1 << 15
    # size the sample so K steps stay within ~60 s: calibrate on a few steps
    cal = cpu_port_run(args.env, n, 3, 1, threads=threads)
    per_step = cal["seconds"] / cal["steps"]
    budget = 60.0
    while n > 1024 and per_step * (args.steps + args.warmup) > budget:
        n //= 2
        per_step /= 2
    res = cpu_port_run(args.env, n, args.steps, args.warmup, threads=threads)
    sample = "%d envs x %d steps of %s, fp64 C port of the reference, OpenMP %d threads" % (n, res["steps"], args.env, threads)
    line = {
        "impl": "reference", "metric": "env_steps_per_sec", "value": res["value"], "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * res["seconds"] / res["steps"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, n),
        "cpu_baseline": {"value": res["value"], "unit": "env-steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": res["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "unmodified Python reference measured in the build container: ~340 env-steps/s/core (BASELINE.md section 2)",
    }
    emit(line)


def workload_config(args, envs_per_gpu):
    return {"workload": "%s fused step (PID/vel target controller + 2x Drone.step + state2rel + reward/done + Philox auto-reset)" % args.env,
            "env": args.env, "envs_per_gpu": envs_per_gpu, "actions": "synthetic U(-1,1)^4 fp32, 8 pre-generated tensors cycled",
            "randomize_init": "episode (Philox4x32-10, +-0.5 m)", "l2": "working set (state 168 MB + 8 action tensors 134 MB + outputs 56 MB) > 126 MB L2, no flush needed",
            "parallelism": "env-sharded, dp%d, no collective on the step path" % args.gpus}


# ------------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz, self.ok = index, [], set(), False, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
             0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def run(self):
        if not self.ok:
            return
        while not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.NAMES.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def result(self):
        self.stop_flag = True
        self.join(1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def ncu_traffic(dtype):
    """dram bytes per launch of the step kernel from the committed ncu --set full capture (profiles/), or None."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        return d.get(dtype)
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import quadsim_b200 as qs
    from quadsim_b200 import distributed as D

    rank, local_rank, world = D.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; quadsim_b200 has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    tdt = torch.float32 if args.dtype == "f32" else torch.float64
    n = args.envs_per_gpu
    env = qs.make_vec(args.env, n, dtype=tdt, device=dev, seed=0, env_id0=rank * n, randomize="episode", want_info=False,
                      reuse_buffers=True)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    lo = 0.0 if args.env.startswith("hover") else -1.0
    acts = [(torch.rand((n, 4), generator=gen, device=dev, dtype=torch.float32) * (1 - lo) + lo).to(tdt) for _ in range(8)]
    env.reset()
    for k in range(max(args.warmup, 3)):
        env.step(acts[k % 8])
    torch.cuda.synchronize(dev)
    env.stats()  # clear warm-up statistics

    sampler = ClockSampler(local_rank)
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize(dev)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for k in range(args.steps):
        env.step(acts[k % 8])
    ev1.record()
    torch.cuda.synchronize(dev)
    if world > 1:
        torch.distributed.barrier()
    clocks = sampler.result()
    ms_local = ev0.elapsed_time(ev1)
    ms = D.max_over_ranks(ms_local, dev)
    stats = env.stats(all_reduce=True)  # the one collective of a rollout: 5 doubles
    value = world * n * args.steps / (ms * 1e-3)

    # ---- e2e: host buffers through the public API (qs_step_host), H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        henv = qs.make_vec(args.env, n, dtype=tdt, device=dev, seed=0, env_id0=rank * n, randomize="episode", want_info=False)
        h_acts, h_obs, h_rew, h_done = henv.pinned_host_buffers(n_action_buffers=4)
        for k, a in enumerate(h_acts):
            a[...] = acts[k].cpu().numpy()
        henv.reset()
        for k in range(3):
            henv.step_host(h_acts[k % 4], out=(h_obs, h_rew, h_done))
        if world > 1:
            torch.distributed.barrier()
        t0 = time.perf_counter()
        for k in range(args.e2e_steps):
            henv.step_host(h_acts[k % 4], out=(h_obs, h_rew, h_done))  # blocking: results are in host memory on return
        el = time.perf_counter() - t0
        el = D.max_over_ranks(el, dev)
        esz = 4 if args.dtype == "f32" else 8
        chunks, left, cur = 0, n, 1 << 15  # qs_step_host's graduated schedule: 32 Ki, 64 Ki, 128 Ki, then 256 Ki envs per chunk
        while left > 0:
            chunks, left, cur = chunks + 1, left - min(left, cur), min(cur * 2, 1 << 18)
        # what bounds this leg: the device->host link.  Plain pinned D2H / H2D copies of the same sizes, same process.
        d_buf = torch.empty(n * (henv.obs_dim * esz + esz + 1), dtype=torch.uint8, device=dev)
        h_buf = torch.empty(d_buf.numel(), dtype=torch.uint8, pin_memory=True)
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h_buf.copy_(d_buf, non_blocking=True)
        pe0.record()
        for _ in range(10):
            h_buf.copy_(d_buf, non_blocking=True)
        pe1.record()
        torch.cuda.synchronize(dev)
        d2h_gbs = 10 * d_buf.numel() / (pe0.elapsed_time(pe1) * 1e-3) / 1e9
        numa = D.gpu_numa_node(local_rank)
        e2e = {"value": world * n * args.e2e_steps / el, "unit": "env-steps/s", "h2d_bytes_per_step": n * 4 * esz,
               "pcie": {"d2h_copy_GBps": d2h_gbs, "d2h_achieved_GBps": n * args.e2e_steps / el * (henv.obs_dim * esz + esz + 1) / 1e9,
                        "note": "this rank; the leg is bound by the device->host link (53 B out vs 16 B in per env-step)"},
               "numa": {"gpu_node": numa, "cpus_allowed": len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else None},
               "d2h_bytes_per_step": n * (henv.obs_dim * esz + esz + 1), "steps": args.e2e_steps,
               "api": "QuadVecEnv.step_host -> qs_step_host (pinned numpy in/out, %d graduated chunks pipelined over 4 streams; obs by DMA, "
                      "reward/done written by the kernel into the pinned buffers)" % chunks,
               "checksum_reward": float(h_rew.sum())}
        henv.close()

    # ---- the path's caller (SURVEY 8f-1 / BASELINE config 5), a few seconds: PPO2 rollout with the tcgen05 policy in the loop
    callers = None
    if not args.no_callers and args.dtype == "f32" and not args.env.startswith("hover"):
        callers = measure_callers(args, rank, dev, world, D)

    if rank != 0:
        return
    peak, peak_src = measured_peak()
    algo = ALGO_BYTES[args.dtype]
    kernel_ms = ms_local / args.steps
    achieved = algo * n / (kernel_ms * 1e-3) / 1e9
    traffic = ncu_traffic(args.dtype)
    pipe = os.environ.get("QUADSIM_B200_KERNEL", "pipe" if args.dtype == "f64" else "direct") == "pipe" and not args.env.startswith("hover")
    kernel = "%s<%s,%s>" % ("dock_step_pipe_kernel" if pipe else "dock_step_kernel", "float" if args.dtype == "f32" else "double", args.env)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": kernel, "algorithmic_bytes_per_env_step": algo, "layout_bytes_per_env_step": LAYOUT_BYTES[args.dtype],
                "layout_GBps": LAYOUT_BYTES[args.dtype] * n / (kernel_ms * 1e-3) / 1e9, "kernel_ms": kernel_ms, "peak_source": peak_src,
                "timing": "CUDA events on the launch stream around the %d timed launches / %d" % (args.steps, args.steps)}
    try:  # informational: the no-arithmetic kernel with the step's exact loads/stores (tools/microbench/step_pattern_bw.cu)
        pc = json.load(open(os.path.join(ROOT, "profiles", "r1_step_pattern_ceiling.json")))["results"]
        hit = [r for r in pc if r["pdl"] and r["envs"] == n and r["pattern"] == "docking " + args.dtype.replace("f", "fp")]
        if hit and args.env.startswith("docking"):
            roofline["pattern_ceiling"] = {"us_per_launch": hit[0]["us_per_launch"], "frac": hit[0]["us_per_launch"] / (kernel_ms * 1e3),
                                           "source": "profiles/r1_step_pattern_ceiling.json (measured once, not in this run)"}
    except Exception:
        pass
    cpu = None
    if not args.no_cpu_baseline:
        D.restore_affinity()  # the CPU leg uses every host core, not just the GPU's NUMA node
        threads = os.cpu_count() or 1
        res = cpu_port_run(args.env, 1 << 15, None, 2, seconds=args.cpu_seconds, threads=threads)
        cpu = {"value": res["value"], "unit": "env-steps/s", "cores": threads, "kind": "port",
               "sample": "%d envs x %d steps of %s (%.1f s), fp64 C port of the reference with OpenMP; the Python reference "
                         "itself measured ~340 env-steps/s/core in the build container (BASELINE.md)" % (res["envs"], res["steps"], args.env, res["seconds"])}
    line = {
        "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": workload_config(args, n),
        "clocks": clocks, "e2e": e2e, "gpu_launches": args.steps, "roofline": roofline, "cpu_baseline": cpu,
        "episode_stats": {k: stats[k] for k in ("episodes", "mean_length", "mean_return", "docked_steps", "nonfinite_steps")},
        "callers": callers,
    }
    emit(line)


def measure_callers(args, rank, dev, world, D, n_steps=16, rollouts=5):
    """Not the headline: the rollout-in-the-loop numbers of the path's caller, measured in the same run.  qs_rollout (policy
    forward on tcgen05 + env step, in place, 2 launches per step) + qs_gae, then one optimisation epoch of the fused PPO2
    update (qs_policy_ppo_grad + library GEMMs + Adam).  CUDA-event timed, max over ranks."""
    import torch
    import quadsim_b200 as qs
    from quadsim_b200 import ppo
    n = args.envs_per_gpu
    env = qs.make_vec(args.env, n, device=dev, seed=0, env_id0=rank * n, randomize="episode", want_info=False, reuse_buffers=True)
    algo = ppo.PPO2(env, n_steps=n_steps, nminibatches=10, noptepochs=1, fused_policy=True, seed=0)
    for _ in range(2):
        batch = algo.runner.run()
        algo.update(batch)  # twice: workspaces, GEMM heuristics and the NCCL communicator are all set up lazily
    torch.cuda.synchronize(dev)
    if world > 1:
        torch.distributed.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    for _ in range(rollouts):
        batch = algo.runner.run()
    ev[1].record()
    algo.update(batch)
    ev[2].record()
    torch.cuda.synchronize(dev)
    roll_ms = D.max_over_ranks(ev[0].elapsed_time(ev[1]) / rollouts, dev)
    upd_ms = D.max_over_ranks(ev[1].elapsed_time(ev[2]), dev)
    env.close()
    return {"what": "PPO2 %s rollout-in-the-loop, %d envs/GPU, %d-step rollouts (BASELINE config 5 shape): tcgen05 policy forward + "
                    "fused env step written in place (qs_rollout) + qs_gae; then one epoch of the fused update" % (args.env, n, n_steps),
            "rollout_env_steps_per_s": world * n * n_steps / (roll_ms * 1e-3), "rollout_ms": roll_ms, "gpu_launches_per_rollout": algo.runner.n_slices * (2 * n_steps + 1) + 1, "env_slices": algo.runner.n_slices,
            "update_samples_per_s": world * n * n_steps / (upd_ms * 1e-3), "update_epoch_ms": upd_ms}


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the process's original stdout; everything else (NCCL's version banner, library chatter
    written straight to fd 1) was redirected to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
