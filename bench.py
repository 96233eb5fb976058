#!/usr/bin/env python
"""Headline benchmark: env-steps/s of the fused docking-v2 step (BASELINE.json `metric`), 2^20 envs per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference's algorithm on the host CPU, rank 0 only
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N   # one rank per GPU; envs shard, no collective per step

One "step" = one VecEnv.step of every env on every GPU (chaser command, target controller, 2x Drone.step, dock ports,
state2rel, reward/done, Philox auto-reset).  Prints ONE JSON line (rank 0).
  value     whole-job env-steps/s, inputs (pre-generated action tensors) resident in HBM, CUDA-event timed, max over ranks;
            the loop cycles 8 action tensors and 8 output buffer sets, so no launch rewrites what the previous one wrote
  e2e       same metric through the public host-buffer API (numpy in/out): every step copies the actions host->device
            from pinned memory and obs/reward/done device->host inside the timed region
  roofline  HBM: algorithmic bytes per env-step (405 B fp32) x envs / kernel time vs MEASURED_PEAKS.json;
            `dram_frac` is the same fraction in measured DRAM bytes (`traffic`, a committed ncu capture -- see traffic_source)
  fp64      the reference's own arithmetic (float64 throughout, dynamics/quadrotor.py:80-113): the same loop on the fp64 kernel
  configs   BASELINE.json configs 2 (hovering-v0, 4096 envs: per-launch and qs_step_n) and 3 (docking-v0, 2^20 envs);
            config 5's shape (PPO2 rollout-in-the-loop) is under `callers`
  cpu_baseline  kind "reference": the UNMODIFIED Python reference (staged under oracle/_ref by oracle/make_ref.py), one
            single-env process per host core for a fixed wall window (run_moving_docking_ppo2.py:65-67 runs 10 such workers);
            `port` = its fp64 C restatement (oracle/) with OpenMP, `config1` = run_sim_PID.py:43-54 on one core
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# SURVEY.md section 8d / DESIGN.md "Algorithmic bytes" per env-step
ALGO_BYTES = {("docking", "f32"): 405, ("docking", "f64"): 793, ("hovering", "f32"): 225, ("hovering", "f64"): 433}
# what the 40-word (docking) / 20-word (hover) layouts really move per env-step (state R+W, actions, obs, reward, done)
LAYOUT_BYTES = {("docking", "f32"): 389, ("docking", "f64"): 777, ("hovering", "f32"): 233, ("hovering", "f64"): 465}


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
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="wall window of the unmodified-reference workers")
    ap.add_argument("--port-seconds", type=float, default=5.0, help="timed window of the C port inside cpu_baseline")
    ap.add_argument("--ref-min-seconds", type=float, default=2.0, help="--impl reference: minimum timed duration")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-callers", action="store_true", help="skip the short PPO2-rollout measurement reported under 'callers'")
    ap.add_argument("--no-extras", action="store_true", help="skip the fp64 and configs 2/3 sub-records")
    return ap.parse_args()


def family(env_id):
    return "hovering" if env_id.startswith("hover") else "docking"


def workload_config(args):
    """Identical for both arms: the reference arm runs THIS workload on the host."""
    n = args.envs_per_gpu
    return {"workload": "%s fused step (PID/vel target controller + 2x Drone.step + state2rel + reward/done + Philox auto-reset)" % args.env,
            "env": args.env, "envs_per_gpu": n, "actions": "synthetic U(-1,1)^4 fp32, 8 pre-generated tensors cycled",
            "randomize_init": "episode (Philox4x32-10, +-0.5 m)",
            "l2": "inputs larger than L2, no flush: per step the kernel streams 168 MB of state in and out plus one of 8 action "
                  "tensors (134 MB) and one of 8 rotating output sets (448 MB) -- 126 MB L2",
            "parallelism": "env-sharded, dp%d, no collective on the step path" % args.gpus}


# ------------------------------------------------------------------------------------------------ CPU legs (test infrastructure)
def cpu_port_run(env_id, n_envs, steps, warmup, min_seconds=0.0, threads=None):
    """Times the fp64 C port of the reference (oracle/quadsim_oracle.c) with OpenMP over envs: `warmup` untimed steps (plus
    as many as it takes to run the thread team for 0.5 s), then passes of `steps` steps until `min_seconds` have elapsed."""
    import numpy as np
    from oracle import oracle as O
    threads = threads or os.cpu_count() or 1
    env = O.OracleVecEnv(env_id, n_envs, seed=0, randomize=O.RAND_EPISODE, nthreads=threads)
    env.reset()
    rng = np.random.RandomState(0)
    lo = 0.0 if env_id.startswith("hover") else -1.0
    acts = [rng.uniform(lo, 1, (n_envs, 4)).astype(np.float32).astype(np.float64) for _ in range(4)]
    t0 = time.perf_counter()
    k = 0
    while k < warmup or time.perf_counter() - t0 < 0.5:
        env.step(acts[k % 4])
        k += 1
    done_steps = 0
    t0 = time.perf_counter()
    while True:
        for _ in range(steps):
            env.step(acts[done_steps % 4])
            done_steps += 1
        el = time.perf_counter() - t0
        if el >= min_seconds:
            break
    env.close()
    return {"value": n_envs * done_steps / el, "seconds": el, "steps": done_steps, "envs": n_envs, "cores": threads, "warmup_steps": k}


def run_reference(args):
    """--impl reference: the reference's algorithm on the host CPU, on OUR arm's config (2^20 envs).  The reference is pure
    Python (~400 env-steps/s/core: K steps of 2^20 envs would take days), so the arm runs its C restatement (oracle/,
    pinned to the reference's own outputs by tests/test_oracle_golden.py) with every host thread; the unmodified Python
    reference itself is timed in the GPU arm's cpu_baseline leg (kind "reference").  Each step is one pass over all 2^20
    envs; the K steps are repeated until the timed region is at least --ref-min-seconds long."""
    threads = os.cpu_count() or 1
    # torchrun exports OMP_NUM_THREADS=1 to every rank; this arm is meant to use all host cores (libgomp reads the
    # variable when it is first loaded, which happens below)
    os.environ["OMP_NUM_THREADS"] = str(threads)
    os.environ.setdefault("OMP_PROC_BIND", "false")
    n = args.envs_per_gpu
    res = cpu_port_run(args.env, n, args.steps, args.warmup, min_seconds=args.ref_min_seconds, threads=threads)
    sample = ("%d envs x %d steps (%d passes of --steps, %.2f s timed after %d warm-up steps) of %s, fp64 C restatement of the reference, OpenMP %d threads"
              % (n, res["steps"], res["steps"] // max(args.steps, 1), res["seconds"], res["warmup_steps"], args.env, threads))
    line = {
        "impl": "reference", "metric": "env_steps_per_sec", "value": res["value"], "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * res["seconds"] / res["steps"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": res["value"], "unit": "env-steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": res["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "timed_steps": res["steps"], "timed_seconds": res["seconds"],
        "note": "a host rate does not depend on the GPU count: at N > 1 rank 0 alone runs the same 2^20-env sample and the other ranks exit "
                "before importing anything; the unmodified Python reference is timed under cpu_baseline of the GPU arm",
    }
    emit(line)


def cpu_baseline(args):
    """Rank 0: the unmodified Python reference on every host core (kind "reference"), its C port beside it, and config 1."""
    threads = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(threads)
    os.environ.setdefault("OMP_PROC_BIND", "false")
    port = cpu_port_run(args.env, args.envs_per_gpu, 4, 2, min_seconds=args.port_seconds, threads=threads)  # the reference arm's own sample
    port_rec = {"value": port["value"], "unit": "env-steps/s", "cores": threads, "kind": "port",
                "sample": "%d envs x %d steps of %s (%.1f s), fp64 C restatement of the reference (oracle/), OpenMP" % (port["envs"], port["steps"], args.env, port["seconds"])}
    try:
        from oracle import make_ref, ref_bench
        if not make_ref.staged():
            raise RuntimeError("oracle/_ref is not staged (run __graft_entry__.build() or `python oracle/make_ref.py` where /root/reference exists)")
        ref = ref_bench.run_workers(args.env, args.cpu_seconds, threads)
        out = {"value": ref["value"], "unit": "env-steps/s", "cores": ref["procs"], "kind": "reference", "per_core": ref["per_core"],
               "sample": "the unmodified Python reference (oracle/_ref: the six files of SURVEY 8a): %d single-env %s processes x %.1f s wall each, "
                         "env.step + reset-on-done with U(-1,1)^4 actions, %d env-steps and %d episodes in total (run_moving_docking_ppo2.py:65-67 runs 10 such workers)"
                         % (ref["procs"], args.env, ref["seconds"], ref["steps"], ref["episodes"]),
               "port": port_rec}
        try:
            c1 = ref_bench.run_config1()
            out["config1"] = {"what": "BASELINE config 1: run_sim_PID.py:43-54, one quadrotor + PID, 2000 steps, one process, unmodified reference",
                              "steps_per_s": c1["steps_per_s"], "seconds": c1["seconds"], "final_pos": c1["final_pos"]}
        except Exception as exc:  # noqa: BLE001
            out["config1"] = {"error": str(exc)[:200]}
        return out
    except Exception as exc:  # numpy/scipy or the staged files missing on this box: say so, report the port
        port_rec["reference_unavailable"] = str(exc)[:300]
        return port_rec


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
    """(dram bytes per launch of the docking-v2 step kernel at 2^20 envs, source file) from the committed ncu --set full
    capture (profiles/roofline_traffic.json) -- a constant of the kernel, NOT measured in this run -- or (None, None)."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        return d.get(dtype), "profiles/" + str(d.get(dtype + "_source"))
    except Exception:
        return None, None


def kernel_name(env_id, dtype):
    t = "float" if dtype == "f32" else "double"
    if env_id.startswith("hover"):
        return "hover_step_kernel<%s>" % t
    pipe = os.environ.get("QUADSIM_B200_KERNEL", "pipe" if dtype == "f64" else "direct") == "pipe"
    return "%s<%s,%s>" % ("dock_step_pipe_kernel" if pipe else "dock_step_kernel", t, env_id)


def roofline_record(env_id, dtype, n, kernel_ms, steps):
    peak, peak_src = measured_peak()
    key = (family(env_id), dtype)
    algo, layout = ALGO_BYTES[key], LAYOUT_BYTES[key]
    achieved = algo * n / (kernel_ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
            "kernel": kernel_name(env_id, dtype), "algorithmic_bytes_per_env_step": algo, "layout_bytes_per_env_step": layout,
            "layout_GBps": layout * n / (kernel_ms * 1e-3) / 1e9, "kernel_ms": kernel_ms, "peak_source": peak_src,
            "timing": "CUDA events on the launch stream around the %d timed launches / %d" % (steps, steps)}


# ------------------------------------------------------------------------------------------------ GPU arm
def timed_step_loop(env_id, n, tdt, dev, rank, world, steps, warmup, D, sampler=None, n_out=8):
    """W warm-up steps, then K steps between a barrier + synchronize on both sides, CUDA events on the launch stream.
    Returns (ms of this rank, max over ranks, the env -- still open, statistics cleared before the timed region)."""
    import torch
    import quadsim_b200 as qs
    env = qs.make_vec(env_id, n, dtype=tdt, device=dev, seed=0, env_id0=rank * n, randomize="episode" if not env_id.startswith("hover") else "once",
                      want_info=False, reuse_buffers=n_out)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    lo = 0.0 if env_id.startswith("hover") else -1.0
    acts = [(torch.rand((n, 4), generator=gen, device=dev, dtype=torch.float32) * (1 - lo) + lo).to(tdt) for _ in range(8)]
    env.reset()
    for k in range(max(warmup, 3)):
        env.step(acts[k % 8])
    torch.cuda.synchronize(dev)
    env.stats()  # clear warm-up statistics
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize(dev)
    if sampler is not None:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for k in range(steps):
        env.step(acts[k % 8])
    ev1.record()
    torch.cuda.synchronize(dev)
    if world > 1:
        torch.distributed.barrier()
    ms_local = ev0.elapsed_time(ev1)
    return ms_local, D.max_over_ranks(ms_local, dev), env, acts


def variant_record(env_id, dtype, n, rank, dev, world, args, D, what):
    import torch
    tdt = torch.float32 if dtype == "f32" else torch.float64
    steps = args.steps
    ms_local, ms, env, _ = timed_step_loop(env_id, n, tdt, dev, rank, world, steps, args.warmup, D)
    env.close()
    del env
    torch.cuda.empty_cache()
    return {"what": what, "env": env_id, "dtype": dtype, "envs_per_gpu": n, "steps": steps, "value": world * n * steps / (ms * 1e-3),
            "unit": "env-steps/s", "ms_per_step": ms / steps, "roofline": roofline_record(env_id, dtype, n, ms_local / steps, steps)}


def config2_record(args, rank, dev, D):
    """BASELINE config 2: hovering-v0, 4096 envs, random actions -- latency-bound: us/step, per launch and with qs_step_n."""
    import torch
    import quadsim_b200 as qs
    n, T = 4096, 1000
    out = {"what": "BASELINE config 2: hovering-v0 with random U(0,1)^4 actions, 4096 envs on one GPU (latency-bound: us/step, not roofline)", "envs": n}
    for dtype, tdt in (("f32", torch.float32), ("f64", torch.float64)):
        env = qs.make_vec("hovering-v0", n, dtype=tdt, device=dev, seed=0, env_id0=rank * n, randomize="once", want_info=False, reuse_buffers=8)
        gen = torch.Generator(device=dev).manual_seed(7)
        acts = torch.rand((T, n, 4), generator=gen, device=dev, dtype=torch.float32).to(tdt)
        env.reset()
        for k in range(50):
            env.step(acts[k])
        torch.cuda.synchronize(dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        for k in range(T):
            env.step(acts[k])
        ev[1].record()
        env.step_many(acts)  # untimed: warms the multi-step kernel and the allocator (T x N x 13 outputs)
        torch.cuda.synchronize(dev)
        ev[2].record()
        env.step_many(acts)
        ev[3].record()
        torch.cuda.synchronize(dev)
        out["binding"] = "PyTorch extension (csrc/qs_torch_ext.cpp)" if env._bound is not None else "ctypes"
        out[dtype] = {"per_launch_us_per_step": 1e3 * ev[0].elapsed_time(ev[1]) / T, "step_n_us_per_step": 1e3 * ev[2].elapsed_time(ev[3]) / T,
                      "step_n_env_steps_per_s": n * T / (ev[2].elapsed_time(ev[3]) * 1e-3), "steps": T}
        env.close()
    return out


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

    sampler = ClockSampler(local_rank)
    ms_local, ms, env, acts = timed_step_loop(args.env, n, tdt, dev, rank, world, args.steps, args.warmup, D, sampler=sampler)
    clocks = sampler.result()
    stats = env.stats(all_reduce=True)  # the one collective of a rollout: 5 doubles
    env.close()
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
               "api": "QuadVecEnv.step_host -> qs_step_host (pinned numpy in/out, %d graduated chunks pipelined over 4 streams forked from the caller's stream; "
                      "obs by DMA, reward/done written by the kernel into the pinned buffers)" % chunks,
               "checksum_reward": float(h_rew.sum())}
        henv.close()
        del d_buf, h_buf
    del acts
    torch.cuda.empty_cache()

    # ---- the path's caller (SURVEY 8f-1 / BASELINE config 5), a few seconds: PPO2 rollout with the tcgen05 policy in the loop
    callers = None
    if not args.no_callers and args.dtype == "f32" and not args.env.startswith("hover"):
        callers = measure_callers(args, rank, dev, world, D)
        torch.cuda.empty_cache()

    # ---- the reference's own precision and BASELINE configs 2 / 3, same loop, same run
    fp64 = configs = None
    if not args.no_extras and not args.env.startswith("hover"):
        other = "f64" if args.dtype == "f32" else "f32"
        fp64 = variant_record(args.env, other, n, rank, dev, world, args, D,
                              "the same workload in %s (%s)" % (other, "the reference's own arithmetic: float64 throughout, dynamics/quadrotor.py:80-113; "
                                                                "done flags bit-exact, 1e-9 parity bar" if other == "f64" else "the fp32 kernel"))
        configs = {"config3": variant_record("docking-v0", args.dtype, n, rank, dev, world, args, D,
                                             "BASELINE config 3: docking-v0 (PID-hovering target), fused PID + dynamics + reward, per-env random actions"),
                   "hovering_2p20": variant_record("hovering-v0", args.dtype, n, rank, dev, world, args, D,
                                                   "hovering-v0 at the headline size (dynamics-only kernel, 225 B per env-step in fp32)")}
        if rank == 0:
            configs["config2"] = config2_record(args, rank, dev, D)

    if rank != 0:
        return
    kernel_ms = ms_local / args.steps
    roofline = roofline_record(args.env, args.dtype, n, kernel_ms, args.steps)
    if args.env == "docking-v2" and n == 1 << 20:
        traffic, src = ncu_traffic(args.dtype)
        if traffic:
            roofline["traffic"] = traffic
            roofline["traffic_source"] = "%s: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` launch of this kernel at this size, a committed constant, NOT measured in this run" % src
            roofline["dram_GBps"] = traffic / (kernel_ms * 1e-3) / 1e9
            roofline["dram_frac"] = roofline["dram_GBps"] / roofline["peak"]
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
        cpu = cpu_baseline(args)
    line = {
        "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": workload_config(args),
        "clocks": clocks, "e2e": e2e, "gpu_launches": args.steps, "roofline": roofline, "cpu_baseline": cpu,
        "episode_stats": {k: stats[k] for k in ("episodes", "mean_length", "mean_return", "docked_steps", "nonfinite_steps")},
        "fp64" if args.dtype == "f32" else "fp32": fp64, "configs": configs, "callers": callers,
    }
    emit(line)


def measure_callers(args, rank, dev, world, D, n_steps=16, rollouts=5):
    """Not the headline: the rollout-in-the-loop numbers of the path's caller (BASELINE config 5's shape), measured in the same
    run.  qs_rollout (policy forward on tcgen05 + env step, in place) + qs_gae, then one optimisation epoch of the fused PPO2
    update.  CUDA-event timed, max over ranks."""
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
    launches, slices = algo.runner.launches_per_rollout(), algo.runner.n_slices
    # the same rollout through the fused step kernel (one launch per step: policy forward + env step in one CTA)
    algo.runner.n_slices = 0
    for _ in range(2):
        algo.runner.run()
    torch.cuda.synchronize(dev)
    fe = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    fe[0].record()
    for _ in range(rollouts):
        algo.runner.run()
    fe[1].record()
    torch.cuda.synchronize(dev)
    fused_ms = D.max_over_ranks(fe[0].elapsed_time(fe[1]) / rollouts, dev)
    fused_launches = algo.runner.launches_per_rollout()
    env.close()
    rec = {"what": "PPO2 %s rollout-in-the-loop, %d envs/GPU, %d-step rollouts (BASELINE config 5 shape): tcgen05 policy forward + "
                   "fused env step written in place (qs_rollout, two env slices on two streams) + qs_gae; then one epoch of the fused update" % (args.env, n, n_steps),
           "rollout_env_steps_per_s": world * n * n_steps / (roll_ms * 1e-3), "rollout_ms": roll_ms,
           "gpu_launches_per_rollout": launches, "env_slices": slices,
           "fused_step_kernel": {"what": "qs_rollout(n_slices=0): one launch per step (policy_env_step_kernel: env warps inside the policy CTA)",
                                 "rollout_env_steps_per_s": world * n * n_steps / (fused_ms * 1e-3), "rollout_ms": fused_ms,
                                 "gpu_launches_per_rollout": fused_launches},
           "update_samples_per_s": world * n * n_steps / (upd_ms * 1e-3), "update_epoch_ms": upd_ms,
           "update": "hand-written kernels only: qs_permutation, qs_policy_ppo_grad (gradient + tcgen05 split-K weight-gradient contraction), qs_adam_step"}
    del algo, batch
    return rec


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
    if args.impl == "reference":
        # Under torchrun every rank runs this file.  Ranks other than 0 leave HERE -- before numpy, torch or the oracle are
        # imported -- so that nothing competes with rank 0's OpenMP team for the host cores.
        if int(os.environ.get("RANK", "0")) != 0:
            return
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
        return
    run_ours(args)
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
