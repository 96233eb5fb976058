#!/usr/bin/env python
"""Timing-hazard stress for the tensor-core policy kernels (warp-specialised forward, TRAIN gradient kernel): the hand-offs
between issuer and epilogue warps are ordered by mbarriers and tcgen05 fences only, so a missing edge would show up as a
rare wrong row, not as a crash.  Repeats random-size launches (1 .. 400k envs), each checked (a) bit-for-bit against a
second launch on the same inputs, (b) against the torch reference within the bf16 tolerance, while env-step kernels run
concurrently on another stream to perturb the timing; same for the gradient kernel's per-sample outputs.

    python tools/stress_policy.py [--seconds 60]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quadsim_b200 as qs
from quadsim_b200 import ppo
from quadsim_b200.fused_policy import FusedPolicy

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=60.0)
args = ap.parse_args()
torch.manual_seed(0)
pol = ppo.MlpPolicy().cuda()
with torch.no_grad():
    pol.pi.weight.mul_(30.0)
fused = FusedPolicy(pol, seed=3)
env = qs.make_vec("docking-v2", 1 << 19, seed=1, randomize="episode", want_info=False, reuse_buffers=True)
env.reset()
side = torch.cuda.Stream()
acts = torch.rand((1 << 19, 4), device="cuda") * 2 - 1
g = torch.Generator(device="cuda").manual_seed(1)
t0, it, worst, worst_g = time.time(), 0, 0.0, 0.0
while time.time() - t0 < args.seconds:
    n = int(torch.randint(1, 400_000, (1,)).item()) if it % 3 else int(torch.randint(1, 600, (1,)).item())
    obs = torch.randn((n, 12), device="cuda", generator=g) * torch.tensor([2, 2, 2, .5, .5, .5, .3, .3, .3, 1, 1, 1], device="cuda")
    with torch.cuda.stream(side):  # concurrent HBM-bound kernels: co-resident CTAs change every timing in the policy kernel
        for _ in range(3):
            env.step(acts)
    a1 = fused.step(obs, step_idx=it)
    a2 = fused.step(obs, step_idx=it)
    assert all(torch.equal(x, y) for x, y in zip(a1, a2)), ("forward not reproducible", it, n)
    with torch.no_grad():
        mean, value = pol.forward(obs)
    det = fused.step(obs, deterministic=True)
    e = max(float((det[0] - mean).abs().max()), float((det[2] - value).abs().max()))
    worst = max(worst, e)
    assert e < 8e-2, ("forward off", it, n, e)
    if it % 4 == 0 and n > 16:
        B = max(1, n // 2)
        idx = torch.randperm(n, device="cuda")[:B].contiguous()
        batch = dict(obs=obs, actions=a1[0], returns=torch.randn(n, device="cuda"), values=a1[2], neglogp=a1[3])
        g1, _ = fused.ppo_grad(batch, idx)
        ps1 = fused._ws["per_sample"][:B].clone()
        tiles = (B + 127) // 128
        dz1 = fused._ws["dz1"][: tiles * 32768].clone()  # operand image of the weight-gradient contraction (bf16, opaque)
        w2 = g1[pol.pi_fc0.weight].clone()
        g2, _ = fused.ppo_grad(batch, idx)
        assert torch.equal(ps1, fused._ws["per_sample"][:B]) and torch.equal(dz1, fused._ws["dz1"][: tiles * 32768]), ("gradient kernel not reproducible", it, n)
        # the contraction adds split-K partial sums with atomics: reproducible to accumulation order, not bit for bit
        assert torch.allclose(w2, g2[pol.pi_fc0.weight], rtol=1e-3, atol=1e-6 * float(w2.abs().max()) + 1e-12), ("weight-gradient contraction off", it, n)
        assert bool(torch.isfinite(ps1).all())
    it += 1
torch.cuda.synchronize()
print("stress ok: %d iterations in %.0f s, worst forward deviation from fp32 torch %.3e" % (it, time.time() - t0, worst))
