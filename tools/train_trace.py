#!/usr/bin/env python
"""Timeline of the PPO2 gradient kernel (policy_kernel<TRAIN>) from SM-clock stamps (a -DQS_WS_TRACE=1 build): mean clocks
between the phase boundaries of a tile for thread 0 of CTA 0 / group 0.

    QUADSIM_B200_LIB=variants/train_trace.so python tools/train_trace.py
"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import quadsim_b200 as qs
from quadsim_b200 import _capi, ppo

n, T = 1 << 20, 16
env = qs.make_vec("docking-v2", n, seed=0, randomize="episode", want_info=False, reuse_buffers=True)
algo = ppo.PPO2(env, n_steps=T, nminibatches=10, noptepochs=1, fused_policy=True)
batch = algo.runner.run()
idx = algo.fused.permutation(n * T, 123)[: n * T // 10].contiguous()
for _ in range(2):
    algo.fused.ppo_grad(batch, idx, 0.2, 0.2, 0.5, 0.0, want_per_sample=False)
torch.cuda.synchronize()
lib = _capi.load()
TILES, SLOTS = 28, 24
buf = np.zeros(TILES * SLOTS, np.int64)
lib.qs_debug_ws_trace.restype = C.c_int
lib.qs_debug_ws_trace.argtypes = [C.c_void_p, C.c_int]
assert lib.qs_debug_ws_trace(buf.ctypes.data_as(C.c_void_p), buf.size) == buf.size
t = buf.reshape(TILES, SLOTS).astype(np.float64)[2:26]
names = ["loop top (A1 built by previous stamps)", "A1 + a1 image stored", "group barrier", "layer 1 issued, gathers issued", "layer 1 done",
         "epilogue 1 (h -> A2 + image)", "group barrier", "layer 2 done (issue + wait)", "epilogue 2 (relu, pack to TMEM, hpv image)", "group barrier",
         "layer 3 done (issue + wait)", "loss, dOut", "backward pi: barrier + MMA + wait", "backward pi: dZ2 pack + image", "backward vf: barrier + MMA + wait",
         "backward vf: dZ2 pack + image", "group barrier", "dA2 done (issue + wait)"]
out = {"period_clk": float(np.mean(np.diff(t[:, 0]))), "phases": []}
for k in range(1, 18):
    out["phases"].append((names[k], round(float(np.mean(t[:, k] - t[:, k - 1])), 1)))
out["phases"].append(("dZ1 pack + image, loop", round(float(np.mean(t[1:, 0] - t[:-1, 17])), 1)))
print(json.dumps(out, indent=1))
