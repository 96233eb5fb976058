#!/usr/bin/env python
"""Small all-paths run meant to be executed under `compute-sanitizer --tool memcheck` (ragged N, every env kind, both
dtypes, info/terminal outputs, reset, export/import, host path, stats, leaf functions)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import quadsim_b200 as qs
from quadsim_b200 import functional as F

for env_id in ("docking-v2", "docking-v0", "hovering-v0"):
    for dtype in (torch.float32, torch.float64):
        n = 1000 + (7 if dtype == torch.float64 else 0)  # not a multiple of 128 or 32
        env = qs.make_vec(env_id, n, dtype=dtype, seed=3, randomize="episode", want_info=True, want_terminal_obs=True)
        env.reset()
        lo = 0.0 if env_id.startswith("hover") else -1.0
        for t in range(90):
            a = (torch.rand((n, 4), device="cuda") * (1 - lo) + lo).to(dtype)
            env.step(a)
        st = env.get_state()
        env.set_state(st)
        env.reset()
        env.step(a.cpu().numpy())  # host path
        print(env_id, dtype, env.stats())
        env.close()
q = torch.randn(300, 4, device="cuda")
F.quat2euler(q); F.quat2rot(q); F.euler2quat(torch.randn(300, 3, device="cuda"))
torch.cuda.synchronize()
print("sanitize smoke done")
