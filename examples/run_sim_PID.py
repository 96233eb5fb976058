#!/usr/bin/env python
"""The reference's run_sim_PID.py (BASELINE config 1: one quadrotor, PID hover to pos_des = [-0.2, 0.2, 0.2], 2000 steps)
on the CUDA library -- and the same scenario for a batch of set-points at once.  Same names as the reference:
controller.PID -> functional.pid, Drone.step -> functional.drone_step (dynamics/quadrotor.py:126-144,
controller/PIDController.py:179-185).  No plotting: prints the final state the reference plots."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quadsim_b200 import functional as F

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1          # 1 = the reference's single drone
dev, dt = "cuda", torch.float64
state = torch.zeros((n, 13), device=dev, dtype=dt)         # run_sim_PID.py:9-16: at the origin, identity attitude
state[:, 6] = 1.0
u_prev = torch.zeros((n, 4), device=dev, dtype=dt)         # Drone.reset: u = 0 (quadrotor.py:76)
state_des = torch.zeros((n, 13), device=dev, dtype=dt)     # :18-22
state_des[:, 0:3] = torch.tensor([-0.2, 0.2, 0.2], device=dev, dtype=dt)
if n > 1:                                                  # a slightly different set-point per drone in the batched run
    state_des[:, 0:3] += 0.04 * (torch.rand((n, 3), device=dev, dtype=dt, generator=torch.Generator(dev).manual_seed(0)) - 0.5)
state_des[:, 6] = 1.0
for t in range(2000):                                      # :43-54
    u, des_q = F.pid(state_des, state)                     # control.PID(state_des, state_now); it rewrites state_des[6:10] (Q4)
    state_des[:, 6:10] = des_q
    state, u_prev = F.drone_step(state, u_prev, u)         # quad1.step(u)
err = (state[:, 0:3] - state_des[:, 0:3]).norm(dim=1)
print("final state of drone 0:", state[0].cpu().numpy())
print("position error after 2000 steps: max %.3e m over %d drones" % (float(err.max()), n))
