#!/usr/bin/env python
"""The reference's run_expert_record.py (generate_PID_expert_traj, :47-193; main :195-232) on the batched env: record PID-expert
flights of docking-v1 as a stable-baselines ExpertDataset archive (keys actions, obs, rewards, episode_returns,
episode_starts).

    python examples/run_expert_record.py [n_episodes=100] [n_envs=64] [out=./expert_PID/expert_PID_rand_initial]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import quadsim_b200
from quadsim_b200.expert import generate_pid_expert_traj

n_episodes = int(sys.argv[1]) if len(sys.argv) > 1 else 100
n_envs = int(sys.argv[2]) if len(sys.argv) > 2 else 64
out = sys.argv[3] if len(sys.argv) > 3 else "./expert_PID/expert_PID_rand_initial"
os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
env = quadsim_b200.make_vec("gym_docking:docking-v1", n_envs)   # before: gym.make('gym_docking:docking-v1')
data = generate_pid_expert_traj(env, n_episodes=n_episodes, save_path=out)
for key, val in data.items():
    print(key, val.shape)
print("mean episode return %.3f over %d episodes -> %s.npz" % (float(np.mean(data["episode_returns"])), len(data["episode_returns"]), out))
