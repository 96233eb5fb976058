#!/usr/bin/env python
"""The reference's replay script run_trained_moving_docking_ppo2.py:27-64 with the drop-in package: load a stable-baselines
PPO2 archive, fly ONE docking env with model.predict(obs, deterministic=True) until done, print the return.

    python examples/run_trained_moving_docking_ppo2.py model.zip [docking-v2]
(e.g. the reference's trained_model/best_model_v0.zip with docking-v0, or an archive written by quadsim_b200.ppo.PPO2.save)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import quadsim_b200
from quadsim_b200 import ppo

path = sys.argv[1]
env_id = sys.argv[2] if len(sys.argv) > 2 else "docking-v2"
env = quadsim_b200.make("gym_docking:" + env_id)               # before: gym.make('gym_docking:docking-v2')
model = ppo.MlpPolicy.from_sb_zip(path, device="cuda")         # before: PPO2.load(path)
rewards, docked = [], 0
obs = env.reset()
for t in range(1000):
    action = model.predict(torch.as_tensor(obs, dtype=torch.float32, device="cuda")[None])[0].cpu().numpy()
    obs, reward, done, info = env.step(action)
    u = (np.linalg.inv(env.chaser.rotor2control) @ action.flatten() - env.action_mean) / env.action_std  # as the script logs it
    rewards.append(reward)
    docked += int(info["flag_docking"])
    if done:
        break
print("steps %d, return %.4f, docked on %d steps, final relative position %s" % (t + 1, np.sum(rewards), docked, obs[:3]))
