"""PID expert for the chaser and expert-dataset recording on the batched CUDA env.

The reference flies the chaser with `controller.vel_controller` towards a point 0.2 m behind the target
(`run_expert_policy.py:44-69`) and records those flights as a stable-baselines GAIL / pretrain dataset
(`run_expert_record.py:47-193`, `generate_PID_expert_traj`: keys actions, obs, rewards, episode_returns, episode_starts).
Here the controller is one CUDA kernel over every env of a docking handle (C ABI `qs_expert_action`); observations,
actions and rewards stay on the device until the dataset is assembled."""
import ctypes as C

import numpy as np
import torch

from . import _capi


class PIDExpert:
    """`alias_ini_state=True` reproduces the scripts literally -- their `state_des` IS `env.chaser_ini_state`
    (`run_expert_record.py:41`), so every episode after the first starts from the initial state the controller has been
    writing into, and an episode's first command is plain hover thrust.  False keeps `state_des` private."""

    def __init__(self, env, alias_ini_state=True):
        if env.kind == _capi.HOVER:
            raise ValueError("the PID expert flies the chaser of a docking environment")
        self.env, self.alias = env, bool(alias_ini_state)
        self.mem = torch.zeros((env.num_envs, 8), dtype=env.dtype, device=env.device)  # des vel xyz, pad, des quaternion
        self.mem[:, 4] = 1.0

    def act(self, obs=None):
        """-> actions[N,4] for the env's current state.  With aliasing, `obs` (the tensor the last reset()/step() returned)
        is updated in place for envs whose chaser was re-seated, so that it is the observation the reference records."""
        env = self.env
        env._check_open()
        if obs is not None and (obs.dtype != env.dtype or not obs.is_contiguous() or obs.device != env.device):
            raise ValueError("obs must be the env's own contiguous observation tensor")
        act = torch.empty((env.num_envs, 4), dtype=env.dtype, device=env.device)
        with torch.cuda.device(env.device):
            _capi.check(env.lib.qs_expert_action(env._h, C.c_void_p(self.mem.data_ptr()), int(self.alias), C.c_void_p(act.data_ptr()),
                                                 C.c_void_p(obs.data_ptr()) if obs is not None else None, env._stream()))
        return act

    predict = act


def generate_pid_expert_traj(env, n_episodes=100, save_path=None, alias_ini_state=True, max_steps=None):
    """`generate_PID_expert_traj` (`run_expert_record.py:47-193`) on N batched envs: steps every env with the expert until
    at least `n_episodes` episodes have finished, then lays the COMPLETED episodes out env after env in the reference's flat
    format: `actions [M,4]`, `obs [M,12]`, `rewards [M]`, `episode_starts [M]` (True on an episode's first transition),
    `episode_returns [E]`.  With `save_path` the dict is also written as `<save_path>.npz` (the file SB's ExpertDataset
    loads)."""
    expert = PIDExpert(env, alias_ini_state)
    n = env.num_envs
    per_env = -(-int(n_episodes) // n)
    obs = env.reset()
    rec_obs, rec_act, rec_rew, rec_done = [], [], [], []
    finished = torch.zeros(n, dtype=torch.int64, device=env.device)
    steps = 0
    while int(finished.min()) < per_env and (max_steps is None or steps < max_steps):
        act = expert.act(obs)          # may re-seat chasers and rewrite their rows of `obs` (alias mode)
        rec_obs.append(obs.clone())
        obs, rew, done, _ = env.step(act)
        rec_act.append(act)
        rec_rew.append(rew.clone() if env.reuse_buffers else rew)
        rec_done.append(done.clone() if env.reuse_buffers else done)
        finished += done
        steps += 1
    O, A = torch.stack(rec_obs).cpu().numpy(), torch.stack(rec_act).cpu().numpy()        # [T,N,...]
    Rw, D = torch.stack(rec_rew).cpu().numpy(), torch.stack(rec_done).cpu().numpy().astype(bool)
    out = dict(actions=[], obs=[], rewards=[], episode_starts=[], episode_returns=[])
    for k in range(n):
        ends = np.nonzero(D[:, k])[0][:per_env]
        if len(ends) == 0:
            continue
        last = ends[-1] + 1
        starts = np.zeros(last, dtype=bool)
        starts[0] = True
        starts[ends[:-1] + 1] = True
        out["actions"].append(A[:last, k])
        out["obs"].append(O[:last, k])
        out["rewards"].append(Rw[:last, k])
        out["episode_starts"].append(starts)
        out["episode_returns"].extend(Rw[a:b + 1, k].sum() for a, b in zip(np.r_[0, ends[:-1] + 1], ends))
    data = {k: (np.concatenate(v) if k != "episode_returns" else np.asarray(v, dtype=np.float64)) for k, v in out.items()}
    if save_path is not None:
        np.savez(save_path, **data)
    return data
