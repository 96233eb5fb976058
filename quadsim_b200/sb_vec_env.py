"""stable-baselines-shaped adapter: numpy float32 observations, numpy rewards/dones and a list of info dicts, exactly
what `SubprocVecEnv([lambda: gym.make('gym_docking:docking-v2')] * n)` hands to `PPO2` in the reference
(run_moving_docking_ppo2.py:65-67, rl_baselines/ppo2/ppo2.py:472-498) -- but stepped by the CUDA library.

Use it for drop-in runs of the reference's own scripts (tens to thousands of envs).  For millions of envs keep the
tensors on the device with QuadVecEnv instead: a python list of a million dicts is the bottleneck, not the simulator."""
import numpy as np
import torch

from .vec_env import QuadVecEnv

try:  # subclass the real base class when stable-baselines is installed, so isinstance checks in SB pass
    from stable_baselines.common.vec_env import VecEnv as _Base  # type: ignore
except Exception:  # pragma: no cover - not installed in this image
    _Base = object


class SBVecEnv(_Base):
    def __init__(self, env_id="docking-v2", num_envs=10, dtype=torch.float64, info_dicts=True, monitor=False, **kw):
        """monitor=True adds stable-baselines' Monitor key to the info dict of an env that finished an episode:
        info['episode'] = {'r': return, 'l': length} -- the one key the reference's trainer reads from infos
        (rl_baselines/ppo2/ppo2.py:494-497).  The reference's scripts do not wrap their envs in Monitor, hence off by default."""
        self._vec = QuadVecEnv(env_id, num_envs, dtype=dtype, want_info=True, want_terminal_obs=True,
                               want_terminal_state=info_dicts, **kw)
        self.num_envs = self._vec.num_envs
        self.observation_space, self.action_space = self._vec.observation_space, self._vec.action_space
        self.info_dicts = info_dicts
        self.monitor = monitor
        self._ep_ret = np.zeros(self.num_envs, np.float64)
        self._ep_len = np.zeros(self.num_envs, np.int64)
        self._actions = None

    def reset(self):
        self._ep_ret[:] = 0.0
        self._ep_len[:] = 0
        return self._vec.reset().float().cpu().numpy()

    def step_async(self, actions):
        self._actions = torch.as_tensor(np.asarray(actions, dtype=np.float64), device=self._vec.device)

    def step_wait(self):
        obs, rew, done, info = self._vec.step(self._actions)
        obs_np, rew_np, done_np = obs.float().cpu().numpy(), rew.float().cpu().numpy(), done.cpu().numpy()
        infos = [{} for _ in range(self.num_envs)]
        if self.info_dicts:
            # the reference's info dict (docking_env.py:226-229): 'chaser', 'target', 'flag_docking', 'done_overlimit', plus
            # SB's 'terminal_observation' on done.  Like the reference's, the dict of an env that finished in this step
            # describes the TERMINAL step: the kernel auto-resets in place, but writes the pre-reset 13-states of exactly
            # those envs to the terminal_chaser / terminal_target rows (qs_step_extras_t).
            bits = info.bits.cpu().numpy()
            term = info.terminal_observation.float().cpu().numpy()
            st = self._vec.get_state()
            chaser = st["chaser"].cpu().numpy()
            target = st["target"].cpu().numpy() if "target" in st else None
            if done_np.any():
                fin = np.nonzero(done_np)[0]
                chaser[fin] = info.terminal_chaser[torch.as_tensor(fin, device=self._vec.device)].cpu().numpy()
                if target is not None:
                    target[fin] = info.terminal_target[torch.as_tensor(fin, device=self._vec.device)].cpu().numpy()
            for i in range(self.num_envs):
                infos[i] = {"chaser": chaser[i], "flag_docking": bool(bits[i] & 4), "done_overlimit": bool(bits[i] & 2)}
                if target is not None:
                    infos[i]["target"] = target[i]
                if done_np[i]:
                    infos[i]["terminal_observation"] = term[i]
        if self.monitor:
            self._ep_ret += rew.double().cpu().numpy()
            self._ep_len += 1
            for i in np.nonzero(done_np)[0]:
                infos[i]["episode"] = {"r": float(self._ep_ret[i]), "l": int(self._ep_len[i])}
                self._ep_ret[i] = 0.0
                self._ep_len[i] = 0
        return obs_np, rew_np, done_np, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        self._vec.close()

    def seed(self, seed=None):
        return self._vec.seed(seed)

    def render(self, *a, **k):
        return None

    def get_attr(self, attr_name, indices=None):
        return self._vec.get_attr(attr_name, indices)

    def set_attr(self, attr_name, value, indices=None):
        return self._vec.set_attr(attr_name, value, indices)

    def env_method(self, method_name, *args, indices=None, **kwargs):
        return self._vec.env_method(method_name, *args, indices=indices, **kwargs)

    def get_images(self):
        return []
