"""Single-environment gym facades with the reference's class names, attributes and return types (numpy, python
float/bool, info dict), for the scripts that drive one env directly (run_trained_moving_docking_ppo2.py:27-64,
run_expert_policy.py:27-80).  Each is a QuadVecEnv with N=1 and gym.Env (no auto-reset) semantics; the arithmetic
runs in the CUDA library, by default in float64 like the reference."""
import numpy as np
import torch

from . import _capi
from .vec_env import ARM_LENGTH, GRAVITY, MASS, ROTOR2CONTROL, QuadVecEnv


class _DroneView:
    """What scripts read from env.chaser / env.target / env.drone (dynamics/quadrotor.py:10-63, :189-211)."""
    mass, gravity, arm_length, dt = MASS, GRAVITY, ARM_LENGTH, 0.02
    rotor2control = ROTOR2CONTROL

    def __init__(self, env, field, ufield):
        self._env, self._field, self._ufield = env, field, ufield

    @property
    def state(self):
        return self._env._vec.get_state()[self._field][0].double().cpu().numpy()

    @property
    def u(self):
        return self._env._vec.get_state()[self._ufield][0].double().cpu().numpy()

    def get_state(self):
        return self.state

    def get_arm_length(self):
        return self.arm_length

    def get_mass(self):
        return self.mass


class _SingleEnv:
    metadata = {"render.modes": ["human"]}
    env_id = None

    def __init__(self, dtype=torch.float64, device=None, seed=0, randomize=None, half_width=None):
        self._vec = QuadVecEnv(self.env_id, 1, dtype=dtype, device=device, seed=seed, randomize=randomize,
                               half_width=half_width, auto_reset=False, want_info=True, want_terminal_obs=False)
        self.observation_space, self.action_space = self._vec.observation_space, self._vec.action_space
        self.action_mean, self.action_std = self._vec.action_mean, self._vec.action_std
        self.done, self.reward = False, 0.0
        self._act = torch.zeros((1, 4), dtype=dtype, device=self._vec.device)

    def reset(self):
        self.done, self.reward = False, 0.0
        return self._vec.reset()[0].double().cpu().numpy()

    def _step(self, action):
        a = np.asarray(action, dtype=np.float64).reshape(1, 4)
        self._act.copy_(torch.from_numpy(a))
        obs, rew, done, info = self._vec.step(self._act)
        self.reward, self.done = float(rew[0]), bool(done[0])
        return obs[0].double().cpu().numpy(), self.reward, self.done, info

    def render(self, mode="human"):
        return None

    def close(self):
        self._vec.close()
        return None

    def seed(self, seed=None):
        self._vec.seed(seed)
        return [seed]


class _DockingBase(_SingleEnv):
    def __init__(self, **kw):
        super().__init__(**kw)
        self.chaser = _DroneView(self, "chaser", "u_chaser")
        self.target = _DroneView(self, "target", "u_target")
        self.chaser_dock_port = np.array([0.1, 0.0, 0.0])   # docking_env.py:38
        self.target_dock_port = np.array([-0.1, 0.0, 0.0])  # docking_env.py:51

    @property
    def state_chaser(self):
        return self.chaser.state

    @property
    def state_target(self):
        return self.target.state

    @property
    def chaser_ini_state(self):
        s = np.zeros(13)
        s[0:3], s[6] = [8, -50, 5], 1.0  # docking_env.py:34-43 (before any jitter)
        return s

    @property
    def target_ini_state(self):
        s = np.zeros(13)
        s[0:3], s[6] = [10, -50, 5], 1.0  # docking_env.py:47-56
        return s

    @property
    def t(self):
        return int(self._vec.get_state()["t"][0])

    def step(self, action):
        obs, r, d, info = self._step(action)
        st = self._vec.get_state()
        return obs, r, d, {"chaser": st["chaser"][0].double().cpu().numpy(), "target": st["target"][0].double().cpu().numpy(),
                           "flag_docking": bool(info.flag_docking[0]), "done_overlimit": bool(info.done_overlimit[0])}


class DockingEnv(_DockingBase):
    """gym_docking.envs.DockingEnv ('docking-v0'), gym_docking/envs/docking_env.py:12-254."""
    env_id = "docking-v0"


class ImitatingDockingEnv(_DockingBase):
    """gym_docking.envs.ImitatingDockingEnv ('docking-v1'): v0 with a once-only U(-0.3,0.3)^3 chaser jitter,
    gym_docking/envs/imitating_docking_env.py:34."""
    env_id = "docking-v1"


class MovingDockingEnv(_DockingBase):
    """gym_docking.envs.MovingDockingEnv ('docking-v2'), gym_docking/envs/moving_docking_env.py:12-219."""
    env_id = "docking-v2"


class HoveringEnv(_SingleEnv):
    """gym_docking.envs.HoveringEnv ('hovering-v0'), gym_docking/envs/hovering_env.py:10-95."""
    env_id = "hovering-v0"

    def __init__(self, **kw):
        super().__init__(**kw)
        self.drone = _DroneView(self, "chaser", "u_chaser")
        self.action_max = self._vec.action_std  # hovering_env.py:42

    @property
    def state(self):
        return self.drone.state

    def step(self, action):
        obs, r, d, _ = self._step(action)
        return obs, r, d, {}


ENTRY_POINTS = {"docking-v0": DockingEnv, "docking-v1": ImitatingDockingEnv, "docking-v2": MovingDockingEnv,
                "hovering-v0": HoveringEnv}
