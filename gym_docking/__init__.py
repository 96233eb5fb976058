"""Drop-in for the reference's `gym_docking` package (gym-docking/gym_docking/__init__.py:3-21): the same module name and the
same env ids, so the reference's own call sites resolve unchanged --

    gym.make('gym_docking:docking-v2')                      # run_moving_docking_ppo2.py:65-67
    from gym_docking.envs.docking_env import DockingEnv     # direct imports

-- but every env is the CUDA-backed facade of quadsim_b200.envs (N = 1, numpy in / out, float64, no auto-reset).
Registration happens with whichever of gymnasium / gym is installed; without either, `gym_docking.make(id)` and the classes
under gym_docking.envs still work.  'docking-v3' (screen-grab observations of a Unity window, video_docking_env.py) is out of
scope and not registered.
"""
from quadsim_b200.envs import ENTRY_POINTS

ENV_IDS = {"docking-v0": "DockingEnv", "docking-v1": "ImitatingDockingEnv", "docking-v2": "MovingDockingEnv",
           "hovering-v0": "HoveringEnv"}


def _register():
    used = []
    for mod in ("gymnasium", "gym"):
        try:
            reg = __import__(mod + ".envs.registration", fromlist=["register"])
        except Exception:
            continue
        for env_id, cls in ENV_IDS.items():
            try:
                reg.register(id=env_id, entry_point="gym_docking.envs:" + cls)  # the reference's own entry-point strings
            except Exception:
                pass  # already registered
        used.append(mod)
    return used


REGISTERED_WITH = _register()


def make(env_id, **kw):
    """gym.make('gym_docking:<id>') without gym: the facade class for `env_id`."""
    return ENTRY_POINTS[env_id.split(":")[-1]](**kw)
