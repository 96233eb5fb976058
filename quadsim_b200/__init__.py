"""quadsim_b200 -- B200-native batched drop-in for the hot path of sd196821/QuadSim (gym_docking envs).

    import quadsim_b200 as qs
    env = qs.make_vec("docking-v2", num_envs=1 << 20, randomize="episode")   # device tensors in/out
    obs = env.reset(); obs, rew, done, info = env.step(actions)
    single = qs.make("docking-v2")                                              # reference-style gym.Env, N=1

All arithmetic runs in hand-written CUDA (quadsim_b200/csrc) behind the C ABI in include/quadsim_b200.h.
There is no CPU fallback: without the built library or without a GPU, constructing an env raises.
"""
from ._capi import QuadSimError  # noqa: F401
from .vec_env import ENV_IDS, QuadVecEnv  # noqa: F401

__version__ = "0.1.0"


def make_vec(env_id="docking-v2", num_envs=1, **kw):
    """Batched, device-resident VecEnv (see QuadVecEnv)."""
    return QuadVecEnv(env_id, num_envs, **kw)


def make(env_id, **kw):
    """gym.make('gym_docking:docking-v2')-style constructor for the single-env facades (quadsim_b200.envs)."""
    from .envs import ENTRY_POINTS
    return ENTRY_POINTS[env_id.split(":")[-1]](**kw)


def register_with_gym():
    """Register the reference's env ids (gym-docking/gym_docking/__init__.py:3-21) with gymnasium or gym if either is
    installed; returns the module used or None."""
    for mod in ("gymnasium", "gym"):
        try:
            reg = __import__(mod + ".envs.registration", fromlist=["register"])
        except Exception:
            continue
        for env_id, cls in (("docking-v0", "DockingEnv"), ("docking-v1", "ImitatingDockingEnv"),
                            ("docking-v2", "MovingDockingEnv"), ("hovering-v0", "HoveringEnv")):
            try:
                reg.register(id=env_id, entry_point="quadsim_b200.envs:" + cls)
            except Exception:
                pass
        return mod
    return None
