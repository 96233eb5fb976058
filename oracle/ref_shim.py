"""TEST INFRASTRUCTURE ONLY -- import shim for the *unmodified* reference (sd196821/QuadSim).

Usable where ``/root/reference`` exists (the build container) or where ``oracle/make_ref.py`` has staged the six
hot-path files under ``oracle/_ref/`` (git-ignored; it travels to the GPU box with the snapshot, so bench.py's
cpu_baseline leg can time the unmodified reference there).  It is used by ``tests/golden/make_golden.py`` (to generate
the committed fixtures), by ``oracle/ref_bench.py`` and by the live-reference tests (skipped when neither tree is
present).  Nothing in ``quadsim_b200/`` may import this module.

The reference's env modules ``import gym`` / ``evdev`` / ``server.pub_server`` at import time
(gym-docking/gym_docking/envs/docking_env.py:1-3, controller/__init__.py:2,
gym-docking/gym_docking/envs/video_docking_env.py:11); none of them is installed here, so tiny stand-ins
are planted in ``sys.modules`` before the reference is imported.  No reference file is modified or copied.
"""
import os
import sys
import types

import numpy as np

STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REF_ROOT = os.environ.get("QUADSIM_REFERENCE_ROOT", "/root/reference")
if not os.path.isdir(os.path.join(REF_ROOT, "dynamics")) and os.path.isdir(os.path.join(STAGED_ROOT, "dynamics")):
    REF_ROOT = STAGED_ROOT  # the GPU box: only the staged copy of the hot-path files exists


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "dynamics"))


def use_staged():
    """Point the shim at oracle/_ref/ even where the full tree exists (bench.py: measure the same files everywhere)."""
    global REF_ROOT
    if _installed:
        raise RuntimeError("ref_shim.use_staged() must be called before install()")
    if not os.path.isdir(os.path.join(STAGED_ROOT, "dynamics")):
        raise RuntimeError("oracle/_ref is not staged: run `python oracle/make_ref.py` where /root/reference exists")
    REF_ROOT = STAGED_ROOT


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.shape = self.low.shape if shape is None else tuple(shape)
        self.dtype = np.dtype(dtype)


class _Env:
    pass


_installed = False


def install():
    """Plant the stubs and put the reference on sys.path (idempotent)."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _mod("evdev", InputDevice=object, categorize=object, ecodes=object, KeyEvent=object)
    seeding = _mod("gym.utils.seeding",
                   np_random=lambda seed=None: (np.random.RandomState(seed), seed))
    spaces = _mod("gym.spaces", Box=_Box)
    utils = _mod("gym.utils", seeding=seeding)
    error = _mod("gym.error")
    logger = _mod("gym.logger", warn=lambda *a, **k: None)
    registration = _mod("gym.envs.registration", register=lambda **kw: None)
    envs = _mod("gym.envs", registration=registration)
    gym = _mod("gym", Env=_Env, spaces=spaces, utils=utils, error=error, logger=logger, envs=envs)
    gym.__path__ = []
    pub = _mod("server.pub_server", pub_server=lambda *a, **k: None)
    srv = _mod("server", pub_server=pub)
    srv.__path__ = []
    sys.path[:0] = [REF_ROOT, os.path.join(REF_ROOT, "gym-docking")]
    for name in [m for m in sys.modules if m == "gym_docking" or m.startswith("gym_docking.")]:
        del sys.modules[name]  # this repository ships a drop-in of the same name: the reference's must be the one imported
    _installed = True


def load():
    """Return a namespace with the reference's hot-path classes and functions."""
    install()
    from dynamics.quadrotor import Drone
    from controller.PIDController import controller
    import utils.transform as transform
    from gym_docking.envs.docking_env import DockingEnv
    from gym_docking.envs import docking_env as docking_env_mod
    from gym_docking.envs.moving_docking_env import MovingDockingEnv
    from gym_docking.envs.hovering_env import HoveringEnv
    from scipy.integrate import RK45
    return types.SimpleNamespace(Drone=Drone, controller=controller, transform=transform,
                                 DockingEnv=DockingEnv, MovingDockingEnv=MovingDockingEnv,
                                 HoveringEnv=HoveringEnv, state2rel=docking_env_mod.state2rel, RK45=RK45)
