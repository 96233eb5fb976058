"""gym_docking/envs/imitating_docking_env.py of the reference: ImitatingDockingEnv ('docking-v1'), CUDA-backed."""
from gym_docking.envs.docking_env import state2rel  # noqa: F401
from quadsim_b200.envs import ImitatingDockingEnv  # noqa: F401
