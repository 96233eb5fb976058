"""gym_docking/envs/moving_docking_env.py of the reference: MovingDockingEnv ('docking-v2', :12-219), CUDA-backed."""
from gym_docking.envs.docking_env import state2rel  # noqa: F401  (the reference defines it again at :222-260)
from quadsim_b200.envs import MovingDockingEnv  # noqa: F401
