"""gym_docking/envs/hovering_env.py of the reference: HoveringEnv ('hovering-v0', :10-95), CUDA-backed."""
from quadsim_b200.envs import HoveringEnv  # noqa: F401
