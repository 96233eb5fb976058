"""`gym_docking.envs` as the reference's scripts import it: the reference's class names, each one the CUDA-backed
single-env facade of quadsim_b200.envs (what 'gym_docking.envs:<Class>' entry points resolve to)."""
from quadsim_b200.envs import ENTRY_POINTS as _ENTRY_POINTS

globals().update({cls.__name__: cls for cls in _ENTRY_POINTS.values()})
__all__ = sorted(cls.__name__ for cls in _ENTRY_POINTS.values())
