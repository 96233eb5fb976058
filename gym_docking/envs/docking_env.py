"""gym_docking/envs/docking_env.py of the reference: DockingEnv ('docking-v0', :12-254) and the module function state2rel
(:257-295), here backed by the CUDA library."""
import numpy as np

from quadsim_b200.envs import DockingEnv  # noqa: F401


def state2rel(state_chaser, state_target, chaser_dp=None, target_dp=None):
    """Relative observation [dp(3), dv(3), euler(3), euler rates(3)] of two 13-states (docking_env.py:257-295).  The dock-port
    dicts of the reference's signature are accepted and ignored: the ports the envs use (+-0.1 m along body x) are folded
    into the device function (qs::state2rel)."""
    import torch
    from quadsim_b200 import functional as F
    c = torch.as_tensor(np.asarray(state_chaser, dtype=np.float64).reshape(1, 13), device="cuda")
    t = torch.as_tensor(np.asarray(state_target, dtype=np.float64).reshape(1, 13), device="cuda")
    return F.state2rel(c, t)[0].cpu().numpy()
