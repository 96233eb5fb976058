"""Device versions of the reference's leaf functions (utils/transform.py, Drone, controller, state2rel) on batched
torch CUDA tensors, through qs_fn_apply of the C ABI.  Same names and argument meaning as the reference, with a leading
batch dimension; float32 or float64 (taken from the first argument)."""
import ctypes as C

import torch

from . import _capi


def _prep(*ts):
    t0 = ts[0]
    if not t0.is_cuda:
        raise _capi.QuadSimError("quadsim_b200.functional needs CUDA tensors: there is no CPU fallback")
    if t0.dtype not in (torch.float32, torch.float64):
        raise ValueError("float32 or float64 tensors only")
    return [None if t is None else t.to(device=t0.device, dtype=t0.dtype).contiguous() for t in ts]


def _apply(fn, ins, out_widths):
    ins = _prep(*ins)
    t0 = ins[0]
    n = t0.shape[0]
    outs = [torch.empty((n, w), dtype=t0.dtype, device=t0.device) for w in out_widths]
    ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    ins = ins + [None] * (3 - len(ins))
    outs_p = outs + [None] * (2 - len(outs))
    with torch.cuda.device(t0.device):
        st = C.c_void_p(torch.cuda.current_stream(t0.device).cuda_stream)
        _capi.check(_capi.load().qs_fn_apply(fn, _capi.F32 if t0.dtype == torch.float32 else _capi.F64, n,
                                             ptr(ins[0]), ptr(ins[1]), ptr(ins[2]), ptr(outs_p[0]), ptr(outs_p[1]), st))
    return outs


def quat2euler(quat):
    """utils/transform.py:94-120 -- [n,4] (w,x,y,z) -> [n,3] (phi, theta, psi)."""
    return _apply(_capi.FN_QUAT2EULER, [quat], [3])[0]


def euler2quat(euler):
    """utils/transform.py:123-136 -- [n,3] -> [n,4]."""
    return _apply(_capi.FN_EULER2QUAT, [euler], [4])[0]


def quat2rot(quat):
    """utils/transform.py:4-20 (the reference's elementwise-square matrix) -- [n,4] -> [n,3,3]."""
    return _apply(_capi.FN_QUAT2ROT, [quat], [9])[0].view(-1, 3, 3)


def rot2euler(R):
    """utils/transform.py:23-46 -- [n,3,3] -> [n,3]."""
    return _apply(_capi.FN_ROT2EULER, [R.reshape(-1, 9)], [3])[0]


def u_limit(u):
    """Drone.u_limit, dynamics/quadrotor.py:171-187 -- [n,4] -> [n,4]."""
    return _apply(_capi.FN_U_LIMIT, [u], [4])[0]


def drone_step(state, u_prev, u_cmd):
    """Drone.step, dynamics/quadrotor.py:126-144 -- returns (new state[n,13], new u_prev[n,4])."""
    return tuple(_apply(_capi.FN_DRONE_STEP, [state, u_prev, u_cmd], [13, 4]))


def pid(state_des, state_now):
    """controller.PID, controller/PIDController.py:179-185 -- returns (u[n,4], des_q written back[n,4])."""
    return tuple(_apply(_capi.FN_PID, [state_des, state_now], [4, 4]))


def vel_controller(state_des, state_now):
    """controller.vel_controller with state_last == state_now, PIDController.py:106-141 -- (u[n,4], des_q[n,4])."""
    return tuple(_apply(_capi.FN_VEL_CTRL, [state_des, state_now], [4, 4]))


def state2rel(state_chaser, state_target):
    """state2rel with the envs' dock ports, gym_docking/envs/docking_env.py:257-295 -- [n,12]."""
    return _apply(_capi.FN_STATE2REL, [state_chaser, state_target], [12])[0]


def scalar_math(y, x):
    """The kernels' own elementary functions (what stands in for numpy's arctan2/sin/cos/sqrt of the reference) on [n]
    tensors: returns [n,6] = atan2(y,x), sin(y), cos(y), 1/sqrt(|x|), sqrt(|x|), asin(y).  For accuracy tests."""
    return _apply(_capi.FN_MATH, [torch.stack([y, x], dim=1)], [6])[0]


def philox4x32_10(ctr, key):
    """ctr int64/uint32-valued [n,4], key [n,2] -> [n,4] (as int64 holding uint32 values)."""
    c = ctr.to(torch.int64).contiguous()
    k = key.to(torch.int64).contiguous()
    dev = c.device
    to_u32 = lambda t: (t & 0xFFFFFFFF).to(torch.int64).to(torch.int32, non_blocking=False) if False else \
        torch.where(t >= 2 ** 31, t - 2 ** 32, t).to(torch.int32).contiguous()
    c32, k32 = to_u32(c), to_u32(k)
    out = torch.empty_like(c32)
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _capi.check(_capi.load().qs_fn_apply(_capi.FN_PHILOX, 0, c.shape[0], C.c_void_p(c32.data_ptr()),
                                             C.c_void_p(k32.data_ptr()), None, C.c_void_p(out.data_ptr()), None, st))
    return out.to(torch.int64) & 0xFFFFFFFF
