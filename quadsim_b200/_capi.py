"""ctypes binding of libquadsim_b200.so (C ABI: include/quadsim_b200.h).

The library is the product: if it cannot be loaded this module raises -- there is no Python/CPU fallback.
"""
import ctypes as C
import os

from . import build as _build

OK = 0
HOVER, DOCK_V0, DOCK_V2 = 0, 1, 2
F32, F64 = 0, 1
RAND_NONE, RAND_EPISODE, RAND_ONCE = 0, 1, 2
INFO_DONE, INFO_OVERLIMIT, INFO_DOCKING, INFO_OVERTIME = 1, 2, 4, 8
(FN_QUAT2EULER, FN_EULER2QUAT, FN_QUAT2ROT, FN_ROT2EULER, FN_U_LIMIT, FN_DRONE_STEP, FN_PID, FN_VEL_CTRL,
 FN_STATE2REL, FN_PHILOX, FN_MATH) = range(11)
ABI_VERSION = 2

# every symbol include/quadsim_b200.h declares (tests check the built library exports all of them)
SYMBOLS = ["qs_abi_version", "qs_last_error", "qs_create", "qs_destroy", "qs_obs_dim", "qs_num_envs", "qs_seed",
           "qs_reset", "qs_step", "qs_step_ex", "qs_step_host", "qs_host_alloc", "qs_host_free", "qs_export_state",
           "qs_import_state", "qs_get_stats", "qs_fn_apply", "qs_gae", "qs_policy_create", "qs_policy_destroy",
           "qs_policy_set_weights", "qs_policy_step", "qs_env_dtype", "qs_env_id0", "qs_rollout",
           "qs_step_range", "qs_expert_action", "qs_peek_state", "qs_policy_ppo_grad", "qs_step_n", "qs_default_params",
           "qs_env_device", "qs_adam_step", "qs_permutation", "qs_policy_set_precision",
           "qs_policy_pack_rollout"]
POLICY_BF16, POLICY_FP32 = 0, 1


_D3 = C.c_double * 3


class EnvParams(C.Structure):
    """qs_env_params_t: the reference's model literals (quadrotor.py:10-63, PIDController.py:10-44, env __init__s)."""
    _fields_ = [("dt", C.c_double), ("gravity", C.c_double), ("mass", C.c_double), ("arm_length", C.c_double),
                ("inertia", _D3), ("motor_lambda", C.c_double), ("att_limit_deg", _D3), ("kp_att", _D3), ("kd_att", _D3),
                ("kp_pos", _D3), ("kd_pos", _D3), ("kp_vel", _D3), ("chaser_ini_pos", _D3), ("target_ini_pos", _D3),
                ("target_des_pos", _D3), ("target_des_vel", _D3), ("dock_port", C.c_double), ("dock_pos_tol", C.c_double),
                ("dock_vel_tol", C.c_double), ("dock_angle_deg", C.c_double), ("range_limit", C.c_double),
                ("min_altitude", C.c_double), ("max_steps", C.c_int32), ("reserved", C.c_int32)]

    def as_dict(self):
        out = {}
        for name, ct in self._fields_:
            v = getattr(self, name)
            out[name] = list(v) if ct is _D3 else v
        out.pop("reserved")
        return out


class Config(C.Structure):
    _fields_ = [("env_kind", C.c_int32), ("dtype", C.c_int32), ("n_envs", C.c_int64), ("seed", C.c_uint64),
                ("env_id0", C.c_uint64), ("randomize", C.c_int32), ("device", C.c_int32), ("half_width", C.c_double),
                ("auto_reset", C.c_int32), ("use_params", C.c_int32), ("params", EnvParams)]


class StepExtras(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("info", "terminal_obs", "terminal_chaser", "terminal_target")]


class StateArrays(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("chaser", "target", "u_chaser", "u_target", "des_q", "last_shaping",
                                          "ep_return", "t", "episode")]


class PolicyWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("shared_w", "shared_b", "pi_fc_w", "pi_fc_b", "vf_fc_w", "vf_fc_b", "pi_w", "pi_b",
                                          "vf_w", "vf_b", "logstd")]


class RolloutBuffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("obs", "actions", "values", "neglogp", "rewards", "dones", "clipped")]


class PPOGrad(C.Structure):
    _fields_ = ([(n, C.c_void_p) for n in ("obs", "actions", "returns", "values", "neglogp", "idx")] + [("batch", C.c_int64), ("adv_stats", C.c_void_p)]
                + [(n, C.c_float) for n in ("cliprange", "cliprange_vf", "vf_coef")]
                + [(n, C.c_void_p) for n in ("a1", "h", "hpv", "dz2", "dz1", "dout", "per_sample", "col_sums", "g1", "g2", "g3", "packed_rows")])


class AdamTensor(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("n", C.c_int64)]


class QuadSimError(RuntimeError):
    pass


_lib = None


def load():
    """Load (building first if the in-tree .so is missing or stale and nvcc is present)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("QUADSIM_B200_LIB") or _build.LIB_PATH  # override: kernel-variant experiments (tools/kbench.py)
    if path == _build.LIB_PATH and not _build.up_to_date():
        try:
            path = _build.build()
        except Exception as exc:  # no nvcc on this box: use the shipped .so if there is one
            if not os.path.exists(path):
                raise QuadSimError("libquadsim_b200.so is not built and cannot be built here (%s); run "
                                   "`python -m quadsim_b200.build`. There is no CPU fallback." % exc) from exc
    L = C.CDLL(path)
    vp, i64, u64 = C.c_void_p, C.c_int64, C.c_uint64
    L.qs_abi_version.restype = C.c_int
    L.qs_last_error.restype = C.c_char_p
    L.qs_create.argtypes = [C.POINTER(vp), C.POINTER(Config)]
    L.qs_destroy.argtypes = [vp]
    L.qs_obs_dim.argtypes = [vp]
    L.qs_num_envs.argtypes = [vp]
    L.qs_num_envs.restype = i64
    L.qs_seed.argtypes = [vp, u64]
    L.qs_reset.argtypes = [vp, vp, vp]
    L.qs_step.argtypes = [vp, vp, vp, vp, vp, vp]
    L.qs_step_ex.argtypes = [vp, vp, vp, vp, vp, C.POINTER(StepExtras), vp]
    L.qs_step_host.argtypes = [vp, vp, vp, vp, vp, vp]
    L.qs_default_params.argtypes = [C.c_int, C.POINTER(EnvParams)]
    L.qs_env_device.argtypes = [vp]
    L.qs_host_alloc.argtypes = [C.POINTER(vp), u64]
    L.qs_host_free.argtypes = [vp]
    L.qs_export_state.argtypes = [vp, C.POINTER(StateArrays), vp]
    L.qs_import_state.argtypes = [vp, C.POINTER(StateArrays), vp]
    L.qs_get_stats.argtypes = [vp, C.POINTER(C.c_double), vp]
    L.qs_fn_apply.argtypes = [C.c_int, C.c_int, i64, vp, vp, vp, vp, vp, vp]
    L.qs_policy_create.argtypes = [C.POINTER(vp), C.c_int]
    L.qs_policy_destroy.argtypes = [vp]
    L.qs_policy_set_weights.argtypes = [vp, C.POINTER(PolicyWeights), vp]
    L.qs_policy_set_precision.argtypes = [vp, C.c_int]
    L.qs_policy_step.argtypes = [vp, i64, vp, u64, u64, C.c_uint32, C.c_int, vp, vp, vp, vp, vp]
    L.qs_env_dtype.argtypes = [vp]
    L.qs_env_id0.argtypes = [vp]
    L.qs_env_id0.restype = u64
    L.qs_rollout.argtypes = [vp, vp, i64, C.POINTER(RolloutBuffers), u64, C.c_uint32, C.c_int, vp]
    L.qs_step_range.argtypes = [vp, i64, i64, vp, vp, vp, vp, vp]
    L.qs_expert_action.argtypes = [vp, vp, C.c_int, vp, vp, vp]
    L.qs_peek_state.argtypes = [vp, vp, i64, vp, vp, vp, vp]
    L.qs_policy_ppo_grad.argtypes = [vp, C.POINTER(PPOGrad), vp]
    L.qs_step_n.argtypes = [vp, i64, vp, vp, vp, vp, vp]
    L.qs_adam_step.argtypes = [C.POINTER(AdamTensor), C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, i64, vp, C.c_float, vp, vp]
    L.qs_permutation.argtypes = [i64, u64, vp, vp]
    L.qs_policy_pack_rollout.argtypes = [vp, vp, vp, vp, vp, i64, vp, vp]
    L.qs_gae.argtypes = [i64, i64, vp, vp, vp, vp, vp, C.c_double, C.c_double, vp, vp, vp]
    if L.qs_abi_version() != ABI_VERSION:
        raise QuadSimError("libquadsim_b200.so ABI %d != expected %d; rebuild" % (L.qs_abi_version(), ABI_VERSION))
    _lib = L
    return L


_ext = False  # False: not tried yet; None: unavailable


def load_ext():
    """The PyTorch-extension face of the library (csrc/qs_torch_ext.cpp: tensors in, PyTorch's current stream, plain pointers
    to the C ABI) used on the per-step path, or None when it is neither built nor buildable here -- the ctypes binding above
    reaches the same entry points, only with a higher per-call cost."""
    global _ext
    if _ext is False:
        _ext = None
        if os.environ.get("QUADSIM_B200_NO_EXT") != "1":
            try:
                import importlib.util
                import torch  # noqa: F401  (libc10 / libtorch must be loaded first)
                load()
                path = _build.ext_path()
                if not _build.ext_up_to_date():
                    try:
                        path = _build.build_ext()
                    except Exception:
                        if not os.path.exists(path):
                            raise
                spec = importlib.util.spec_from_file_location(_build.EXT_NAME, path)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                if mod.abi_version() == ABI_VERSION:
                    _ext = mod
            except Exception:
                _ext = None
    return _ext


def default_params(kind):
    """EnvParams filled with the reference's literals for `kind` (qs_default_params)."""
    p = EnvParams()
    check(load().qs_default_params(kind, C.byref(p)))
    return p


def check(rc):
    if rc != OK:
        raise QuadSimError("quadsim_b200 error %d: %s" % (rc, load().qs_last_error().decode()))
