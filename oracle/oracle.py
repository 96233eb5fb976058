"""TEST INFRASTRUCTURE ONLY -- ctypes face of the fp64 CPU oracle (oracle/quadsim_oracle.c).

Importers allowed: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs.
The product package quadsim_b200 never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libquadsim_oracle.so")

HOVER, DOCK_V0, DOCK_V2 = 0, 1, 2
RAND_NONE, RAND_EPISODE, RAND_ONCE = 0, 1, 2
INFO_DONE, INFO_OVERLIMIT, INFO_DOCKING, INFO_OVERTIME = 1, 2, 4, 8
KINDS = {"hovering-v0": HOVER, "docking-v0": DOCK_V0, "docking-v2": DOCK_V2}


def build(force=False):
    src = [os.path.join(HERE, f) for f in ("quadsim_oracle.c", "quadsim_oracle.h", "Makefile")]
    if (not force and os.path.exists(LIB_PATH)
            and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return LIB_PATH
    subprocess.check_call(["make", "-s", "-C", HERE])
    return LIB_PATH


_D3 = C.c_double * 3


class Params(C.Structure):
    """qo_params == the CUDA library's qs_env_params_t, field for field."""
    _fields_ = [("dt", C.c_double), ("gravity", C.c_double), ("mass", C.c_double), ("arm_length", C.c_double),
                ("inertia", _D3), ("motor_lambda", C.c_double), ("att_limit_deg", _D3), ("kp_att", _D3), ("kd_att", _D3),
                ("kp_pos", _D3), ("kd_pos", _D3), ("kp_vel", _D3), ("chaser_ini_pos", _D3), ("target_ini_pos", _D3),
                ("target_des_pos", _D3), ("target_des_vel", _D3), ("dock_port", C.c_double), ("dock_pos_tol", C.c_double),
                ("dock_vel_tol", C.c_double), ("dock_angle_deg", C.c_double), ("range_limit", C.c_double),
                ("min_altitude", C.c_double), ("max_steps", C.c_int32), ("reserved", C.c_int32)]


class _Config(C.Structure):
    _fields_ = [("kind", C.c_int), ("n_envs", C.c_int64), ("seed", C.c_uint64), ("env_id0", C.c_uint64),
                ("randomize", C.c_int), ("half_width", C.c_double), ("auto_reset", C.c_int), ("id_stride", C.c_int64),
                ("params", C.POINTER(Params))]


def default_params(kind):
    if isinstance(kind, str):
        kind = KINDS[kind]
    p = Params()
    lib().qo_default_params(kind, C.byref(p))
    return p


def make_params(kind, overrides):
    """Params with the reference's literals for `kind`, then `overrides` (dict: scalars or 3-sequences) applied."""
    p = default_params(kind)
    for k, v in (overrides or {}).items():
        cur = getattr(p, k)
        if isinstance(cur, (int, float)):
            setattr(p, k, type(cur)(v))
        else:
            cur[:] = [float(x) for x in v]
    return p


_lib = None
_dp = C.POINTER(C.c_double)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.qo_vec_create.restype = C.c_void_p
        L.qo_vec_create.argtypes = [C.POINTER(_Config)]
        L.qo_vec_destroy.argtypes = [C.c_void_p]
        L.qo_vec_obs_dim.argtypes = [C.c_void_p]
        L.qo_vec_field.restype = C.c_void_p
        L.qo_vec_field.argtypes = [C.c_void_p, C.c_int]
        L.qo_vec_reset.argtypes = [C.c_void_p, C.c_void_p]
        L.qo_vec_step.argtypes = [C.c_void_p] + [C.c_void_p] * 6 + [C.c_int]
        L.qo_vec_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.qo_default_params.argtypes = [C.c_int, C.POINTER(Params)]
        L.qo_philox_jitter3.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_float, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _apply(name, n_in, shapes_in, shape_out, *arrs, inplace=()):
    """Row-wise application of a leaf function over [n, w] float64 arrays."""
    L = lib()
    fn = getattr(L, name)
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in arrs]
    n = arrs[0].shape[0]
    outs = [np.zeros((n,) + s) for s in shape_out]
    for i in range(n):
        fn(*[_p(a[i]) for a in arrs], *[_p(o[i]) for o in outs])
    return outs[0] if len(outs) == 1 else outs


def quat2euler(q):
    return _apply("qo_quat2euler", 1, None, [(3,)], q)


def euler2quat(e):
    return _apply("qo_euler2quat", 1, None, [(4,)], e)


def quat2rot(q):
    return _apply("qo_quat2rot", 1, None, [(9,)], q).reshape(-1, 3, 3)


def rot2euler(R):
    return _apply("qo_rot2euler", 1, None, [(3,)], np.asarray(R).reshape(-1, 9))


def u_limit(u):
    return _apply("qo_u_limit", 1, None, [(4,)], u)


def df(s, u):
    return _apply("qo_df", 2, None, [(13,)], s, u)


def state2rel(sc, st):
    return _apply("qo_state2rel", 2, None, [(12,)], sc, st)


def drone_step(s, u_prev, u_cmd):
    """Returns (new_state, new_u_prev); inputs untouched."""
    L = lib()
    s = np.array(s, dtype=np.float64, copy=True)
    up = np.array(u_prev, dtype=np.float64, copy=True)
    uc = np.ascontiguousarray(u_cmd, dtype=np.float64)
    for i in range(s.shape[0]):
        L.qo_drone_step(_p(s[i]), _p(up[i]), _p(uc[i]))
    return s, up


def dock_port(s, d):
    L = lib()
    L.qo_dock_port.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    s = np.ascontiguousarray(s, dtype=np.float64)
    pos, vel = np.zeros((s.shape[0], 3)), np.zeros((s.shape[0], 3))
    for i in range(s.shape[0]):
        L.qo_dock_port(_p(s[i]), d, _p(pos[i]), _p(vel[i]))
    return pos, vel


def pid(des, now):
    """Returns (u, des_after)."""
    L = lib()
    des = np.array(des, dtype=np.float64, copy=True)
    now = np.ascontiguousarray(now, dtype=np.float64)
    u = np.zeros((des.shape[0], 4))
    for i in range(des.shape[0]):
        L.qo_pid(_p(des[i]), _p(now[i]), _p(u[i]))
    return u, des


def vel_controller(des, now, last):
    L = lib()
    des = np.array(des, dtype=np.float64, copy=True)
    now = np.ascontiguousarray(now, dtype=np.float64)
    last = np.ascontiguousarray(last, dtype=np.float64)
    u = np.zeros((des.shape[0], 4))
    for i in range(des.shape[0]):
        L.qo_vel_controller(_p(des[i]), _p(now[i]), _p(last[i]), _p(u[i]))
    return u, des


def philox4x32_10(ctr, key):
    L = lib()
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    L.qo_philox4x32_10(_p(c), _p(k), _p(out))
    return out


def philox_jitter3(seed, env_id, episode, stream, scale):
    out = np.zeros(3, dtype=np.float32)
    lib().qo_philox_jitter3(seed, env_id, episode, stream, scale, _p(out))
    return out


_FIELDS = {"chaser": (0, 13, np.float64), "target": (1, 13, np.float64), "u_c": (2, 4, np.float64),
           "u_t": (3, 4, np.float64), "des_q": (4, 4, np.float64), "last_shaping": (5, 1, np.float64),
           "ep_return": (6, 1, np.float64), "t": (7, 1, np.int64), "episode": (8, 1, np.int64)}


class OracleVecEnv:
    """Batched fp64 CPU environment with stable-baselines auto-reset semantics."""

    def __init__(self, kind, n_envs, seed=0, env_id0=0, randomize=RAND_NONE, half_width=0.0, nthreads=1, auto_reset=True,
                 id_stride=1, params=None):
        """id_stride > 1: env i carries the global id env_id0 + i * id_stride (a strided sample of a larger batch).
        params: dict of qo_params overrides (or a Params); None = the reference's literals."""
        if isinstance(kind, str):
            kind = KINDS[kind]
        self.kind, self.n, self.nthreads = kind, int(n_envs), int(nthreads)
        self.params = params if isinstance(params, Params) else (make_params(kind, params) if params else None)
        cfg = _Config(kind, self.n, seed, env_id0, randomize, half_width, int(auto_reset), int(id_stride),
                      C.pointer(self.params) if self.params is not None else None)
        self._h = lib().qo_vec_create(C.byref(cfg))
        self.obs_dim = lib().qo_vec_obs_dim(self._h)

    def field(self, name):
        """Writable numpy view of a per-env field (state injection / inspection)."""
        fid, w, dt = _FIELDS[name]
        ptr = lib().qo_vec_field(self._h, fid)
        ct = C.c_double if dt is np.float64 else C.c_int64
        arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(self.n * w,))
        return arr.reshape(self.n, w) if w > 1 else arr

    def reset(self):
        obs = np.zeros((self.n, self.obs_dim))
        lib().qo_vec_reset(self._h, _p(obs))
        return obs

    def step(self, actions, want_term_obs=False):
        a = np.ascontiguousarray(actions, dtype=np.float64)
        assert a.shape == (self.n, 4)
        obs = np.zeros((self.n, self.obs_dim))
        rew = np.zeros(self.n)
        done = np.zeros(self.n, dtype=np.uint8)
        info = np.zeros(self.n, dtype=np.uint8)
        term = np.zeros((self.n, self.obs_dim)) if want_term_obs else None
        lib().qo_vec_step(self._h, _p(a), _p(obs), _p(rew), _p(done), _p(info), _p(term) if want_term_obs else None,
                          self.nthreads)
        return (obs, rew, done, info, term) if want_term_obs else (obs, rew, done, info)

    def stats(self):
        out = np.zeros(5)
        lib().qo_vec_stats(self._h, _p(out))
        return out

    def close(self):
        if self._h:
            lib().qo_vec_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class OraclePIDExpert:
    """The reference's PID expert loop (run_expert_record.py:40-139) driving an OracleVecEnv (docking kinds)."""

    def __init__(self, env, alias_ini_state=True):
        self.env, self.alias = env, int(bool(alias_ini_state))
        self.mem = np.zeros((env.n, 8))
        self.mem[:, 4] = 1.0

    def act(self, obs=None):
        """-> actions[N,4]; in alias mode rewrites rows of `obs` (the reset observations of re-seated chasers) in place."""
        L, env = lib(), self.env
        chaser, target, t = env.field("chaser"), env.field("target"), env.field("t")
        act = np.zeros((env.n, 4))
        for i in range(env.n):
            first = int(t[i] == 0)
            if L.qo_pid_expert_reseat(_p(chaser[i]), _p(self.mem[i]), first, self.alias) and obs is not None:
                L.qo_state2rel(_p(chaser[i]), _p(target[i]), _p(obs[i]))
            L.qo_pid_expert_action(_p(chaser[i]), _p(target[i]), _p(self.mem[i]), first, self.alias, _p(act[i]))
        return act


def gae(rewards, values, dones, last_values, last_dones, gamma, lam):
    """Runner._run's GAE loop (ppo2.py:511-521) on [T,N] float32 arrays -> (advs, returns)."""
    r = np.ascontiguousarray(rewards, dtype=np.float32)
    v = np.ascontiguousarray(values, dtype=np.float32)
    d = np.ascontiguousarray(dones, dtype=np.uint8)
    lv = np.ascontiguousarray(last_values, dtype=np.float32)
    ld = np.ascontiguousarray(last_dones, dtype=np.uint8)
    T, n = r.shape
    adv, ret = np.zeros_like(r), np.zeros_like(r)
    L = lib()
    L.qo_gae.argtypes = [C.c_int64, C.c_int64] + [C.c_void_p] * 5 + [C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    L.qo_gae(T, n, _p(r), _p(v), _p(d), _p(lv), _p(ld), gamma, lam, _p(adv), _p(ret))
    return adv, ret
