"""TEST-ONLY ctypes face of tests/host_harness.cpp: the kernels' arithmetic (qs_core.cuh) compiled for the host."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libqs_host_harness.so")
SRC = [os.path.join(HERE, "host_harness.cpp"), os.path.join(HERE, "..", "quadsim_b200", "csrc", "qs_core.cuh"),
       os.path.join(HERE, "..", "include", "quadsim_b200.h")]
KINDS = {"hovering-v0": 0, "docking-v0": 1, "docking-v2": 2}
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or any(os.path.getmtime(LIB) < os.path.getmtime(s) for s in SRC):
            os.makedirs(os.path.dirname(LIB), exist_ok=True)
            # -ffp-contract=fast mirrors nvcc's default FMA contraction as far as the host ISA allows
            subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fno-fast-math",
                                   "-o", LIB, SRC[0]])
        L = C.CDLL(LIB)
        L.hh_create.restype = C.c_void_p
        L.hh_create.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_uint64, C.c_uint64, C.c_int, C.c_double, C.c_int, C.c_void_p]
        L.hh_destroy.argtypes = [C.c_int, C.c_void_p]
        L.hh_reset.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.hh_step.argtypes = [C.c_int, C.c_void_p] + [C.c_void_p] * 6
        L.hh_io.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.hh_expert.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.hh_set_spec.argtypes = [C.c_int, C.c_void_p, C.c_int]
        L.hh_replays.restype = C.c_int64
        L.hh_replays.argtypes = [C.c_int, C.c_void_p]
        L.hh_math.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def scalar_math(y, x):
    """qs::m_atan2 / m_sincos / m_rsqrt / m_sqrt (double) -> [n,6]"""
    y = np.ascontiguousarray(y, np.float64)
    x = np.ascontiguousarray(x, np.float64)
    out = np.zeros((y.size, 6), np.float64)
    lib().hh_math(y.size, y.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class HostVecEnv:
    def __init__(self, env_id, n, dtype=np.float32, seed=0, env_id0=0, randomize=0, half_width=0.0, auto_reset=True, params=None,
                 runtime_params=False):
        """params: dict of qs_env_params_t overrides -> the RuntimeParams arithmetic (what use_params = 1 handles run);
        runtime_params=True takes that path with the reference's literals (it must then equal the compiled-in constants)."""
        self.kind, self.n, self.dtype = KINDS[env_id], n, np.dtype(dtype)
        self.f64 = int(self.dtype == np.float64)
        self.obs_dim = 13 if self.kind == 0 else 12
        self._params = None
        if params or runtime_params:
            from oracle import oracle as O  # the oracle's Params mirrors qs_env_params_t field for field (test infrastructure)
            self._params = O.make_params(self.kind, params or {})
        self._h = lib().hh_create(self.f64, self.kind, n, seed, env_id0, randomize, half_width, int(auto_reset),
                                  C.cast(C.pointer(self._params), C.c_void_p) if self._params is not None else None)

    def reset(self):
        obs = np.zeros((self.n, self.obs_dim), self.dtype)
        lib().hh_reset(self.f64, self._h, _p(obs))
        return obs

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=self.dtype)
        obs = np.zeros((self.n, self.obs_dim), self.dtype)
        term = np.zeros((self.n, self.obs_dim), self.dtype)
        rew = np.zeros(self.n, self.dtype)
        done = np.zeros(self.n, np.uint8)
        info = np.zeros(self.n, np.uint8)
        lib().hh_step(self.f64, self._h, _p(a), _p(obs), _p(rew), _p(done), _p(info), _p(term))
        return obs, rew, done, info, term

    def set_speculative(self, on):
        """fp64 docking: False = the guarded arithmetic only (what a replay runs)"""
        lib().hh_set_spec(self.f64, self._h, int(bool(on)))

    def replays(self):
        """fp64 docking: env-steps whose speculative (qs::sdouble) evaluation was redone with the guarded arithmetic"""
        return int(lib().hh_replays(self.f64, self._h))

    def expert_act(self, mem, alias, obs=None):
        """qs::pid_expert_* on every env; mem [n,8] (updated in place), obs [n,12] rewritten for re-seated chasers."""
        act = np.zeros((self.n, 4), self.dtype)
        lib().hh_expert(self.f64, self._h, _p(mem), int(alias), _p(act), _p(obs) if obs is not None else None)
        return act

    FIELDS = {"chaser": (0, 13), "u_c": (13, 17), "target": (17, 30), "u_t": (30, 34), "des_q": (34, 38),
              "last_shaping": (38, 39), "ep_return": (39, 40), "t": (40, 41), "episode": (41, 42)}

    def get_state(self):
        buf = np.zeros((self.n, 42), np.float64)
        lib().hh_io(self.f64, self._h, _p(buf), 0)
        return {k: buf[:, a:b].copy() if b - a > 1 else buf[:, a].copy() for k, (a, b) in self.FIELDS.items()}

    def set_state(self, **kw):
        buf = np.zeros((self.n, 42), np.float64)
        lib().hh_io(self.f64, self._h, _p(buf), 0)
        for k, v in kw.items():
            a, b = self.FIELDS[k]
            buf[:, a:b] = np.asarray(v, dtype=np.float64).reshape(self.n, b - a)
        lib().hh_io(self.f64, self._h, _p(buf), 1)

    def close(self):
        if self._h:
            lib().hh_destroy(self.f64, self._h)
            self._h = None
