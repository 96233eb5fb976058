"""Batched, device-resident drop-in for the reference's gym environments.

Mirrors the stable-baselines VecEnv protocol the reference's trainers drive
(`rl_baselines/ppo2/ppo2.py:472-498`: ``env.reset()``, ``env.step(actions[n,4]) -> obs[n,12], rews[n], dones[n], infos``)
for the env ids `gym-docking/gym_docking/__init__.py:3-21` registers.  Observations, rewards and dones stay on
the GPU as torch tensors; every call goes through the C ABI of libquadsim_b200.so (include/quadsim_b200.h).
"""
import ctypes as C

import numpy as np
import torch

from . import _capi
from .spaces import make_box

MASS, GRAVITY, ARM_LENGTH = 0.18, 9.81, 0.086          # dynamics/quadrotor.py:16-21
MOTOR_LAMBDA = 1.5e-9 / 6.11e-8                         # quadrotor.py:43-45
ROTOR2CONTROL = np.array([[1, 1, 1, 1], [0, ARM_LENGTH, 0, -ARM_LENGTH], [-ARM_LENGTH, 0, ARM_LENGTH, 0],
                          [MOTOR_LAMBDA, -MOTOR_LAMBDA, MOTOR_LAMBDA, -MOTOR_LAMBDA]])  # quadrotor.py:56-59

# id -> (kind, default randomize, default half width)
ENV_IDS = {
    "hovering-v0": (_capi.HOVER, _capi.RAND_ONCE, 1.0),     # hovering_env.py:23-24: sampled once at construction
    "docking-v0": (_capi.DOCK_V0, _capi.RAND_NONE, 0.5),    # docking_env.py:34: jitter commented out
    "docking-v1": (_capi.DOCK_V0, _capi.RAND_ONCE, 0.3),    # imitating_docking_env.py:34: U(-0.3,0.3)^3 once
    "docking-v2": (_capi.DOCK_V2, _capi.RAND_NONE, 0.5),    # moving_docking_env.py:34
}
_RAND = {None: None, False: _capi.RAND_NONE, "none": _capi.RAND_NONE, True: _capi.RAND_EPISODE,
         "episode": _capi.RAND_EPISODE, "once": _capi.RAND_ONCE}


def _space_bounds(kind):
    if kind == _capi.HOVER:  # hovering_env.py:37-41 with Drone.state_lim_* (quadrotor.py:35-39)
        lo = np.array([-100, -100, 0, -100, -100, -100, -100, -100, -100, -100] + [-10 * 2 * np.pi] * 3)
        hi = np.array([100] * 10 + [10 * 2 * np.pi] * 3)
        return lo, hi, np.zeros(4), np.ones(4)
    lo = np.array([-np.inf] * 3 + [-100] * 3 + [-np.pi, -np.pi / 2, -np.pi] + [-10 * np.pi] * 3)  # docking_env.py:85-89
    return lo, -lo, -np.ones(4), np.ones(4)                                                          # :93


class VecInfo:
    """Lazy stand-in for SB's per-env list of info dicts (docking_env.py:226-229): tensors, not N python dicts."""

    def __init__(self, env, bits, terminal_obs, terminal_chaser=None, terminal_target=None):
        self._env, self.bits, self.terminal_observation = env, bits, terminal_obs
        # [N,13] float64 world coordinates; only the rows of envs whose episode ended in this step are meaningful: the
        # reference's info['chaser'] / info['target'] of the terminal step (docking_env.py:226-229)
        self.terminal_chaser, self.terminal_target = terminal_chaser, terminal_target

    @property
    def flag_docking(self):
        return (self.bits & _capi.INFO_DOCKING) != 0

    @property
    def done_overlimit(self):
        return (self.bits & _capi.INFO_OVERLIMIT) != 0

    @property
    def done_overtime(self):
        return (self.bits & _capi.INFO_OVERTIME) != 0

    def __len__(self):
        return self._env.num_envs

    def __getitem__(self, i):
        b = int(self.bits[i])
        d = {"flag_docking": bool(b & _capi.INFO_DOCKING), "done_overlimit": bool(b & _capi.INFO_OVERLIMIT)}
        if self.terminal_observation is not None and (b & _capi.INFO_DONE):
            d["terminal_observation"] = self.terminal_observation[i]
        return d


class QuadVecEnv:
    """N independent QuadSim environments stepped by one CUDA kernel launch (one thread per env)."""

    def __init__(self, env_id="docking-v2", num_envs=1, *, dtype=torch.float32, device=None, seed=0, env_id0=None,
                 randomize=None, half_width=None, auto_reset=True, want_info=True, want_terminal_obs=False,
                 want_terminal_state=False, reuse_buffers=False, params=None):
        """params: dict of qs_env_params_t overrides (mass, inertia, gains, dt, thresholds ...; include/quadsim_b200.h);
        unset members keep the reference's literals."""
        env_id = env_id.split(":")[-1]  # accept 'gym_docking:docking-v2'
        if env_id not in ENV_IDS:
            raise ValueError("unknown env id %r (have %s)" % (env_id, sorted(ENV_IDS)))
        if not torch.cuda.is_available():
            raise _capi.QuadSimError("quadsim_b200 needs a CUDA device: there is no CPU fallback")
        if dtype not in (torch.float32, torch.float64):
            raise ValueError("dtype must be torch.float32 or torch.float64")
        self.lib = _capi.load()
        self.env_id, self.num_envs, self.dtype = env_id, int(num_envs), dtype
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        kind, rand_default, hw_default = ENV_IDS[env_id]
        self.kind = kind
        rand = _RAND[randomize] if not isinstance(randomize, int) or isinstance(randomize, bool) else randomize
        self.randomize = rand_default if rand is None else rand
        self.half_width = float(hw_default if half_width is None else half_width)
        if env_id0 is None:
            env_id0 = 0
            if torch.distributed.is_available() and torch.distributed.is_initialized():
                # global env ids.  Exclusive prefix sum of the ranks' env counts, so unequal shards
                # (distributed.shard_range) neither overlap nor leave gaps in the Philox counter space.
                from .distributed import default_env_id0
                env_id0 = default_env_id0(self.num_envs, self.device)
        self.env_id0 = int(env_id0)
        self.want_info, self.want_terminal_obs, self.reuse_buffers = want_info, want_terminal_obs, reuse_buffers
        self.want_terminal_state = want_terminal_state
        self.params = _capi.default_params(kind)
        for k, v in (params or {}).items():
            if not hasattr(self.params, k) or k == "reserved":
                raise ValueError("unknown env parameter %r (have %s)" % (k, sorted(self.params.as_dict())))
            cur = getattr(self.params, k)
            if isinstance(cur, (int, float)):
                setattr(self.params, k, type(cur)(v))
            else:
                vals = [float(x) for x in np.asarray(v, dtype=np.float64).reshape(-1)]
                if len(vals) != len(cur):
                    raise ValueError("env parameter %r needs %d values" % (k, len(cur)))
                cur[:] = vals
        cfg = _capi.Config(kind, _capi.F32 if dtype == torch.float32 else _capi.F64, self.num_envs, int(seed),
                           self.env_id0, self.randomize, self.device.index, self.half_width, int(bool(auto_reset)),
                           int(bool(params)), self.params)
        h = C.c_void_p()
        _capi.check(self.lib.qs_create(C.byref(h), C.byref(cfg)))
        self._h = h
        self.obs_dim = self.lib.qs_obs_dim(self._h)
        # per-step path: the PyTorch extension (tensors + current stream -> C ABI) when built, else ctypes
        self._ext = _capi.load_ext()
        self._bound = self._ext.bind(h.value) if self._ext is not None else None
        lo, hi, alo, ahi = _space_bounds(kind)
        self.observation_space = make_box(lo, hi, np.float32)
        self.action_space = make_box(alo, ahi, np.float32)
        # what the reference's scripts reach into (run_expert_policy.py:61-69, run_trained_moving_docking_ppo2.py:49)
        mg = self.params.mass * self.params.gravity
        scale = mg if kind == _capi.HOVER else mg / 2.0
        self.action_mean = np.ones(4) * (0.0 if kind == _capi.HOVER else scale)
        self.action_std = np.ones(4) * scale
        self._bufs = None
        self._fast = None  # (action shape, cached output pointers, return tuple): see step()
        self._pending = None
        self._host = None

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _new_bufs(self):
        n, dev = self.num_envs, self.device
        obs = torch.empty((n, self.obs_dim), dtype=self.dtype, device=dev)
        rew = torch.empty((n,), dtype=self.dtype, device=dev)
        done = torch.empty((n,), dtype=torch.uint8, device=dev)
        bits = torch.empty((n,), dtype=torch.uint8, device=dev) if self.want_info else None
        term = torch.empty((n, self.obs_dim), dtype=self.dtype, device=dev) if self.want_terminal_obs else None
        tc = tt = None
        if self.want_terminal_state:
            tc = torch.zeros((n, 13), dtype=torch.float64, device=dev)
            tt = torch.zeros((n, 13), dtype=torch.float64, device=dev) if self.kind != _capi.HOVER else None
        p = lambda t: t.data_ptr() if t is not None else None
        extras = _capi.StepExtras(p(bits), p(term), p(tc), p(tt)) if (bits is not None or term is not None or tc is not None) else None
        return obs, rew, done, bits, term, tc, tt, extras

    def _info(self, bits, term, tc, tt):
        return VecInfo(self, bits, term, tc, tt) if (bits is not None or term is not None or tc is not None) else None

    def _out(self):
        """Output tensors of the next step.  reuse_buffers: False = fresh tensors every step; True / 1 = one set overwritten
        in place; k > 1 = k sets used round-robin (a result stays valid for k - 1 further steps)."""
        if not self.reuse_buffers:
            return self._new_bufs()
        if self._bufs is None:
            self._bufs, self._fast, self._turn = [], [], 0
            for _ in range(max(int(self.reuse_buffers), 1)):
                b = self._new_bufs()
                obs, rew, done, bits, term, tc, tt, extras = b
                ret = (obs, rew, done.view(torch.bool), self._info(bits, term, tc, tt))
                self._bufs.append(b)
                self._fast.append(((self.num_envs, 4), obs.data_ptr(), rew.data_ptr(), done.data_ptr(),
                                   C.byref(extras) if extras is not None else None, ret,
                                   obs, rew, done, C.addressof(extras) if extras is not None else 0))
        self._turn = (self._turn + 1) % len(self._bufs)
        return self._bufs[self._turn]

    @staticmethod
    def _ptr(t):
        return C.c_void_p(t.data_ptr()) if t is not None else None

    def _check_open(self):
        if self._h is None:
            raise _capi.QuadSimError("environment is closed")

    # ------------------------------------------------------------------ VecEnv protocol
    def reset(self):
        """-> obs[N, obs_dim].  Envs that stepped since their last (auto-)reset start a new episode."""
        self._check_open()
        obs = torch.empty((self.num_envs, self.obs_dim), dtype=self.dtype, device=self.device)
        with torch.cuda.device(self.device):
            _capi.check(self.lib.qs_reset(self._h, self._ptr(obs), self._stream()))
        return obs

    def step_async(self, actions):
        self._check_open()
        if isinstance(actions, np.ndarray):
            self._pending = ("host", actions)
            return
        a = actions
        if a.device != self.device or a.dtype != self.dtype:
            a = a.to(device=self.device, dtype=self.dtype)
        a = a.contiguous()
        if a.shape != (self.num_envs, 4):
            raise ValueError("actions must have shape (%d, 4), got %s" % (self.num_envs, tuple(a.shape)))
        obs, rew, done, bits, term, tc, tt, extras = self._out()
        with torch.cuda.device(self.device):
            _capi.check(self.lib.qs_step_ex(self._h, self._ptr(a), self._ptr(obs), self._ptr(rew), self._ptr(done),
                                            C.byref(extras) if extras is not None else None, self._stream()))
        self._pending = ("dev", (obs, rew, done, bits, term, tc, tt, a))

    def step_wait(self):
        kind, payload = self._pending
        self._pending = None
        if kind == "host":
            return self.step_host(payload)
        obs, rew, done, bits, term, tc, tt, _keepalive = payload
        return obs, rew, done.view(torch.bool), self._info(bits, term, tc, tt)

    def step(self, actions):
        """actions[N,4] (torch CUDA tensor, or numpy for the host path) -> obs, rewards, dones, infos."""
        # fast path (reuse_buffers=True, a ready-made device tensor): the output pointers are cached and the call goes
        # straight to the C ABI -- the python side of a step drops from ~35 us to ~10 us, which is what bounds small batches
        # and the 35 us hover kernel
        f = self._fast
        if f is not None:
            self._turn = (self._turn + 1) % len(f)
            f = f[self._turn]
        if f is not None and self._bound is not None and type(actions) is torch.Tensor and actions.dtype == self.dtype and actions.is_cuda:
            # the extension checks device / shape / contiguity itself and takes PyTorch's current stream of the handle's device
            try:
                rc = self._ext.step(self._bound, actions, f[6], f[7], f[8], f[9])
            except RuntimeError:
                rc = None  # another device / a view / a wrong shape: the general path below converts or explains
            if rc is not None:
                if rc != 0:
                    _capi.check(rc)
                return f[5]
        if (f is not None and type(actions) is torch.Tensor and actions.dtype == self.dtype and actions.device == self.device
                and actions.shape == f[0] and actions.is_contiguous() and self._h is not None):
            if torch.cuda.current_device() == self.device.index:
                rc = self.lib.qs_step_ex(self._h, actions.data_ptr(), f[1], f[2], f[3], f[4],
                                         torch.cuda.current_stream().cuda_stream)
                if rc != 0:
                    _capi.check(rc)
                return f[5]
        self.step_async(actions)
        return self.step_wait()

    def step_many(self, actions):
        """T consecutive steps in one kernel launch (qs_step_n): actions[T,N,4] -> obs[T,N,obs_dim], rewards[T,N], dones[T,N]
        (time-major device tensors).  The arithmetic of T step() calls (same parity bars); the env state never leaves
        registers in between."""
        self._check_open()
        a = actions
        if a.device != self.device or a.dtype != self.dtype:
            a = a.to(device=self.device, dtype=self.dtype)
        a = a.contiguous()
        if a.dim() != 3 or tuple(a.shape[1:]) != (self.num_envs, 4):
            raise ValueError("actions must have shape (T, %d, 4), got %s" % (self.num_envs, tuple(a.shape)))
        T, n, dev = a.shape[0], self.num_envs, self.device
        obs = torch.empty((T, n, self.obs_dim), dtype=self.dtype, device=dev)
        rew = torch.empty((T, n), dtype=self.dtype, device=dev)
        done = torch.empty((T, n), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _capi.check(self.lib.qs_step_n(self._h, T, self._ptr(a), self._ptr(obs), self._ptr(rew), self._ptr(done), self._stream()))
            if a is not actions:
                torch.cuda.current_stream(dev).synchronize()  # the converted copy must outlive the kernel
        return obs, rew, done.view(torch.bool)

    def pinned_host_buffers(self, n_action_buffers=1):
        """Pinned numpy staging buffers for the host path: ([actions...], obs, rewards, dones).  Passing these to
        step_host avoids every extra host-side copy (the DMA engines read/write them directly)."""
        npdt = self.dtype
        mk = lambda *shape, dt=npdt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()
        acts = [mk(self.num_envs, 4) for _ in range(n_action_buffers)]
        return acts, mk(self.num_envs, self.obs_dim), mk(self.num_envs), mk(self.num_envs, dt=torch.uint8)

    def step_host(self, actions, out=None):
        """numpy in / numpy out through qs_step_host: H2D of the actions, the step kernel and D2H of obs/reward/done are
        chunked over internal streams so they overlap.  `out=(obs, rew, done)` numpy buffers are written in place (use
        pinned_host_buffers() for full-speed DMA); without `out`, fresh arrays are returned."""
        self._check_open()
        npdt = np.float32 if self.dtype == torch.float32 else np.float64
        a = np.asarray(actions)
        if a.shape != (self.num_envs, 4):
            raise ValueError("actions must have shape (%d, 4), got %s" % (self.num_envs, a.shape))
        if a.dtype != npdt or not a.flags["C_CONTIGUOUS"]:
            a = np.ascontiguousarray(a, dtype=npdt)
        if out is None:
            if self._host is None:
                self._host = self.pinned_host_buffers()[1:]
            ho, hr, hd = self._host
        else:
            ho, hr, hd = out
            if (ho.shape != (self.num_envs, self.obs_dim) or ho.dtype != npdt or hr.shape != (self.num_envs,)
                    or hr.dtype != npdt or hd.shape != (self.num_envs,) or hd.dtype != np.uint8
                    or not (ho.flags["C_CONTIGUOUS"] and hr.flags["C_CONTIGUOUS"] and hd.flags["C_CONTIGUOUS"])):
                raise ValueError("out buffers must be C-contiguous (obs[N,%d], rew[N]) of %s and done[N] uint8" % (self.obs_dim, npdt))
        vp = lambda x: C.c_void_p(x.ctypes.data)
        with torch.cuda.device(self.device):
            _capi.check(self.lib.qs_step_host(self._h, vp(a), vp(ho), vp(hr), vp(hd), self._stream()))
        if out is None:
            return ho.copy(), hr.copy(), hd.astype(bool), None
        return ho, hr, hd.view(np.bool_), None

    def close(self):
        if getattr(self, "_h", None) is not None:
            if getattr(self, "_bound", None) is not None:
                self._ext.unbind(self._bound)
                self._bound = None
            self.lib.qs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def seed(self, seed=None):
        """env.seed (docking_env.py:252-254): re-keys the Philox stream used for randomised (re)starts."""
        self._check_open()
        _capi.check(self.lib.qs_seed(self._h, int(0 if seed is None else seed)))
        return [seed] * min(self.num_envs, 1)

    def render(self, mode="human"):
        return None

    def get_attr(self, name, indices=None):
        n = self.num_envs if indices is None else len(np.atleast_1d(indices))
        return [getattr(self, name)] * n

    def set_attr(self, name, value, indices=None):
        setattr(self, name, value)

    def env_method(self, name, *args, indices=None, **kwargs):
        raise NotImplementedError("per-env python methods do not exist on the batched CUDA env")

    # ------------------------------------------------------------------ state access / checkpoint
    _FIELDS = (("chaser", 13), ("target", 13), ("u_chaser", 4), ("u_target", 4), ("des_q", 4), ("last_shaping", 0),
               ("ep_return", 0), ("t", 0), ("episode", 0))
    _DOCK_ONLY = ("target", "u_target", "des_q", "last_shaping")

    def _arrays(self, tensors):
        sa = _capi.StateArrays()
        for name, _ in self._FIELDS:
            t = tensors.get(name)
            setattr(sa, name, t.data_ptr() if t is not None else None)
        return sa

    def get_state(self):
        """dict of device tensors (float64 whatever the env dtype; t/episode int32): env.chaser.state, env.target.state,
        Drone.u x2, target_state_des[6:10], last_shaping, ep_return.  get_state -> set_state is bit-transparent."""
        self._check_open()
        out = {}
        for name, w in self._FIELDS:
            if self.kind == _capi.HOVER and name in self._DOCK_ONLY:
                continue
            dt = torch.int32 if name in ("t", "episode") else torch.float64
            shape = (self.num_envs, w) if w else (self.num_envs,)
            out[name] = torch.empty(shape, dtype=dt, device=self.device)
        sa = self._arrays(out)
        with torch.cuda.device(self.device):
            _capi.check(self.lib.qs_export_state(self._h, C.byref(sa), self._stream()))
        return out

    def set_state(self, state):
        """Inverse of get_state; missing keys are left untouched."""
        self._check_open()
        conv = {}
        for name, w in self._FIELDS:
            if name not in state or state[name] is None:
                continue
            if self.kind == _capi.HOVER and name in self._DOCK_ONLY:
                raise ValueError("%s does not exist in hovering-v0" % name)
            dt = torch.int32 if name in ("t", "episode") else torch.float64
            t = torch.as_tensor(state[name]).to(device=self.device, dtype=dt).contiguous()
            shape = (self.num_envs, w) if w else (self.num_envs,)
            if tuple(t.shape) != shape:
                raise ValueError("%s must have shape %s, got %s" % (name, shape, tuple(t.shape)))
            conv[name] = t
        sa = self._arrays(conv)
        with torch.cuda.device(self.device):
            _capi.check(self.lib.qs_import_state(self._h, C.byref(sa), self._stream()))
            torch.cuda.current_stream(self.device).synchronize()  # conv tensors must outlive the kernel

    STAT_NAMES = ("return_sum", "length_sum", "episodes", "docked_steps", "nonfinite_steps")

    def stats(self, all_reduce=False):
        """Episode statistics since the last call (device-accumulated); optionally summed over ranks (32+8 bytes)."""
        self._check_open()
        buf = (C.c_double * 5)()
        with torch.cuda.device(self.device):
            _capi.check(self.lib.qs_get_stats(self._h, buf, self._stream()))
        vals = torch.tensor(list(buf), dtype=torch.float64)
        if all_reduce and torch.distributed.is_available() and torch.distributed.is_initialized():
            from .distributed import all_reduce_stats
            vals = all_reduce_stats(vals, self.device)
        d = dict(zip(self.STAT_NAMES, vals.tolist()))
        d["mean_return"] = d["return_sum"] / d["episodes"] if d["episodes"] else float("nan")
        d["mean_length"] = d["length_sum"] / d["episodes"] if d["episodes"] else float("nan")
        return d
